/* libawr_b200.so -- C-ABI of the B200-native AWR hot path.
 *
 * The reference (Elody-07/AWR-Adaptive-Weighting-Regression) has no FFI layer: its boundary is
 * four Python symbols (train.py:13-17).  These entry points are what a ctypes binding for that
 * path binds; awr_b200/_lib.py is that binding and INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; tensors are dense, row-major;
 *   - `stream` is a cudaStream_t passed as void*; no call synchronises, allocates or frees;
 *   - return 0 on success, >0 = cudaError_t of the launch, <0 = AWR_ERR_* argument errors;
 *   - all calls are re-entrant and CUDA-graph capturable.
 */
#ifndef AWR_B200_H
#define AWR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define AWR_B200_VERSION 100

#define AWR_DTYPE_F32 0
#define AWR_DTYPE_BF16 1

#define AWR_HUBER_MAX_BLOCKS 1184 /* 148 SMs x 8 */

int awr_version(void);

/* ---- adaptive-weighting head + SmoothL1 (util/feature_tool.py:12-65, model/loss.py:8-25) ------------------ */

/* Workspace size in floats for awr_head_fwd/bwd: softmax stats [B*J][2] + loss partials [B*J][2] + ticket.
 * Must be zero-filled once before first use (the ticket self-resets). */
#define AWR_HEAD_WS_FLOATS(B, J) (4 * (B) * (J) + 4)

/* FeatureModule.offset2joint_softmax (feature_tool.py:41-65), optionally fused with both SmoothL1 terms of
 * train.py:119-120.  pred: (B,4J,F,F) NCHW, fp32 or bf16.  img: (B,1,H,H) fp32 (H % F == 0, nearest resample).
 * uvd_gt: (B,J,3) or NULL.  uvd_out: (B,J,3).  loss_out: [2] = {SmoothL1(uvd,uvd_gt), SmoothL1(pred, joint2offset(uvd_gt))}
 * (unweighted means) or NULL. */
int awr_head_fwd(const void* pred, int pred_dtype, const float* img, const float* uvd_gt, float* uvd_out, float* loss_out,
                 float* ws, int B, int J, int F, int H, float kernel_size, void* stream);

/* Backward of the above w.r.t. pred (fp32 NCHW out).  uvd / ws are the forward's outputs.
 *   g_uvd  (B,J,3) or NULL : upstream gradient of the UVD output (autograd use);
 *   uvd_gt (B,J,3) or NULL : when given adds coord_weight*dSmoothL1(uvd,uvd_gt) and dense_weight*dSmoothL1(pred,GT volume),
 *                            both scaled by *loss_grad (device scalar, NULL == 1). */
int awr_head_bwd(const void* pred, int pred_dtype, const float* img, const float* uvd_gt, const float* uvd, const float* ws,
                 const float* g_uvd, const float* loss_grad, float* dpred, int B, int J, int F, int H, float kernel_size,
                 float coord_weight, float dense_weight, void* stream);

/* FeatureModule.joint2offset (feature_tool.py:12-39): jt_uvd (B,J,3), img (B,1,H,H) -> out (B,4J,F,F) fp32. */
int awr_joint2offset(const float* jt_uvd, const float* img, float* out, int B, int J, int F, int H, float kernel_size,
                     void* stream);

/* My_SmoothL1Loss.forward (loss.py:8-19): mean Huber(delta=0.01) over n elements.
 * ws: 2*AWR_HUBER_MAX_BLOCKS + 4 floats, zero-filled once. out: device scalar. */
int awr_huber_fwd(const float* x, const float* y, long long n, float* ws, float* out, void* stream);
/* d/dx of the above times *grad_out (device scalar, NULL == 1). */
int awr_huber_bwd(const float* x, const float* y, long long n, const float* grad_out, float* dx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AWR_B200_H */
