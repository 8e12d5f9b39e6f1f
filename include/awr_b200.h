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

/* ---- evaluation of predicted joints (util/eval_tool.py:20-122, util/util.py:13-20) -------------------------------- */

/* EvalUtil.feed for a batch.  uvd_pred (B,J,3) normalised crop coordinates; xyz_gt_norm (B,J,3) cube-normalised GT; center_xyz (B,3) mm;
 * M (B,3,3) crop affine; cube (B,3) mm; vis (B,J) bytes or NULL (all visible).  Outputs: uvd_img (B,J,3) = pixel u, v in the original
 * image + depth in mm (what test.py:103-108 writes to results/*.txt); dist (B,J) Euclidean error in mm, -1 where not visible;
 * diff_mean (B,3) or NULL = mean signed error per frame (eval_tool.py:50). */
int awr_eval_feed(const float* uvd_pred, const float* xyz_gt_norm, const float* center_xyz, const float* M, const float* cube,
                  const unsigned char* vis, int B, int J, float img_size, float fx, float fy, float fu, float fv, float flip, float* uvd_img,
                  float* dist, float* diff_mean, void* stream);

/* Reductions of EvalUtil.get_measures over dist (N,J) (entries < 0 ignored): sum[J] (double) and count[J] of the errors per joint and
 * pck_count[J][nthr] = number of errors <= linspace(0, thr_max, nthr)[k]. */
int awr_eval_measures(const float* dist, long long N, int J, int nthr, float thr_max, double* sum, unsigned* count, unsigned* pck_count,
                      void* stream);

/* ---- depth preprocessing (dataloader/loader.py:19-51,88-101,190-207; nyu_loader.py:71-74) -------------------------- */

/* Loader.crop + Loader.normalize of the non-augmented path for N raw frames: src (N,Hs,Ws) float32 millimetres (src_format 0) or
 * (N,Hs,Ws,3) uint8 BGR with depth = B + 256*G (src_format 1); params (N,12) doubles = crop-box geometry computed on the host
 * (layout in csrc/preprocess.cu, built by awr_b200/preprocess.py); out (N,1,img_size,img_size) float32 in [-1,1], background +1. */
int awr_crop_normalize(const void* src, int src_format, int N, int Hs, int Ws, const double* params, int img_size, float* out,
                       void* stream);

/* The training data path for N raw frames in one launch: Loader.crop, then Loader.augment's image half (dataloader/loader.py:74-86 --
 * translate / scale = Loader.recrop :125-139 (cv2.warpPerspective INTER_LINEAR, BORDER_CONSTANT 0, pixels below min(depth>0)-1 dropped,
 * cube clamp), rotate :141-161 (cv2.warpAffine INTER_LINEAR)), then Loader.normalize (:88-101) -- bit-identical to the reference running
 * cv2 4.13.  params (N,32) doubles: [0..11] as awr_crop_normalize (10/11 = the centre z and half cube AFTER augmentation), 12 = op
 * (0 none, 1 perspective, 2 affine), 13..21 = the INVERSE map (row-major 3x3; affine: 13..18), 22/23 = cube front/back after
 * augmentation, 24 = tile width of cv2's perspective loop (min(64, img_size) for img_size >= 16); built by awr_b200/preprocess.py
 * (train_frame_geometry).  One 8-CTA cluster per frame; img_size <= 640.  A frame without any positive depth yields all background
 * (the reference raises on it). */
int awr_crop_augment_normalize(const void* src, int src_format, int N, int Hs, int Ws, const double* params, int img_size, float* out,
                               void* stream);

/* ---- NHWC elementwise / normalisation kernels (storage dtype: AWR_DTYPE_F32 or AWR_DTYPE_BF16) ---------------
 * Internal activation layout is NHWC (M = N*H*W pixels x C channels, C a power of two in [64,2048] for the
 * per-channel reductions).  These replace nn.BatchNorm2d / nn.ReLU / residual adds / nn.MaxPool2d / nn.Upsample
 * of model/resnet_deconv.py:31-36,145-215 and model/hourglass.py:28-88. */

/* Per-channel accumulators shared by many CTAs (BatchNorm batch statistics `sums`, their backward sums `dsums`) are
 * ORDER-INDEPENDENT: awr_acc_t[n], value = hi * 2^-24 + lo * 2^-72, filled with 64-bit integer atomics (every fp32 partial, |p| < 2^39,
 * is split exactly), so the totals -- and with them the whole step's BatchNorm path -- are bit-reproducible run to run, which fp32
 * atomics are not (the reference is bit-deterministic on CPU).  16 bytes per value, zero-filled by the caller. */
typedef struct { long long hi, lo; } awr_acc_t;

/* sums[0:C] += sum_m x[m,c];  with_sq: sums[C:2C] += sum_m x[m,c]^2   (sums: awr_acc_t[C or 2C], caller zero-fills).
 * out_f32 + counter (both or neither; counter: one zero-initialised unsigned): the CTA finishing last ADDS the totals to the fp32
 * array out_f32 (conv bias gradients inside the flat gradient buffer; single writer => reproducible) and re-arms the counter. */
int awr_channel_stats(const void* x, int dtype, long long M, int C, void* sums, int with_sq, float* out_f32, unsigned* counter, void* stream);

/* BatchNorm2d statistics -> per-channel affine.  training: batch stats from sums/count, running stats updated with
 * `momentum` (unbiased var), *num_batches_tracked += 1;  eval: running stats.  scale_shift[2C]; mean_invstd[2C] or NULL. */
int awr_bn_finalize(const void* sums, long long count, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, long long* num_batches_tracked, float* scale_shift, float* mean_invstd, int C,
                    float momentum, float eps, int training, void* stream);

/* One-launch BatchNorm2d forward: out = act( BN(y) [+ res | + BN_res(res)] ).  training: batch statistics from sums[2C] = {sum, sum of
 * squares} over the M pixels (produced by awr_channel_stats or by awr_conv_tc's epilogue), running statistics updated (momentum,
 * unbiased variance), *num_batches_tracked += 1; eval: running statistics.  mean_invstd[2C] (or NULL) is saved for backward.
 * The res_* set (all NULL when absent) is the BatchNorm of the residual branch (ResNet downsample path, resnet_deconv.py:62-68). */
int awr_bn_act(const void* y, const void* sums, const float* gamma, const float* beta, float* running_mean, float* running_var,
               long long* num_batches_tracked, float* mean_invstd, const void* res, const void* res_sums, const float* res_gamma,
               const float* res_beta, float* res_running_mean, float* res_running_var, long long* res_num_batches_tracked,
               float* res_mean_invstd, void* out, int dtype, long long M, int C, float momentum, float eps, int training, int relu,
               void* stream);

/* Fused stem tail (resnet_deconv.py:33-35): out = MaxPool_{k,s,p}(ReLU(BN(y))) read from the raw conv output in ONE pass; idx = arg-max
 * tap byte per pooled element.  BN arguments as in awr_bn_act.  The full-resolution normalised tensor is never written. */
int awr_bn_relu_maxpool_fwd(const void* y, const void* sums, const float* gamma, const float* beta, float* running_mean, float* running_var,
                            long long* num_batches_tracked, float* mean_invstd, void* out, unsigned char* idx, int dtype, int N, int H, int W,
                            int C, int k, int s, int p, float momentum, float eps, int training, void* stream);
/* Backward of the above straight to dy (gradient of the raw conv output): pass 0 accumulates dsums[2C] = {sum dz, sum dz*yhat} (caller
 * zero-fills), pass 1 writes dy and dgamma/dbeta.  dz is rebuilt from dpool + idx (gather over <= 4 windows) and the ReLU mask from y. */
int awr_maxpool_bn_bwd(const void* dpool, const unsigned char* idx, const void* y, const float* mean_invstd, const float* gamma,
                       const float* beta, void* dsums, void* dy, float* dgamma, float* dbeta, int dtype, int N, int H, int W, int C, int k,
                       int s, int p, int pass, int accumulate_param_grads, void* stream);

/* Pass 0 of awr_maxpool_bn_bwd at pooled resolution (any k/s/p): dsums[0:C] += sum dpool*[pool_out>0], dsums[C:2C] += sum dpool*[pool_out>0]*
 * (pool_out-beta)/gamma, which equal sum dz and sum dz*yhat of the full-resolution pass because the pooled value is the activation at the
 * arg-max pixel.  dpool / pool_out: (M = N*Ho*Wo, C) NHWC. */
int awr_pool_bn_bwd_reduce(const void* dpool, const void* pool_out, const float* gamma, const float* beta, void* dsums, int dtype, long long M,
                           int C, void* stream);

/* out = act( ss(y) + res_ss(res) ),  ss(v)[c] = v*scale[c] + shift[c]; scale_shift / res / res_scale_shift may be NULL. */
int awr_affine_act(const void* y, const float* scale_shift, const void* res, const float* res_scale_shift, void* out, int dtype,
                   long long M, int C, int relu, void* stream);

/* BN backward pass 1: dz = dout*mask; dsums[0:C]+=sum dz, dsums[C:2C]+=sum dz*yhat.  ReLU mask: act_out>0 when act_out is given; or, for
 * ReLU(BN(y)) WITHOUT residual, recomputed from y alone when mask_gamma/mask_beta (the BN affine) are given -- one tensor read less;
 * all three NULL: no ReLU. */
int awr_bn_bwd_reduce(const void* dout, const void* act_out, const void* y, const float* mean_invstd, const float* mask_gamma,
                      const float* mask_beta, int dtype, long long M, int C, void* dsums, void* stream);
/* BN backward pass 2 (mask_beta non-NULL: ReLU mask recomputed from y with gamma/mask_beta instead of reading act_out):
 * dy = bn_grad [+ dy_addend]; optional dres = dz [+ dres_addend] (addends may alias their outputs);
 * dgamma/dbeta (NULL to skip) written or accumulated. */
int awr_bn_bwd_apply(const void* dout, const void* act_out, const void* y, const float* mean_invstd, const void* dsums,
                     const float* gamma, void* dy, const void* dy_addend, void* dres, const void* dres_addend, float* dgamma,
                     float* dbeta, const float* mask_beta, int dtype, long long M, int C, int accumulate_param_grads, void* stream);

/* BN backward in ONE launch (both passes above) for tensors whose operands fit in the shared memory of one CTA per SM: every CTA stages its
 * slice of dout / y (/ act_out) with 1-D bulk TMA copies, reduces, meets the other CTAs at a grid-wide barrier and writes dy / dres from the
 * shared-memory copy.  `barrier`: one zero-initialised unsigned per call (re-zero before every launch; the plan keeps it in the per-step
 * zeroed arena).  ReLU mask: recomputed from y when mask_beta is given (gamma is then the BN weight), else act_out > 0 when act_out is given,
 * else none.  awr_bn_bwd_fused returns AWR_ERR_UNSUPPORTED when the operands do not fit; awr_bn_bwd_fused_ok returns 1 when the tensor
 * qualifies AND the path is enabled (AWR_BN_FUSED=1; measured slower than reduce + apply at the headline sizes, so off by default).
 * The launch needs all its CTAs co-resident (grid <= SM count, one CTA per SM). */
int awr_bn_bwd_fused_ok(long long M, int C, int dtype, int with_act);
int awr_bn_bwd_fused(const void* dout, const void* act_out, const void* y, const float* mean_invstd, const float* gamma, const float* mask_beta,
                     void* dsums, unsigned* barrier, void* dy, const void* dy_addend, void* dres, const void* dres_addend, float* dgamma,
                     float* dbeta, int dtype, long long M, int C, int accumulate_param_grads, void* stream);

/* dx = dout*(act_out>0) [+ addend]   (act_out / addend may be NULL); n elements, n % 8 == 0 */
int awr_relu_bwd(const void* dout, const void* act_out, const void* addend, void* dx, int dtype, long long n, void* stream);

/* MaxPool2d(k,s,p) NHWC; idx (uint8 per output element, arg-max tap, first max in scan order) may be NULL in fwd. */
int awr_maxpool_fwd(const void* x, void* out, unsigned char* idx, int dtype, int N, int H, int W, int C, int k, int s, int p,
                    void* stream);
int awr_maxpool_bwd(const void* dout, const unsigned char* idx, void* dx, int dtype, int N, int H, int W, int C, int k, int s, int p,
                    int accumulate, void* stream);

/* out = up + nearest_upsample_x2(low)  (hourglass.py:87-88); H,W are the fine (output) size.  bwd: dlow = 2x2 block sums. */
int awr_upsample2_add(const void* up, const void* low, void* out, int dtype, int N, int H, int W, int C, void* stream);
int awr_upsample2_bwd(const void* dout, void* dlow, int dtype, int N, int H, int W, int C, int accumulate, void* stream);

/* NCHW fp32 (N,Csrc,P) <-> NHWC dtype (N,P,C): to_nhwc zero-pads channels up to Cdst; to_nchw keeps the first Cdst. */
int awr_nchw_to_nhwc(const float* src, void* dst, int dtype, int N, int Csrc, int Cdst, int P, void* stream);
int awr_nhwc_to_nchw(const void* src, float* dst, int dtype, int N, int Csrc, int Cdst, int P, void* stream);

/* torch.optim.Adam (train.py:67) fused over a flat fp32 buffer; g is multiplied by grad_scale (1/world_size under DP);
 * bf16_shadow (or NULL) receives the rounded parameters.  step_dev: device float holding the 1-based step count. */
int awr_adam_flat(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n, const float* step_dev, float lr,
                  float beta1, float beta2, float eps, float weight_decay, float grad_scale, void* stream);
int awr_adam_tick(float* step_dev, void* stream);
/* The same Adam step / torch.optim.SGD(momentum, dampening 0; train.py:69) with the schedule-driven hyper-parameters in DEVICE memory:
 * hyper_dev = [step count (1-based, as float; advance it with awr_adam_tick), learning rate], so a captured CUDA graph follows
 * StepLR / ReduceLROnPlateau (train.py:89-92,157-160) without re-capture.  skip_spans_dev (or NULL): n_skip <= 256 pairs [begin, end)
 * of flat indices (multiples of 4) that the step leaves untouched -- parameters whose gradient is None in the reference (the
 * Hourglass skip_layer convs forward never calls, model/hourglass.py:38,45-48), which torch.optim skips.
 * zero_grad != 0: g is zero-filled once consumed (the next step's weight-gradient kernels accumulate into it), which replaces a separate
 * 4*n-byte fill per step. */
int awr_optim_adam(float* p, float* g, float* m, float* v, void* bf16_shadow, long long n, const float* hyper_dev, float beta1,
                   float beta2, float eps, float weight_decay, float grad_scale, const long long* skip_spans_dev, int n_skip, int zero_grad,
                   void* stream);
int awr_optim_sgd(float* p, float* g, float* momentum_buf, void* bf16_shadow, long long n, const float* hyper_dev, float momentum,
                  float weight_decay, float grad_scale, const long long* skip_spans_dev, int n_skip, int zero_grad, void* stream);
/* Deterministic build (awr_deterministic() == 1, libawr_b200_det.so): the weight-gradient kernels (awr_conv_wgrad_tc / _simt,
 * awr_stem_wgrad) take dW / dbias as awr_acc_t slots instead of floats -- same element offsets -- so split-K partial sums combine
 * order-independently; this folds n slots into the fp32 gradients (grads[i] += value) and re-zeroes them.  In the default build the
 * kernels add fp32 atomics straight into the float gradient buffer and this call is not needed. */
int awr_deterministic(void);
int awr_grad_acc_finalize(void* acc, float* grads, long long n, void* stream);
/* cudaMemsetAsync(p, 0, nbytes) on `stream` (a memset node when captured): the per-step zero fill of the accumulator arena. */
int awr_memset_zero(void* p, long long nbytes, void* stream);
/* Grid cap (SMs) of the persistent tensor-core kernels launched from now on, 8..148; returns the previous value.  The data-parallel
 * trainer lowers it for the launches that overlap a gradient bucket's NCCL all-reduce, so the collective finds free SMs. */
int awr_set_sm_budget(int n);
int awr_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);

/* ---- convolutions, CUDA-core fp32-accumulate path (fp32 precision mode; also the 1-channel stem) --------------
 * Activations NHWC of `dtype`; weights fp32 in the physical order [kh][kw][Cout][Cin] (the canonical OIHW
 * nn.Parameter is a permuted view of it; ConvTranspose2d IOHW likewise).  Replaces nn.Conv2d / nn.ConvTranspose2d
 * forward + autograd (resnet_deconv.py:31-32,78-86,141-142,182-188; hourglass.py:10). */

/* Generic gather convolution  out[m,n] = sum_{tap,k} in[gather(m,tap),k] * w[tap*w_tap + k*w_sk + n*w_sn] (+bias[n]).
 *   transposed=0: input pixel = out*stride - pad + tap        (Conv2d fprop; ConvTranspose2d dgrad)
 *   transposed=1: input pixel = (out + pad - tap)/stride      (ConvTranspose2d fprop; Conv2d dgrad)
 *   fprop contracts Cin (w_sk=1, w_sn=Cin); dgrad contracts Cout (w_sk=Cin, w_sn=1).
 *   out_mode 0: NHWC `dtype` (accumulate=1 adds into out); out_mode 1: NCHW fp32, first n_valid channels. */
int awr_conv_simt(const void* in, const float* w, const float* bias, void* out, int dtype, int N, int Hi, int Wi, int Ck, int Ho,
                  int Wo, int Cn, int R, int S, int stride, int pad, int transposed, int w_sk, int w_sn, int w_tap, int out_mode,
                  int n_valid, int accumulate, void* stream);

/* Weight gradient  dW[tap*w_tap + i*s_p + j*s_g] += sum_q pointwise[q,i] * gathered[q*stride - pad + tap, j]
 * (q over the coarse grid N x Hc x Wc).  Conv2d: pointwise=dy, gathered=x, s_p=Cin, s_g=1.
 * ConvTranspose2d: pointwise=x, gathered=dy, s_p=1, s_g=Cin.  dW fp32, caller zero-fills. */
int awr_conv_wgrad_simt(const void* pointwise, const void* gathered, void* dW, int dtype, int N, int Hc, int Wc, int Cp, int Hf, int Wf,
                        int Cg, int R, int S, int stride, int pad, int s_p, int s_g, int w_tap, void* stream);

/* 1-channel k x k stride-1 'same' stem convolution: x (N,H,W) fp32, w [k*k][Cout] fp32, bias or NULL -> y NHWC.
 * stats (or NULL; k = 5 only): awr_acc_t[2*Cout] += per-channel sum / sum of squares of the stored outputs (BatchNorm statistics). */
int awr_stem_conv(const float* x, const float* w, const float* bias, void* y, void* stats, int dtype, int N, int H, int W, int Cout, int k,
                  void* stream);
/* dW[k*k][Cout] += ..., dbias[Cout] += ... (dbias may be NULL); caller zero-fills. */
int awr_stem_wgrad(const float* x, const void* dy, void* dW, void* dbias, int dtype, int N, int H, int W, int Cout, int k,
                   void* stream);

/* ---- convolutions, tcgen05 tensor-core path (bf16 precision mode) ------------------------------------------------
 * NHWC bf16 activations, bf16 weights (the Adam kernel's shadow copy, same physical order [kh][kw][Cout][Cin]), fp32
 * accumulation in TMEM, operands staged by TMA (shifted boxes = implicit im2col, zero OOB fill = padding).
 * Same argument meaning as awr_conv_simt (w strides select fprop [w_sk==1] or dgrad [w_sn==1]); Ck, Cn multiples of 64,
 * feature-map sides powers of two <= 256, stride 1 or 2.  stats (or NULL): awr_acc_t[2*Cn], the epilogue adds the per-channel sum and
 * sum of squares of the (bf16-rounded) outputs -- the BatchNorm batch statistics -- so no separate reduction pass is needed
 * (caller zero-fills).  Returns AWR_ERR_DRIVER (-3) if the driver cannot encode a tensor map. */
int awr_conv_tc(const void* in, const void* w, const float* bias, void* out, void* stats, int N, int Hi, int Wi, int Ck, int Ho, int Wo,
                int Cn, int R, int S, int stride, int pad, int transposed, int w_sk, int w_sn, int w_tap, int out_mode, int n_valid,
                int accumulate, void* stream);

/* Weight gradient on tensor cores; same argument meaning as awr_conv_wgrad_simt (bf16 NHWC operands, fp32 dW accumulated with
 * atomic adds: caller zero-fills).  The contraction runs over pixels; both operands are MN-major TMA boxes of the NHWC tensors. */
int awr_conv_wgrad_tc(const void* pointwise, const void* gathered, void* dW, int N, int Hc, int Wc, int Cp, int Hf, int Wf, int Cg, int R,
                      int S, int stride, int pad, int s_p, int s_g, int w_tap, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AWR_B200_H */
