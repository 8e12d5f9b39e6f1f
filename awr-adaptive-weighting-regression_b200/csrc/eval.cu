// Device-side evaluation of predicted joints: the per-sample numpy loop of the reference's
// util/eval_tool.py::EvalUtil.feed (:20-58) + util/util.py::uvd2xyz (:13-20) for a whole batch, and the reductions of
// EvalUtil.get_measures (:60-122).  Removes the B x 5 `.cpu()` round trips per training step of train.py:141-148.
#include "common.cuh"
#include "awr_b200.h"

namespace {

// One warp per frame.  Arithmetic follows the reference's dtypes: float32 where numpy stays in float32 (eval_tool.py:38-39,45,48-49),
// double where numpy promotes (the float64 homogeneous dot :40-41, uvd2xyz with python-float intrinsics util.py:16), values rounded to
// float32 where the reference stores into float32 arrays.
__global__ void __launch_bounds__(128) eval_feed_kernel(const float* __restrict__ uvd_pred, const float* __restrict__ xyz_gt_norm,
                                                        const float* __restrict__ center_xyz, const float* __restrict__ Mat,
                                                        const float* __restrict__ cube, const unsigned char* __restrict__ vis, int B, int J,
                                                        float img_size, float fx, float fy, float fu, float fv, float flip,
                                                        float* __restrict__ uvd_img, float* __restrict__ dist, float* __restrict__ diff_mean) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const int b = warp;
  // M^-1 (eval_tool.py:33) in double from the adjugate, then rounded to float32 like np.linalg.inv of a float32 matrix
  const float* m = Mat + (size_t)b * 9;
  const double a = m[0], bb = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double det = a * (e * i - f * h) - bb * (d * i - f * g) + c * (d * h - e * g);
  const double id = 1.0 / det;
  const float i00 = (float)((e * i - f * h) * id), i01 = (float)((c * h - bb * i) * id), i02 = (float)((bb * f - c * e) * id);
  const float i10 = (float)((f * g - d * i) * id), i11 = (float)((a * i - c * g) * id), i12 = (float)((c * d - a * f) * id);
  const float cx = center_xyz[b * 3], cy = center_xyz[b * 3 + 1], cz = center_xyz[b * 3 + 2];
  const float hx = cube[b * 3] / 2.f, hy = cube[b * 3 + 1] / 2.f, hz = cube[b * 3 + 2] / 2.f;
  float sdx = 0.f, sdy = 0.f, sdz = 0.f;
  for (int j = lane; j < J; j += 32) {
    const size_t o = ((size_t)b * J + j) * 3;
    const float u = (uvd_pred[o] + 1.f) * img_size / 2.f, v = (uvd_pred[o + 1] + 1.f) * img_size / 2.f;      // :38
    const float dep = uvd_pred[o + 2] * cube[b * 3 + 2] / 2.f + cz;                                           // :39
    const float u0 = (float)((double)i00 * u + (double)i01 * v + (double)i02);                                // :40-41 (float64 dot, float32 store)
    const float v0 = (float)((double)i10 * u + (double)i11 * v + (double)i12);
    uvd_img[o] = u0; uvd_img[o + 1] = v0; uvd_img[o + 2] = dep;                                               // :42 (what test.py:103-108 saves)
    const float x = (float)(((double)u0 - (double)fu) * (double)dep / (double)fx);                            // util.py:16
    const float y = (float)(((double)v0 - (double)fv) * (double)dep / (double)fy) * flip;                     // util.py:16-17
    const float gx = xyz_gt_norm[o] * hx + cx, gy = xyz_gt_norm[o + 1] * hy + cy, gz = xyz_gt_norm[o + 2] * hz + cz;   // :45
    const float dx = gx - x, dy = gy - y, dz = gz - dep;                                                      // :48
    sdx += dx; sdy += dy; sdz += dz;
    const float dd = sqrtf(dx * dx + dy * dy + dz * dz);                                                      // :49
    dist[(size_t)b * J + j] = (vis == nullptr || vis[(size_t)b * J + j]) ? dd : -1.f;                         // :53-58 (not visible: skipped)
  }
  sdx = warp_sum(sdx); sdy = warp_sum(sdy); sdz = warp_sum(sdz);
  if (lane == 0 && diff_mean) { diff_mean[b * 3] = sdx / (float)J; diff_mean[b * 3 + 1] = sdy / (float)J; diff_mean[b * 3 + 2] = sdz / (float)J; }   // :50
}

// One CTA per joint: sum / count of the visible errors and, for thresholds t_k = k * thr_max / (nthr - 1) (np.linspace, :82), the number
// of errors <= t_k (eval_tool.py:61-68, compared in float64 like numpy does for a float32 array against a float64 scalar).
__global__ void __launch_bounds__(256) eval_measures_kernel(const float* __restrict__ dist, long long N, int J, int nthr, float thr_max,
                                                            double* __restrict__ sum, unsigned* __restrict__ count, unsigned* __restrict__ pck) {
  pdl_entry();
  extern __shared__ unsigned hist[];            // [nthr + 1]: bin k = first threshold index with d <= t_k (nthr: above every threshold)
  __shared__ double red[8];
  __shared__ unsigned cnt_s;
  const int j = blockIdx.x;
  for (int k = threadIdx.x; k <= nthr; k += blockDim.x) hist[k] = 0u;
  if (threadIdx.x == 0) cnt_s = 0u;
  __syncthreads();
  const double step = (double)thr_max / (double)(nthr - 1);
  double s = 0.0;
  unsigned n = 0;
  for (long long r = threadIdx.x; r < N; r += blockDim.x) {
    const float df = dist[r * J + j];
    if (df < 0.f) continue;
    const double d = (double)df;
    s += d; ++n;
    int k = (int)ceil(d / step);
    if (k < 0) k = 0;
    if (k > nthr) k = nthr;
    // exact fix-up against the same threshold values numpy compares with (the last one is thr_max itself)
    auto thr = [&](int q) { return (q == nthr - 1) ? (double)thr_max : (double)q * step; };
    while (k > 0 && d <= thr(k - 1)) --k;
    while (k < nthr && d > thr(k)) ++k;
    atomicAdd(&hist[k], 1u);
  }
  // block reduce
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); n += __shfl_xor_sync(0xffffffffu, n, o); }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s; atomicAdd(&cnt_s, n); }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    sum[j] = t; count[j] = cnt_s;
    unsigned run = 0u;
    for (int k = 0; k < nthr; ++k) { run += hist[k]; pck[(size_t)j * nthr + k] = run; }
  }
}

}  // namespace

extern "C" {

int awr_eval_feed(const float* uvd_pred, const float* xyz_gt_norm, const float* center_xyz, const float* M, const float* cube,
                  const unsigned char* vis, int B, int J, float img_size, float fx, float fy, float fu, float fv, float flip, float* uvd_img,
                  float* dist, float* diff_mean, void* stream) {
  AWR_HOST_CHECK(uvd_pred && xyz_gt_norm && center_xyz && M && cube && uvd_img && dist && B > 0 && J > 0);
  const int warps_per_block = 4;
  launch_pdl(eval_feed_kernel, dim3((B + warps_per_block - 1) / warps_per_block), dim3(32 * warps_per_block), 0, (cudaStream_t)stream, uvd_pred,
             xyz_gt_norm, center_xyz, M, cube, vis, B, J, img_size, fx, fy, fu, fv, flip, uvd_img, dist, diff_mean);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

int awr_eval_measures(const float* dist, long long N, int J, int nthr, float thr_max, double* sum, unsigned* count, unsigned* pck_count,
                      void* stream) {
  AWR_HOST_CHECK(dist && sum && count && pck_count && N > 0 && J > 0 && nthr >= 2 && nthr <= 4096 && thr_max > 0.f);
  launch_pdl(eval_measures_kernel, dim3(J), dim3(256), (size_t)(nthr + 1) * sizeof(unsigned), (cudaStream_t)stream, dist, N, J, nthr, thr_max, sum,
             count, pck_count);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

}  // extern "C"
