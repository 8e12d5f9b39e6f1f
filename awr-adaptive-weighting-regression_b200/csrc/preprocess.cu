// Depth preprocessing of the data path on the device, one launch per batch of raw frames:
//   Loader.crop (bounds2crop + cv2.resize INTER_NEAREST + centring pad, dataloader/loader.py:19-51,190-207), optionally the NYU wire-format
//   decode of nyu_loader.py:71-74 (16-bit depth in the B/G bytes of a PNG),
//   Loader.augment's image half (loader.py:74-86): translate / scale = Loader.recrop (:125-139: cv2.warpPerspective INTER_LINEAR, drop pixels
//   below min(depth > 0) - 1, cube clamp), rotate (:141-161: cv2.warpAffine INTER_LINEAR),
//   Loader.normalize (:88-101).
// The O(1)-per-frame geometry (center2bounds / center2transmat / the warp matrices and their inverses, float64) stays on the host
// (preprocess.py) and arrives as `params`.
//
// One frame = one thread-block cluster of 8 CTAs.  Each CTA gathers 1/8 of the crop into its shared memory (the crop never goes to HBM),
// the cluster reduces max / min-positive depth through distributed shared memory, and the warp reads its four bilinear taps from whichever
// CTA of the cluster holds them (DSMEM).  cv2's arithmetic is kept operation by operation: coordinates in float64 with round-half-even to
// 1/32 pixel (10-bit fixed point for the affine map), float32 table weights, the four products summed left to right without FMA
// contraction -- the result is bit-identical to the reference running the real cv2 (tests/golden/augment_cases.npz).
#include <cooperative_groups.h>

#include "common.cuh"
#include "awr_b200.h"

namespace cg = cooperative_groups;

namespace {

constexpr int kFrameCtas = 8;        // cluster size: CTAs per frame
constexpr int kPreThreads = 256;

// params[n][pstride] (double).  0 ustart, 1 vstart, 2 w, 3 h (crop box in source pixels), 4 size_w, 5 size_h (box after the nearest resize),
// 6 x0, 7 y0 (paste offset in the output), 8 zstart, 9 zend (cube front / back of the crop, mm), 10 centre z, 11 cube_z / 2 (normalize);
// pstride 32 adds: 12 op (0 none, 1 perspective = translate / scale, 2 affine = rotate), 13..21 the INVERSE map (3x3 row-major; affine
// uses 13..18), 22 zstart, 23 zend of the cube after the augmentation (recrop's clamp), 24 tile width of cv2's perspective loop.
struct Slice {
  float* tile;      // this CTA's part of the crop
  int per_cta;      // pixels per CTA
};

__device__ __forceinline__ float crop_tap(cg::cluster_group& cl, const Slice& sl, int D, int yy, int xx) {
  if (yy < 0 || yy >= D || xx < 0 || xx >= D) return 0.f;          // BORDER_CONSTANT, borderValue 0 (loader.py:126,152)
  const int idx = yy * D + xx, r = idx / sl.per_cta;
  return cl.map_shared_rank(sl.tile, r)[idx - r * sl.per_cta];
}

__device__ __forceinline__ int clamp_short(int v) { return max(-32768, min(32767, v)); }

__global__ void __cluster_dims__(kFrameCtas, 1, 1) __launch_bounds__(kPreThreads)
preprocess_kernel(const void* __restrict__ src, int src_format, int Hs, int Ws, const double* __restrict__ params, int pstride, int D,
                  float* __restrict__ out) {
  pdl_entry();
  extern __shared__ float tile[];
  __shared__ float red[2][kPreThreads / 32];
  __shared__ float cta_stat[2];                     // this CTA's max and min-positive depth (read by the whole cluster)
  cg::cluster_group cl = cg::this_cluster();
  const int n = blockIdx.x / kFrameCtas, rank = (int)cl.block_rank();
  const double* p = params + (size_t)n * pstride;
  const int ustart = (int)p[0], vstart = (int)p[1], w = (int)p[2], h = (int)p[3], sw = (int)p[4], sh = (int)p[5], x0 = (int)p[6], y0 = (int)p[7];
  const double zstart = p[8], zend = p[9], cz = p[10], half = p[11];
  const double ifx = 1.0 / ((double)sw / (double)w), ify = 1.0 / ((double)sh / (double)h);       // cv2 resizeNN: sx = min(floor(x * ifx), w - 1)
  const int per_cta = (D * D + kFrameCtas - 1) / kFrameCtas;
  const int base = rank * per_cta, lim = min(D * D, base + per_cta);
  const Slice sl{tile, per_cta};

  // ---- phase 1: this CTA's slice of the crop -> shared memory; max and min-positive depth ------------------------------------------
  float mx = 0.f, mn = __int_as_float(0x7f800000);  // the crop is >= 0 everywhere (padding and invalid pixels are 0)
  for (int i = base + threadIdx.x; i < lim; i += kPreThreads) {
    const int y = i / D, x = i - y * D;
    float d = 0.f;
    if (y >= y0 && y < y0 + sh && x >= x0 && x < x0 + sw) {
      const int sx = min((int)floor((double)(x - x0) * ifx), w - 1), sy = min((int)floor((double)(y - y0) * ify), h - 1);
      const int u = ustart + sx, v = vstart + sy;
      if (u >= 0 && u < Ws && v >= 0 && v < Hs) {                     // outside the image: bounds2crop pads with 0
        const size_t q = ((size_t)n * Hs + v) * Ws + u;
        if (src_format == 0) d = reinterpret_cast<const float*>(src)[q];
        else { const unsigned char* b = reinterpret_cast<const unsigned char*>(src) + q * 3; d = (float)((unsigned)b[0] + 256u * (unsigned)b[1]); }
        if (d != 0.f && (double)d < zstart) d = (float)zstart;          // loader.py:200-205
        else if (d != 0.f && (double)d > zend) d = 0.f;
      }
    }
    tile[i - base] = d;
    mx = fmaxf(mx, d);
    if (d > 0.f) mn = fminf(mn, d);
  }
  mx = warp_max(mx);
  mn = -warp_max(-mn);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = mx; red[1][threadIdx.x >> 5] = mn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < kPreThreads / 32; ++k) { mx = fmaxf(mx, red[0][k]); mn = fminf(mn, red[1][k]); }
    cta_stat[0] = mx; cta_stat[1] = mn;
  }
  cl.sync();                                        // every slice and every CTA's statistics are in place
  mx = 0.f; mn = __int_as_float(0x7f800000);
  for (int r = 0; r < kFrameCtas; ++r) {
    const float* st = cl.map_shared_rank(cta_stat, r);
    mx = fmaxf(mx, st[0]); mn = fminf(mn, st[1]);
  }
  const float depth_max = mx;                       // loader.py:75: the maximum of the crop BEFORE the warp
  const float nv_val = __fsub_rn(mn, 1.f);          // :116,173: np.min(img[img > 0]) - 1 (float32)

  // ---- phase 2: warp (taps over DSMEM) + recrop clean-up + normalize ------------------------------------------------------------------
  const int op = pstride > 12 ? (int)p[12] : 0;
  double m[9];
  double zs2 = 0.0, ze2 = 0.0;
  int bw = D;
  if (op) {
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = p[13 + k];
    zs2 = p[22]; ze2 = p[23]; bw = max(1, (int)p[24]);
  }
  const float bg = (float)(cz + half);
  float* o = out + (size_t)n * D * D;
  for (int i = base + threadIdx.x; i < lim; i += kPreThreads) {
    float v;
    if (op == 0) {
      v = tile[i - base];
    } else {
      const int y = i / D, x = i - y * D;
      int X, Y;
      if (op == 2) {                                // WarpAffineInvoker: AB_SCALE = 1024, x and y terms rounded separately, round_delta = 16
        const int ad = __double2int_rn(__dmul_rn(__dmul_rn(m[0], (double)x), 1024.0)), bd = __double2int_rn(__dmul_rn(__dmul_rn(m[3], (double)x), 1024.0));
        const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)y), m[2]), 1024.0)) + 16;
        const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)y), m[5]), 1024.0)) + 16;
        X = (X0 + ad) >> 5; Y = (Y0 + bd) >> 5;
      } else {                                      // WarpPerspectiveInvoker: row term at the tile's first column, 32 / W, round to 1/32 pixel
        const int bx = x / bw * bw;
        const double xb = (double)bx, x1 = (double)(x - bx), yd = (double)y;
        const double Xr = __dadd_rn(__dadd_rn(__dmul_rn(m[0], xb), __dmul_rn(m[1], yd)), m[2]);
        const double Yr = __dadd_rn(__dadd_rn(__dmul_rn(m[3], xb), __dmul_rn(m[4], yd)), m[5]);
        const double Wr = __dadd_rn(__dadd_rn(__dmul_rn(m[6], xb), __dmul_rn(m[7], yd)), m[8]);
        double Wv = __dadd_rn(Wr, __dmul_rn(m[6], x1));
        Wv = Wv != 0.0 ? __ddiv_rn(32.0, Wv) : 0.0;
        const double fX = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(Xr, __dmul_rn(m[0], x1)), Wv)));
        const double fY = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(Yr, __dmul_rn(m[3], x1)), Wv)));
        X = __double2int_rn(fX); Y = __double2int_rn(fY);
      }
      const int sx = clamp_short(X >> 5), sy = clamp_short(Y >> 5);
      const float fx = __fmul_rn((float)(X & 31), 1.0f / 32), fy = __fmul_rn((float)(Y & 31), 1.0f / 32);
      const float wx0 = __fsub_rn(1.f, fx), wy0 = __fsub_rn(1.f, fy);
      const float t00 = crop_tap(cl, sl, D, sy, sx), t01 = crop_tap(cl, sl, D, sy, sx + 1);
      const float t10 = crop_tap(cl, sl, D, sy + 1, sx), t11 = crop_tap(cl, sl, D, sy + 1, sx + 1);
      v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00, __fmul_rn(wy0, wx0)), __fmul_rn(t01, __fmul_rn(wy0, fx))), __fmul_rn(t10, __fmul_rn(fy, wx0))),
                    __fmul_rn(t11, __fmul_rn(fy, fx)));
      if (op == 1) {                                // Loader.recrop :128-137
        if (v < nv_val) v = 0.f;
        if (v != 0.f && (double)v < zs2) v = (float)zs2;
        else if (v != 0.f && (double)v > ze2) v = 0.f;
      }
    }
    if (v == depth_max) v = bg;                     // loader.py:89 (the crop's maximum becomes background)
    if (v == 0.f) v = bg;                           // :91 invalid points are background
    const double r = fmin(fmax((double)v, cz - half), cz + half);      // :93-95, float64 like numpy with np.float64 bounds
    o[i] = (float)((r - cz) / half);                // :98-99
  }
  cl.sync();                                        // no CTA leaves while a peer may still read its slice
}

int launch_preprocess(const void* src, int src_format, int N, int Hs, int Ws, const double* params, int pstride, int D, float* out, cudaStream_t st) {
  const size_t smem = (size_t)((D * D + kFrameCtas - 1) / kFrameCtas) * sizeof(float);
  if (smem > 200 * 1024) return AWR_ERR_UNSUPPORTED; // img_size <= 640
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    const cudaError_t e = cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  launch_pdl_cluster(preprocess_kernel, dim3(N * kFrameCtas), dim3(kPreThreads), smem, st, kFrameCtas, src, src_format, Hs, Ws, params, pstride, D, out);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

}  // namespace

extern "C" int awr_crop_normalize(const void* src, int src_format, int N, int Hs, int Ws, const double* params, int img_size, float* out,
                                  void* stream) {
  AWR_HOST_CHECK(src && params && out && N > 0 && Hs > 0 && Ws > 0 && img_size > 0 && (src_format == 0 || src_format == 1));
  return launch_preprocess(src, src_format, N, Hs, Ws, params, 12, img_size, out, (cudaStream_t)stream);
}

extern "C" int awr_crop_augment_normalize(const void* src, int src_format, int N, int Hs, int Ws, const double* params, int img_size, float* out,
                                          void* stream) {
  AWR_HOST_CHECK(src && params && out && N > 0 && Hs > 0 && Ws > 0 && img_size > 0 && (src_format == 0 || src_format == 1));
  return launch_preprocess(src, src_format, N, Hs, Ws, params, 32, img_size, out, (cudaStream_t)stream);
}
