// Depth preprocessing of the non-augmented data path on the device: Loader.crop (bounds2crop + cv2.resize
// INTER_NEAREST + centring pad, dataloader/loader.py:19-51,190-207) fused with Loader.normalize (:88-101) and, optionally, the
// NYU wire-format decode of nyu_loader.py:71-74 (16-bit depth in the G/B channels of a PNG).  The O(1)-per-frame box geometry
// (center2bounds / center2transmat, float64) stays on the host (preprocess.py) and arrives as `params`.
#include "common.cuh"
#include "awr_b200.h"

namespace {

// params[n][12] (double): 0 ustart, 1 vstart, 2 w, 3 h (crop box in source pixels), 4 size_w, 5 size_h (box after the nearest resize),
// 6 x0, 7 y0 (paste offset in the output), 8 zstart, 9 zend (cube front / back, mm), 10 centre z, 11 cube_z / 2
__global__ void __launch_bounds__(256) crop_normalize_kernel(const void* __restrict__ src, int src_format, int Hs, int Ws,
                                                             const double* __restrict__ params, int D, float* __restrict__ out) {
  pdl_entry();
  __shared__ float red[32];
  const int n = blockIdx.x;
  const double* p = params + (size_t)n * 12;
  const int ustart = (int)p[0], vstart = (int)p[1], w = (int)p[2], h = (int)p[3], sw = (int)p[4], sh = (int)p[5], x0 = (int)p[6], y0 = (int)p[7];
  const double zstart = p[8], zend = p[9], cz = p[10], half = p[11];
  const double ifx = 1.0 / ((double)sw / (double)w), ify = 1.0 / ((double)sh / (double)h);       // cv2 resizeNN: sx = min(floor(x * ifx), w - 1)
  float* o = out + (size_t)n * D * D;
  float mx = 0.f;                                   // the crop is >= 0 everywhere (padding and invalid pixels are 0)
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
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
    o[i] = d;
    mx = fmaxf(mx, d);
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();                                  // also orders this CTA's writes of o[] before the reads below
  mx = red[0];
  for (int k = 1; k < (int)(blockDim.x >> 5); ++k) mx = fmaxf(mx, red[k]);
  const float bg = (float)(cz + half);
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    float v = o[i];
    if (v == mx) v = bg;                            // loader.py:89 (the crop's maximum becomes background)
    if (v == 0.f) v = bg;                           // :91 invalid points are background
    double r = fmin(fmax((double)v, cz - half), cz + half);            // :93-95, float64 like numpy with np.float64 bounds
    o[i] = (float)((r - cz) / half);                // :98-99
  }
}

}  // namespace

extern "C" int awr_crop_normalize(const void* src, int src_format, int N, int Hs, int Ws, const double* params, int img_size, float* out,
                                  void* stream) {
  AWR_HOST_CHECK(src && params && out && N > 0 && Hs > 0 && Ws > 0 && img_size > 0 && (src_format == 0 || src_format == 1));
  launch_pdl(crop_normalize_kernel, dim3(N), dim3(256), 0, (cudaStream_t)stream, src, src_format, Hs, Ws, params, img_size, out);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}
