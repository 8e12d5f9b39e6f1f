// Fused adaptive-weighting head + SmoothL1 kernels (bandwidth-bound; one kernel per direction).
//
// Replaces, for the reference path util/feature_tool.py:12-65 + model/loss.py:8-25:
//   coord-grid construction (feature_tool.py:23-27,50-55), nearest depth resample (:20,:44),
//   joint2offset (:12-39), offset2joint_softmax (:41-65), My_SmoothL1Loss (loss.py:8-25).
// Nothing is materialised: the (u,v,d) grid and the dense GT volume are recomputed per pixel in
// registers; each (b,j) CTA streams its 4 planes of the prediction volume exactly once.
#include "common.cuh"
#include "awr_b200.h"
#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace {

constexpr float kSoftmaxScale = 30.0f;     // feature_tool.py:60
constexpr float kDepthBg = 0.99f;          // feature_tool.py:35,57
constexpr int kHeadThreads = 128;
// A (frame, joint) pair is streamed by S CTAs (S = 1 by default; 2 or 4 on request: forward = one thread-block cluster whose
// online-softmax states are combined through distributed shared memory, backward = independent CTAs).  See head_split().

struct Px4 { float v[4]; };

template <typename T> __device__ __forceinline__ Px4 load4(const T* p);
template <> __device__ __forceinline__ Px4 load4<float>(const float* p) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  Px4 r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; return r;
}
template <> __device__ __forceinline__ Px4 load4<bf16>(const bf16* p) {
  uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
  float2 x = __bfloat1622float2(h[0]), y = __bfloat1622float2(h[1]);
  Px4 r; r.v[0] = x.x; r.v[1] = x.y; r.v[2] = y.x; r.v[3] = y.y; return r;
}

// depth of 4 consecutive feature pixels (row r, cols c..c+3) = img[b, r*step, (c+i)*step]   (nearest resample)
__device__ __forceinline__ Px4 load_depth4(const float* img_b, int H, int step, int r, int c) {
  const float* p = img_b + (size_t)(r * step) * H + (size_t)c * step;
  Px4 d;
  if (step == 1) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    d.v[0] = a.x; d.v[1] = a.y; d.v[2] = a.z; d.v[3] = a.w;
  } else if (step == 2) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    d.v[0] = a.x; d.v[1] = a.z; d.v[2] = b.x; d.v[3] = b.z;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) d.v[i] = __ldg(p + i * step);
  }
  return d;
}

__device__ __forceinline__ float coord_of(int i, float Ff) { return (2.0f * ((float)i + 0.5f)) / Ff - 1.0f; }
// the F pixel-centre coordinates, evaluated once per CTA with the reference's exact expression (feature_tool.py:23-24) and then
// looked up, so the streaming loops carry no divisions
__device__ __forceinline__ void fill_axis(float* ax, int F) {
  for (int i = threadIdx.x; i < F; i += blockDim.x) ax[i] = coord_of(i, (float)F);
  __syncthreads();
}

// GT volume of joint2offset for one pixel (feature_tool.py:29-38):  off/dis, (ks-dis)/ks, mask.  One MUFU (rsqrt) instead of a
// square root and four IEEE divisions -- the fused kernels are instruction-bound, not HBM-bound, with the exact forms
// (measured 1.9 TB/s at 239 MB).  Differences to the reference's op order are <= 2 ulp of the volume's values.
__device__ __forceinline__ void gt_pixel(float ju, float jv, float jd, float u, float v, float d, float ks, float inv_ks,
                                         float& g0, float& g1, float& g2, float& gh) {
  const float ox = ju - u, oy = jv - v, oz = jd - d;
  const float d2 = ox * ox + oy * oy + oz * oz + 1e-8f;
  const float inv = rsqrtf(d2);
  const float dis = d2 * inv;
  const float hm = (ks - dis) * inv_ks;
  const float mk = (hm >= 0.f && d < kDepthBg) ? inv : 0.f;          // mask folded into the normalisation factor
  g0 = ox * mk; g1 = oy * mk; g2 = oz * mk; gh = (mk != 0.f) ? hm : 0.f;
}

// deterministic "last CTA reduces the partials" epilogue: partial[nblk][2] -> out[2]
__device__ void finalize_partials(const float* partial, int nblk, float inv0, float inv1, unsigned* counter,
                                  float* out, float* smem) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(counter, 1u);
    is_last = (t == (unsigned)nblk - 1u);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float acc[2] = {0.f, 0.f};
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
    acc[0] += __ldcg(partial + 2 * i);
    acc[1] += __ldcg(partial + 2 * i + 1);
  }
  block_sum<2>(acc, smem);
  if (threadIdx.x == 0) {
    out[0] = acc[0] * inv0;
    out[1] = acc[1] * inv1;
    *counter = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// forward: uvd[b,j,:] (+ softmax stats for backward) (+ joint & dense SmoothL1 sums when GT given)
// ------------------------------------------------------------------------------------------------
template <typename T, int S>
__global__ void __launch_bounds__(kHeadThreads, 4)
head_fwd_kernel(const T* __restrict__ pred, const float* __restrict__ img, const float* __restrict__ jt_gt,
                float* __restrict__ uvd_out, float* __restrict__ stats, float* __restrict__ partial,
                unsigned* __restrict__ counter, float* __restrict__ loss_out, int B, int J, int F, int H, float ks) {
  pdl_entry();
  __shared__ float red[6 * 32];
  __shared__ __align__(16) float ax[256];
  __shared__ float xch[8];               // this CTA's online-softmax state, read by cluster rank 0 when S > 1
  fill_axis(ax, F);
  const int bj = blockIdx.x / S, rank = blockIdx.x - bj * S, b = bj / J, j = bj - b * J;
  const int P = F * F, step = H / F;
  const T* p0 = pred + ((size_t)b * 4 * J + 3 * j) * P;
  const T* ph = pred + ((size_t)b * 4 * J + 3 * J + j) * P;
  const float* img_b = img + (size_t)b * H * H;
  const bool has_gt = (jt_gt != nullptr);
  float ju = 0.f, jv = 0.f, jd = 0.f;
  if (has_gt) { ju = jt_gt[bj * 3]; jv = jt_gt[bj * 3 + 1]; jd = jt_gt[bj * 3 + 2]; }

  float mx = -INFINITY, s = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, hub = 0.f;
  const float inv_ks = 1.0f / ks;
  const int ngroups = P >> 2, gpr = F >> 2;
  // one 4-pixel group: online-softmax update + weighted sums (+ dense SmoothL1 against the on-the-fly GT volume)
  auto consume = [&](const Px4& x0, const Px4& x1, const Px4& x2, const Px4& xh, const Px4& d, int r, int c) {
    const float v = ax[r];
    const float4 u4 = *reinterpret_cast<const float4*>(ax + c);
    const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
    float l[4], gmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float m = (d.v[i] < kDepthBg) ? 1.f : 0.f;
      l[i] = kSoftmaxScale * (xh.v[i] * m);
      gmax = fmaxf(gmax, l[i]);
    }
    if (gmax > mx) {
      float sc = __expf(mx - gmax);
      s *= sc; a0 *= sc; a1 *= sc; a2 *= sc; mx = gmax;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float m = (d.v[i] < kDepthBg) ? 1.f : 0.f;
      float u = uu[i];
      float h = xh.v[i] * m;
      float dis = ks - h * ks;
      float e = __expf(l[i] - mx);
      s += e;
      a0 += e * (x0.v[i] * m * dis + u);
      a1 += e * (x1.v[i] * m * dis + v);
      a2 += e * (x2.v[i] * m * dis + d.v[i]);
      if (has_gt) {
        float g0, g1, g2, gh;
        gt_pixel(ju, jv, jd, u, v, d.v[i], ks, inv_ks, g0, g1, g2, gh);
        hub += huber_val(x0.v[i] - g0) + huber_val(x1.v[i] - g1) + huber_val(x2.v[i] - g2) + huber_val(xh.v[i] - gh);
      }
    }
  };
  // two groups per iteration: all 10 loads of both groups are issued before either is consumed (bytes in flight hide HBM latency)
  const int g_hi = (ngroups / S) * (rank + 1);
  int g = (ngroups / S) * rank + threadIdx.x;
  for (; g + kHeadThreads < g_hi; g += 2 * kHeadThreads) {
    const int gb = g + kHeadThreads;
    const int ra = g / gpr, ca = (g - ra * gpr) << 2, rb = gb / gpr, cb = (gb - rb * gpr) << 2;
    const int oa = g << 2, ob = gb << 2;
    Px4 x0a = load4<T>(p0 + oa), x1a = load4<T>(p0 + P + oa), x2a = load4<T>(p0 + 2 * P + oa), xha = load4<T>(ph + oa);
    Px4 x0b = load4<T>(p0 + ob), x1b = load4<T>(p0 + P + ob), x2b = load4<T>(p0 + 2 * P + ob), xhb = load4<T>(ph + ob);
    Px4 da = load_depth4(img_b, H, step, ra, ca), db = load_depth4(img_b, H, step, rb, cb);
    consume(x0a, x1a, x2a, xha, da, ra, ca);
    consume(x0b, x1b, x2b, xhb, db, rb, cb);
  }
  for (; g < g_hi; g += kHeadThreads) {
    const int r = g / gpr, c = (g - r * gpr) << 2, off = g << 2;
    Px4 x0 = load4<T>(p0 + off), x1 = load4<T>(p0 + P + off), x2 = load4<T>(p0 + 2 * P + off), xh = load4<T>(ph + off);
    Px4 d = load_depth4(img_b, H, step, r, c);
    consume(x0, x1, x2, xh, d, r, c);
  }
  // block combine of the online-softmax states
  float wm = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = wm;
  __syncthreads();
  float bm = red[0];
#pragma unroll
  for (int i = 1; i < kHeadThreads / 32; ++i) bm = fmaxf(bm, red[i]);
  float sc = __expf(mx - bm);            // threads without pixels: exp(-inf) = 0
  float acc[5] = {s * sc, a0 * sc, a1 * sc, a2 * sc, hub};
  block_sum<5>(acc, red);
  if (S > 1) {
    // cluster combine: every CTA publishes (max, sum, 3 weighted sums, huber sum); rank 0 merges them in rank order (deterministic)
    cg::cluster_group cluster = cg::this_cluster();
    if (threadIdx.x == 0) { xch[0] = bm; xch[1] = acc[0]; xch[2] = acc[1]; xch[3] = acc[2]; xch[4] = acc[3]; xch[5] = acc[4]; }
    cluster.sync();
    if (rank == 0 && threadIdx.x == 0) {
      float st[S][6];
      float M = -INFINITY;
#pragma unroll
      for (int r = 0; r < S; ++r) {
        const float* rx = cluster.map_shared_rank(xch, r);
#pragma unroll
        for (int k = 0; k < 6; ++k) st[r][k] = rx[k];
        M = fmaxf(M, st[r][0]);
      }
      bm = M;
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[k] = 0.f;
#pragma unroll
      for (int r = 0; r < S; ++r) {
        const float w = __expf(st[r][0] - M);
        acc[0] += st[r][1] * w; acc[1] += st[r][2] * w; acc[2] += st[r][3] * w; acc[3] += st[r][4] * w; acc[4] += st[r][5];
      }
    }
    cluster.sync();                      // remote reads of xch are complete before any CTA of the cluster may exit
    if (rank != 0) return;
  }
  if (threadIdx.x == 0) {
    float inv = 1.0f / acc[0];
    float o0 = acc[1] * inv, o1 = acc[2] * inv, o2 = acc[3] * inv;
    uvd_out[bj * 3] = o0; uvd_out[bj * 3 + 1] = o1; uvd_out[bj * 3 + 2] = o2;
    stats[bj * 2] = bm; stats[bj * 2 + 1] = acc[0];
    if (has_gt) {
      partial[bj * 2] = huber_val(o0 - ju) + huber_val(o1 - jv) + huber_val(o2 - jd);
      partial[bj * 2 + 1] = acc[4];
    }
  }
  if (has_gt && loss_out != nullptr)
    finalize_partials(partial, B * J, 1.0f / (float)(B * J * 3), 1.0f / ((float)(B * J * 4) * (float)P), counter, loss_out, red);
}

// ------------------------------------------------------------------------------------------------
// backward: dpred = d(head)/dpred . g_uvd  [+ cw * dHuber(uvd,jt)]  [+ dw * dHuber(pred, gt_volume)]
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kHeadThreads, 4)
head_bwd_kernel(const T* __restrict__ pred, const float* __restrict__ img, const float* __restrict__ jt_gt,
                const float* __restrict__ uvd, const float* __restrict__ stats, const float* __restrict__ g_uvd,
                const float* __restrict__ loss_grad, float* __restrict__ dpred, int B, int J, int F, int H, float ks,
                float cw, float dw, int split) {
  pdl_entry();
  __shared__ __align__(16) float ax[256];
  fill_axis(ax, F);
  const int bj = blockIdx.x / split, rank = blockIdx.x - bj * split, b = bj / J, j = bj - b * J;
  const int P = F * F, step = H / F;
  const size_t o0 = ((size_t)b * 4 * J + 3 * j) * P, oh = ((size_t)b * 4 * J + 3 * J + j) * P;
  const float* img_b = img + (size_t)b * H * H;
  const bool has_gt = (jt_gt != nullptr);
  const float lg = loss_grad ? __ldg(loss_grad) : 1.0f;
  const float mx = stats[bj * 2], inv_s = 1.0f / stats[bj * 2 + 1];
  const float q0 = uvd[bj * 3], q1 = uvd[bj * 3 + 1], q2 = uvd[bj * 3 + 2];
  float ju = 0.f, jv = 0.f, jd = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
  if (g_uvd) { g0 = g_uvd[bj * 3]; g1 = g_uvd[bj * 3 + 1]; g2 = g_uvd[bj * 3 + 2]; }
  if (has_gt) {
    ju = jt_gt[bj * 3]; jv = jt_gt[bj * 3 + 1]; jd = jt_gt[bj * 3 + 2];
    const float k = cw * lg / (float)(B * J * 3);
    g0 += k * huber_grad(q0 - ju); g1 += k * huber_grad(q1 - jv); g2 += k * huber_grad(q2 - jd);
  }
  const float kd = (has_gt ? dw * lg : 0.f) / ((float)(B * J * 4) * (float)P);
  const float inv_ks = 1.0f / ks;
  const int ngroups = P >> 2, gpr = F >> 2;
  auto emit = [&](const Px4& x0, const Px4& x1, const Px4& x2, const Px4& xh, const Px4& d, int r, int c, int off) {
    const float v = ax[r];
    const float4 u4 = *reinterpret_cast<const float4*>(ax + c);
    const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
    float r0[4], r1[4], r2[4], rh[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float m = (d.v[i] < kDepthBg) ? 1.f : 0.f;
      float u = uu[i];
      float h = xh.v[i] * m;
      float dis = ks - h * ks;
      float w = __expf(kSoftmaxScale * h - mx) * inv_s;
      float v0 = x0.v[i] * m, v1 = x1.v[i] * m, v2 = x2.v[i] * m;
      float wd = w * dis * m;
      r0[i] = g0 * wd; r1[i] = g1 * wd; r2[i] = g2 * wd;
      float t = g0 * (-ks * v0 + kSoftmaxScale * (v0 * dis + u - q0)) + g1 * (-ks * v1 + kSoftmaxScale * (v1 * dis + v - q1)) +
                g2 * (-ks * v2 + kSoftmaxScale * (v2 * dis + d.v[i] - q2));
      rh[i] = m * w * t;
      if (has_gt) {
        float t0, t1, t2, th;
        gt_pixel(ju, jv, jd, u, v, d.v[i], ks, inv_ks, t0, t1, t2, th);
        r0[i] += kd * huber_grad(x0.v[i] - t0); r1[i] += kd * huber_grad(x1.v[i] - t1);
        r2[i] += kd * huber_grad(x2.v[i] - t2); rh[i] += kd * huber_grad(xh.v[i] - th);
      }
    }
    __stcs(reinterpret_cast<float4*>(dpred + o0 + off), make_float4(r0[0], r0[1], r0[2], r0[3]));
    __stcs(reinterpret_cast<float4*>(dpred + o0 + P + off), make_float4(r1[0], r1[1], r1[2], r1[3]));
    __stcs(reinterpret_cast<float4*>(dpred + o0 + 2 * P + off), make_float4(r2[0], r2[1], r2[2], r2[3]));
    __stcs(reinterpret_cast<float4*>(dpred + oh + off), make_float4(rh[0], rh[1], rh[2], rh[3]));
  };
  const int g_hi = (ngroups / split) * (rank + 1);
  int g = (ngroups / split) * rank + threadIdx.x;
  for (; g + kHeadThreads < g_hi; g += 2 * kHeadThreads) {
    const int gb = g + kHeadThreads;
    const int ra = g / gpr, ca = (g - ra * gpr) << 2, rb = gb / gpr, cb = (gb - rb * gpr) << 2;
    const int oa = g << 2, ob = gb << 2;
    Px4 x0a = load4<T>(pred + o0 + oa), x1a = load4<T>(pred + o0 + P + oa), x2a = load4<T>(pred + o0 + 2 * P + oa), xha = load4<T>(pred + oh + oa);
    Px4 x0b = load4<T>(pred + o0 + ob), x1b = load4<T>(pred + o0 + P + ob), x2b = load4<T>(pred + o0 + 2 * P + ob), xhb = load4<T>(pred + oh + ob);
    Px4 da = load_depth4(img_b, H, step, ra, ca), db = load_depth4(img_b, H, step, rb, cb);
    emit(x0a, x1a, x2a, xha, da, ra, ca, oa);
    emit(x0b, x1b, x2b, xhb, db, rb, cb, ob);
  }
  for (; g < g_hi; g += kHeadThreads) {
    const int r = g / gpr, c = (g - r * gpr) << 2, off = g << 2;
    Px4 x0 = load4<T>(pred + o0 + off), x1 = load4<T>(pred + o0 + P + off), x2 = load4<T>(pred + o0 + 2 * P + off), xh = load4<T>(pred + oh + off);
    Px4 d = load_depth4(img_b, H, step, r, c);
    emit(x0, x1, x2, xh, d, r, c, off);
  }
}

// ------------------------------------------------------------------------------------------------
// stand-alone joint2offset (drop-in for FeatureModule.joint2offset; the fused path never calls it)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kHeadThreads)
joint2offset_kernel(const float* __restrict__ jt, const float* __restrict__ img, float* __restrict__ out, int B, int J,
                    int F, int H, float ks) {
  pdl_entry();
  const int bj = blockIdx.x, b = bj / J, j = bj - b * J;
  const int P = F * F, step = H / F;
  const float Ff = (float)F;
  const size_t o0 = ((size_t)b * 4 * J + 3 * j) * P, oh = ((size_t)b * 4 * J + 3 * J + j) * P;
  const float* img_b = img + (size_t)b * H * H;
  const float ju = jt[bj * 3], jv = jt[bj * 3 + 1], jd = jt[bj * 3 + 2];
  const int ngroups = P >> 2, gpr = F >> 2;
  for (int g = threadIdx.x; g < ngroups; g += kHeadThreads) {
    const int r = g / gpr, c = (g - r * gpr) << 2;
    const int off = g << 2;
    Px4 d = load_depth4(img_b, H, step, r, c);
    const float v = coord_of(r, Ff);
    float r0[4], r1[4], r2[4], rh[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gt_pixel(ju, jv, jd, coord_of(c + i, Ff), v, d.v[i], ks, 1.0f / ks, r0[i], r1[i], r2[i], rh[i]);
    *reinterpret_cast<float4*>(out + o0 + off) = make_float4(r0[0], r0[1], r0[2], r0[3]);
    *reinterpret_cast<float4*>(out + o0 + P + off) = make_float4(r1[0], r1[1], r1[2], r1[3]);
    *reinterpret_cast<float4*>(out + o0 + 2 * P + off) = make_float4(r2[0], r2[1], r2[2], r2[3]);
    *reinterpret_cast<float4*>(out + oh + off) = make_float4(rh[0], rh[1], rh[2], rh[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// stand-alone SmoothL1 (drop-in for My_SmoothL1Loss)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
huber_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n, float* __restrict__ partial,
                 unsigned* __restrict__ counter, float* __restrict__ out) {
  pdl_entry();
  __shared__ float red[2 * 32];
  float acc[2] = {0.f, 0.f};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc[0] += huber_val(x[i] - y[i]);
  block_sum<2>(acc, red);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = acc[0]; partial[2 * blockIdx.x + 1] = 0.f; }
  finalize_partials(partial, gridDim.x, 1.0f / (float)n, 0.f, counter, out, red);
}

__global__ void __launch_bounds__(256)
huber_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n, const float* __restrict__ gout,
                 float* __restrict__ dx) {
  pdl_entry();
  const float k = (gout ? __ldg(gout) : 1.0f) / (float)n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dx[i] = k * huber_grad(x[i] - y[i]);
}

// CTAs per (frame, joint) pair.  Measured on B200 (tools/bench_head.py, B=32 J=14 F=64, cold L2): forward 20.5 / 24.6 / 28.7 us and
// backward 21.5 / 22.6 / 22.7 us for S = 1 / 2 / 4 -- the kernels are bound by instruction issue per pixel, not by CTA count, and the
// cluster launch + two cluster barriers cost more than the finer granularity returns.  S = 1 unless AWR_HEAD_SPLIT=2|4 forces a split.
int head_split(int B, int J, int F) {
  static const int forced = [] { const char* e = getenv("AWR_HEAD_SPLIT"); const int v = e ? atoi(e) : 0; return (v == 2 || v == 4) ? v : 1; }();
  (void)B; (void)J;
  return ((F * F / 4) % forced == 0) ? forced : 1;
}

// forward launch: S-CTA clusters + programmatic dependent launch
template <typename T, int S>
cudaError_t launch_head_fwd(cudaStream_t st, const T* pred, const float* img, const float* uvd_gt, float* uvd_out, float* stats, float* partial,
                            unsigned* counter, float* loss_out, int B, int J, int F, int H, float ks) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B * J * S); cfg.blockDim = dim3(kHeadThreads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = S; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = (S > 1) ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, head_fwd_kernel<T, S>, pred, img, uvd_gt, uvd_out, stats, partial, counter, loss_out, B, J, F, H, ks);
}
template <typename T>
cudaError_t launch_head_fwd_s(int S, cudaStream_t st, const T* pred, const float* img, const float* uvd_gt, float* uvd_out, float* stats,
                              float* partial, unsigned* counter, float* loss_out, int B, int J, int F, int H, float ks) {
  if (S == 4) return launch_head_fwd<T, 4>(st, pred, img, uvd_gt, uvd_out, stats, partial, counter, loss_out, B, J, F, H, ks);
  if (S == 2) return launch_head_fwd<T, 2>(st, pred, img, uvd_gt, uvd_out, stats, partial, counter, loss_out, B, J, F, H, ks);
  return launch_head_fwd<T, 1>(st, pred, img, uvd_gt, uvd_out, stats, partial, counter, loss_out, B, J, F, H, ks);
}

bool head_args_ok(int B, int J, int F, int H) { return B > 0 && J > 0 && F >= 4 && F <= 256 && (F % 4) == 0 && H >= F && (H % F) == 0; }

}  // namespace

extern "C" {

int awr_head_fwd(const void* pred, int pred_dtype, const float* img, const float* uvd_gt, float* uvd_out, float* loss_out,
                 float* ws, int B, int J, int F, int H, float kernel_size, void* stream) {
  AWR_HOST_CHECK(pred && img && uvd_out && ws && head_args_ok(B, J, F, H));
  AWR_HOST_CHECK(uvd_gt != nullptr || loss_out == nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  float* stats = ws;
  float* partial = ws + 2 * (size_t)B * J;
  unsigned* counter = reinterpret_cast<unsigned*>(ws + 4 * (size_t)B * J);
  const int S = head_split(B, J, F);
  if (pred_dtype == AWR_DTYPE_F32)
    launch_head_fwd_s<float>(S, st, (const float*)pred, img, uvd_gt, uvd_out, stats, partial, counter, loss_out, B, J, F, H, kernel_size);
  else if (pred_dtype == AWR_DTYPE_BF16)
    launch_head_fwd_s<bf16>(S, st, (const bf16*)pred, img, uvd_gt, uvd_out, stats, partial, counter, loss_out, B, J, F, H, kernel_size);
  else
    return AWR_ERR_UNSUPPORTED;
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

int awr_head_bwd(const void* pred, int pred_dtype, const float* img, const float* uvd_gt, const float* uvd, const float* ws,
                 const float* g_uvd, const float* loss_grad, float* dpred, int B, int J, int F, int H, float kernel_size,
                 float coord_weight, float dense_weight, void* stream) {
  AWR_HOST_CHECK(pred && img && uvd && ws && dpred && head_args_ok(B, J, F, H));
  AWR_HOST_CHECK(uvd_gt != nullptr || g_uvd != nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  const int S = head_split(B, J, F);
  if (pred_dtype == AWR_DTYPE_F32)
    launch_pdl(head_bwd_kernel<float>, dim3(B * J * S), dim3(kHeadThreads), 0, st, (const float*)pred, img, uvd_gt, uvd, ws, g_uvd, loss_grad, dpred,
               B, J, F, H, kernel_size, coord_weight, dense_weight, S);
  else if (pred_dtype == AWR_DTYPE_BF16)
    launch_pdl(head_bwd_kernel<bf16>, dim3(B * J * S), dim3(kHeadThreads), 0, st, (const bf16*)pred, img, uvd_gt, uvd, ws, g_uvd, loss_grad, dpred,
               B, J, F, H, kernel_size, coord_weight, dense_weight, S);
  else
    return AWR_ERR_UNSUPPORTED;
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

int awr_joint2offset(const float* jt_uvd, const float* img, float* out, int B, int J, int F, int H, float kernel_size,
                     void* stream) {
  AWR_HOST_CHECK(jt_uvd && img && out && head_args_ok(B, J, F, H));
  launch_pdl(joint2offset_kernel, dim3(B * J), dim3(kHeadThreads), 0, (cudaStream_t)stream, jt_uvd, img, out, B, J, F, H, kernel_size);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

int awr_huber_fwd(const float* x, const float* y, long long n, float* ws, float* out, void* stream) {
  AWR_HOST_CHECK(x && y && ws && out && n > 0);
  int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (blocks > AWR_HUBER_MAX_BLOCKS) blocks = AWR_HUBER_MAX_BLOCKS;
  unsigned* counter = reinterpret_cast<unsigned*>(ws + 2 * AWR_HUBER_MAX_BLOCKS);
  launch_pdl(huber_fwd_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, y, n, ws, counter, out);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

int awr_huber_bwd(const float* x, const float* y, long long n, const float* grad_out, float* dx, void* stream) {
  AWR_HOST_CHECK(x && y && dx && n > 0);
  int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(huber_bwd_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, y, n, grad_out, dx);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

}  // extern "C"
