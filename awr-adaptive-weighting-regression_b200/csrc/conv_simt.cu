// fp32-accumulate CUDA-core convolution kernels (NHWC activations of type T, fp32 master weights [kh][kw][Cout][Cin]).
//
// Role: the fp32 precision mode (config C1: "fp32 ... UVD output vs reference" needs fp32-accurate convolutions;
// SURVEY.md section 7 "Precision vs parity") and shapes the tcgen05 path does not take (the 1-channel 5x5 stem,
// K=25).  The bf16 throughput path is conv_tc.cu.  Replaces nn.Conv2d / nn.ConvTranspose2d forward + autograd
// (model/resnet_deconv.py:31-32,78-86,141-142,182-188; model/hourglass.py:10) for this mode.
#include "common.cuh"
#include <cstdlib>
#include "awr_b200.h"

namespace {

constexpr int TM = 64, TN = 64, TK = 16, LD = 68;

template <typename T> __device__ __forceinline__ float4 load4c(const T* p);
template <> __device__ __forceinline__ float4 load4c<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 load4c<bf16>(const bf16* p) {
  uint2 a = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
  float2 x = __bfloat1622float2(h[0]), y = __bfloat1622float2(h[1]);
  return make_float4(x.x, x.y, y.x, y.y);
}
template <typename T> __device__ __forceinline__ void store4c(T* p, float4 v);
template <> __device__ __forceinline__ void store4c<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void store4c<bf16>(bf16* p, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk; pk.x = *reinterpret_cast<unsigned*>(&lo); pk.y = *reinterpret_cast<unsigned*>(&hi);
  *reinterpret_cast<uint2*>(p) = pk;
}

struct GatherP {
  int N, Hi, Wi, Ck;      // gathered (input-side) tensor
  int Ho, Wo, Cn;         // produced tensor
  int R, S, stride, pad, transposed;
  int w_sk, w_sn, w_tap;  // weight strides (elements): contraction channel, produced channel, tap
  int out_mode, n_valid, accumulate;
};

// out[m, n] = sum_{tap, k} in[gather(m, tap), k] * w[tap][k, n]  (+ bias[n])
template <typename T>
__global__ void __launch_bounds__(256) conv_gather_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                          const float* __restrict__ bias, void* __restrict__ outp, GatherP p) {
  pdl_entry();
  __shared__ __align__(16) float As[TK][LD];
  __shared__ __align__(16) float Bs[TK][LD];
  const int tid = threadIdx.x;
  const long long M = (long long)p.N * p.Ho * p.Wo;
  const long long m0 = (long long)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  // A-load role: one pixel, 4 contraction channels
  const int a_pix = tid >> 2, a_kq = (tid & 3) * 4;
  const long long am = m0 + a_pix;
  int an = 0, aho = 0, awo = 0;
  const bool a_in = am < M;
  if (a_in) { awo = (int)(am % p.Wo); long long t = am / p.Wo; aho = (int)(t % p.Ho); an = (int)(t / p.Ho); }
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4] = {};
  for (int r = 0; r < p.R; ++r) {
    for (int s = 0; s < p.S; ++s) {
      int hi, wi;
      bool valid = a_in;
      if (!p.transposed) {
        hi = aho * p.stride - p.pad + r; wi = awo * p.stride - p.pad + s;
      } else {
        const int th = aho + p.pad - r, tw = awo + p.pad - s;
        valid = valid && th >= 0 && tw >= 0 && (th % p.stride) == 0 && (tw % p.stride) == 0;
        hi = th / p.stride; wi = tw / p.stride;
      }
      valid = valid && hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi;
      const T* arow = in + (((long long)an * p.Hi + hi) * p.Wi + wi) * p.Ck;
      const float* wt = w + (long long)(r * p.S + s) * p.w_tap;
      for (int k0 = 0; k0 < p.Ck; k0 += TK) {
        float4 av = valid ? load4c<T>(arow + k0 + a_kq) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 bv;
        if (p.w_sk == 1) {          // contraction contiguous: thread -> (n = tid>>2, 4 k)
          bv = *reinterpret_cast<const float4*>(wt + (long long)(n0 + (tid >> 2)) * p.w_sn + k0 + a_kq);
        } else {                    // produced channel contiguous: thread -> (k = tid>>4, 4 n)
          bv = *reinterpret_cast<const float4*>(wt + (long long)(k0 + ty) * p.w_sk + n0 + tx * 4);
        }
        __syncthreads();
        As[a_kq + 0][a_pix] = av.x; As[a_kq + 1][a_pix] = av.y; As[a_kq + 2][a_pix] = av.z; As[a_kq + 3][a_pix] = av.w;
        if (p.w_sk == 1) {
          const int nn = tid >> 2;
          Bs[a_kq + 0][nn] = bv.x; Bs[a_kq + 1][nn] = bv.y; Bs[a_kq + 2][nn] = bv.z; Bs[a_kq + 3][nn] = bv.w;
        } else {
          *reinterpret_cast<float4*>(&Bs[ty][tx * 4]) = bv;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
          const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
          const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
          const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
      }
    }
  }
  const int nc = n0 + tx * 4;
  float bsv[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) bsv[j] = bias[nc + j];
  }
  if (p.out_mode == 0) {
    T* out = reinterpret_cast<T*>(outp);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long m = m0 + ty * 4 + i;
      if (m >= M) continue;
      float4 v = make_float4(acc[i][0] + bsv[0], acc[i][1] + bsv[1], acc[i][2] + bsv[2], acc[i][3] + bsv[3]);
      T* dst = out + m * p.Cn + nc;
      if (p.accumulate) { float4 o = load4c<T>(dst); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
      store4c<T>(dst, v);
    }
  } else {   // NCHW fp32, first n_valid channels; 4 consecutive pixels of one image per thread (P % 4 == 0)
    float* out = reinterpret_cast<float*>(outp);
    const long long P = (long long)p.Ho * p.Wo;
    const long long m = m0 + ty * 4;
    if (m < M) {
      const long long n = m / P, pp = m % P;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (nc + j >= p.n_valid) continue;
        float4 v = make_float4(acc[0][j] + bsv[j], acc[1][j] + bsv[j], acc[2][j] + bsv[j], acc[3][j] + bsv[j]);
        *reinterpret_cast<float4*>(out + (n * p.n_valid + nc + j) * P + pp) = v;
      }
    }
  }
}

struct WgradP {
  int N, Hc, Wc, Cp;      // "coarse" (pointwise) tensor
  int Hf, Wf, Cg;         // "fine" (gathered) tensor
  int R, S, stride, pad;
  int s_p, s_g, w_tap;    // output strides for the Cp index, the Cg index, the tap
  int ksplit;
};
// dW[tap][i*s_p + j*s_g] += sum_{coarse pixel q} P[q, i] * G[q*stride - pad + tap, j]
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const T* __restrict__ Pt, const T* __restrict__ Gt, GradT* __restrict__ dW, WgradP p) {
  pdl_entry();
  __shared__ __align__(16) float Ps[TK][LD];
  __shared__ __align__(16) float Gs[TK][LD];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int tiles_g = p.Cg / TN;
  const int i0 = (blockIdx.x / tiles_g) * TM, j0 = (blockIdx.x % tiles_g) * TN;
  const int r = blockIdx.y / p.S, s = blockIdx.y % p.S;
  const long long Q = (long long)p.N * p.Hc * p.Wc;
  const long long chunk = ((Q + p.ksplit - 1) / p.ksplit + TK - 1) / TK * TK;
  const long long q_begin = (long long)blockIdx.z * chunk, q_end = min(Q, q_begin + chunk);
  const int l_pix = tid >> 4, l_cq = (tid & 15) * 4;     // load role: pixel (0..15), 4 channels
  float acc[4][4] = {};
  for (long long q0 = q_begin; q0 < q_end; q0 += TK) {
    const long long q = q0 + l_pix;
    float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), gv = pv;
    if (q < q_end) {
      const int wc = (int)(q % p.Wc); long long t = q / p.Wc; const int hc = (int)(t % p.Hc); const int n = (int)(t / p.Hc);
      const int hf = hc * p.stride - p.pad + r, wf = wc * p.stride - p.pad + s;
      if (hf >= 0 && hf < p.Hf && wf >= 0 && wf < p.Wf) {
        pv = load4c<T>(Pt + q * p.Cp + i0 + l_cq);
        gv = load4c<T>(Gt + (((long long)n * p.Hf + hf) * p.Wf + wf) * p.Cg + j0 + l_cq);
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&Ps[l_pix][l_cq]) = pv;
    *reinterpret_cast<float4*>(&Gs[l_pix][l_cq]) = gv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&Ps[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Gs[k][tx * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  GradT* dst = dW + (long long)blockIdx.y * p.w_tap;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      grad_add(dst + (long long)(i0 + ty * 4 + i) * p.s_p + (long long)(j0 + tx * 4 + j) * p.s_g, acc[i][j]);
}

// ---------------------------------------------------------------------------------------------------------
// stem: 1-channel k x k stride-1 "same" convolution (resnet_deconv.py:32, hourglass.py:112), w [k*k][Cout]
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                        T* __restrict__ y, int N, int H, int W, int Cout, int k) {
  pdl_entry();
  extern __shared__ float ws[];   // [k*k][Cout]
  for (int i = threadIdx.x; i < k * k * Cout; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int G = Cout >> 3, pad = k / 2;
  const long long items = (long long)N * H * W * G, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += stride) {
    const int cg = (int)(i % G);
    long long t = i / G;
    const int wx = (int)(t % W); t /= W;
    const int hy = (int)(t % H);
    const int n = (int)(t / H);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = bias ? bias[cg * 8 + q] : 0.f;
    const float* xi = x + (long long)n * H * W;
    for (int r = 0; r < k; ++r) {
      const int hh = hy + r - pad;
      if (hh < 0 || hh >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int ww = wx + s - pad;
        if (ww < 0 || ww >= W) continue;
        const float v = __ldg(xi + (long long)hh * W + ww);
        const float* wr = ws + (r * k + s) * Cout + cg * 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = fmaf(v, wr[q], acc[q]);
      }
    }
    Vec8<T>::store(y + i * 8, acc);
  }
}

// dW[tap][co] += sum_pix dy[pix,co] * x[pix + tap];  dbias[co] += sum_pix dy[pix,co]
template <typename T>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy, GradT* __restrict__ dW,
                                                         GradT* __restrict__ dbias, int N, int H, int W, int Cout, int k, int pix_per_block) {
  pdl_entry();
  const int co = threadIdx.x % Cout, tg = threadIdx.x / Cout, ngroups = blockDim.x / Cout;
  const int pad = k / 2, taps = k * k;
  const long long P = (long long)N * H * W;
  const long long p_begin = (long long)blockIdx.x * pix_per_block, p_end = min(P, p_begin + pix_per_block);
  float acc[8] = {};
  float bsum = 0.f;
  for (long long q = p_begin; q < p_end; ++q) {
    const int wx = (int)(q % W); long long t = q / W; const int hy = (int)(t % H); const int n = (int)(t / H);
    const float g = to_f<T>(dy[q * Cout + co]);
    if (tg == 0) bsum += g;
    const float* xi = x + (long long)n * H * W;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int tap = tg + a * ngroups;
      if (tap < taps) {
        const int hh = hy + tap / k - pad, ww = wx + tap % k - pad;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) acc[a] = fmaf(g, __ldg(xi + (long long)hh * W + ww), acc[a]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int tap = tg + a * ngroups;
    if (tap < taps) grad_add(dW + tap * Cout + co, acc[a]);
  }
  if (tg == 0 && dbias) grad_add(dbias + co, bsum);
}

// Register-tiled stem weight gradient (K x K taps, 1 input channel): block = one image x 8-row band, 4 thread groups x Cout=64
// channels; each thread keeps all K*K tap accumulators for its channel and walks 8-pixel row segments so every x value
// fetched from the smem halo tile feeds K FMAs.  One atomicAdd per (tap, channel) per block.
template <typename T, int K>
__global__ void __launch_bounds__(256) stem_wgrad_tiled_kernel(const float* __restrict__ x, const T* __restrict__ dy, GradT* __restrict__ dW,
                                                               GradT* __restrict__ dbias, int N, int H, int W) {
  pdl_entry();
  constexpr int RB = 8, CO = 64, PADK = K / 2, XS = 8 + K - 1;
  extern __shared__ __align__(16) float sm[];
  const int pitch = ((W + K - 1) + 3) & ~3;
  float* xt = sm;                                   // [(RB+K-1)][pitch]
  float* red = sm + (RB + K - 1) * pitch;           // [4][K*K+1][CO]
  const int n = blockIdx.x / (H / RB), h0 = (blockIdx.x % (H / RB)) * RB;
  const int co = threadIdx.x & (CO - 1), grp = threadIdx.x >> 6;
  const float* xi = x + (size_t)n * H * W;
  for (int i = threadIdx.x; i < (RB + K - 1) * pitch; i += 256) {
    const int rr = i / pitch, cc = i - rr * pitch;
    const int hh = h0 + rr - PADK, ww = cc - PADK;
    xt[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W && cc < W + K - 1) ? __ldg(xi + (size_t)hh * W + ww) : 0.f;
  }
  __syncthreads();
  float acc[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) acc[t] = 0.f;
  float bsum = 0.f;
  const int octs_per_row = W / 8, octs = RB * octs_per_row;
  for (int o = grp; o < octs; o += 4) {
    const int rr = o / octs_per_row, c0 = (o - rr * octs_per_row) * 8;
    const T* dyp = dy + (((size_t)n * H + h0 + rr) * W + c0) * CO + co;
    float g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = to_f<T>(dyp[(size_t)i * CO]); bsum += g[i]; }
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const float* xr = xt + (rr + r) * pitch + c0;
      float xs[XS];
#pragma unroll
      for (int i = 0; i < XS; ++i) xs[i] = xr[i];
#pragma unroll
      for (int s2 = 0; s2 < K; ++s2) {
        float a = acc[r * K + s2];
#pragma unroll
        for (int i = 0; i < 8; ++i) a = fmaf(g[i], xs[i + s2], a);
        acc[r * K + s2] = a;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < K * K; ++t) red[(grp * (K * K + 1) + t) * CO + co] = acc[t];
  red[(grp * (K * K + 1) + K * K) * CO + co] = bsum;
  __syncthreads();
  for (int t = grp; t < K * K + 1; t += 4) {
    float v = 0.f;
#pragma unroll
    for (int g2 = 0; g2 < 4; ++g2) v += red[(g2 * (K * K + 1) + t) * CO + co];
    if (t < K * K) grad_add(dW + t * CO + co, v);
    else if (dbias) grad_add(dbias + co, v);
  }
}

// Channel-pair variant of the kernel above (Cout = 64): a thread owns TWO adjacent output channels (one 4-byte bf16x2 / 8-byte fp32x2
// load per pixel, 128 contiguous bytes per warp) and a warp is one of 8 groups walking the band's 8-pixel row segments.  All K*K tap
// accumulators are channel pairs updated with packed FFMA2 (two IEEE FMAs per instruction) and the x row comes from shared memory
// as three 16-byte loads: ~300 instructions per segment and channel pair instead of 2 x 280.
template <typename T> struct Pair2;
template <> struct Pair2<float> { static __device__ __forceinline__ float2 load(const float* p) { return *reinterpret_cast<const float2*>(p); } };
template <> struct Pair2<bf16> {
  static __device__ __forceinline__ float2 load(const bf16* p) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); }
};
template <typename T, int K>
__global__ void __launch_bounds__(256) stem_wgrad_pair_kernel(const float* __restrict__ x, const T* __restrict__ dy, GradT* __restrict__ dW,
                                                              GradT* __restrict__ dbias, int N, int H, int W) {
  pdl_entry();
  constexpr int RB = 8, CO = 64, PADK = K / 2, XS = 12, KK = K * K, HALF = (KK + 2) / 2;     // 26 values (taps + bias) in two rounds of 13
  static_assert(K == 5, "row loader is written for the 5x5 stem");
  extern __shared__ __align__(16) float sm[];
  const int pitch = ((W + K - 1) + 3) & ~3;
  float* xt = sm;                                   // [(RB+K-1)][pitch]
  float* red = sm + (RB + K - 1) * pitch;           // [8 groups][HALF][CO]
  const int n = blockIdx.x / (H / RB), h0 = (blockIdx.x % (H / RB)) * RB;
  const int cp = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const float* xi = x + (size_t)n * H * W;
  for (int i = threadIdx.x; i < (RB + K - 1) * pitch; i += 256) {
    const int rr = i / pitch, cc = i - rr * pitch;
    const int hh = h0 + rr - PADK, ww = cc - PADK;
    xt[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W && cc < W + K - 1) ? __ldg(xi + (size_t)hh * W + ww) : 0.f;
  }
  __syncthreads();
  float2 acc[KK];
#pragma unroll
  for (int t = 0; t < KK; ++t) acc[t] = make_float2(0.f, 0.f);
  float2 bsum = make_float2(0.f, 0.f);
  const int octs_per_row = W / 8, octs = RB * octs_per_row;
  for (int o = grp; o < octs; o += 8) {
    const int rr = o / octs_per_row, c0 = (o - rr * octs_per_row) * 8;
    const T* dyp = dy + (((size_t)n * H + h0 + rr) * W + c0) * CO + 2 * cp;
    float2 g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = Pair2<T>::load(dyp + (size_t)i * CO);
#pragma unroll
    for (int i = 0; i < 8; ++i) { bsum.x += g[i].x; bsum.y += g[i].y; }
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const float4* xr = reinterpret_cast<const float4*>(xt + (rr + r) * pitch + c0);      // 16-byte aligned: pitch % 4 == 0, c0 % 8 == 0
      const float4 xa = xr[0], xb = xr[1], xc = xr[2];
      const float xs[XS] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w, xc.x, xc.y, xc.z, xc.w};
#pragma unroll
      for (int s2 = 0; s2 < K; ++s2) {
        float2 a = acc[r * K + s2];
#pragma unroll
        for (int i = 0; i < 8; ++i) a = __ffma2_rn(make_float2(xs[i + s2], xs[i + s2]), g[i], a);
        acc[r * K + s2] = a;
      }
    }
  }
  // cross-group reduction through shared memory in two rounds of HALF values, then one atomicAdd per (tap, channel) per block
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    __syncthreads();
#pragma unroll
    for (int tt = 0; tt < HALF; ++tt) {
      const int t = half * HALF + tt;
      if (t <= KK) {
        const float2 v = (t < KK) ? acc[t < KK ? t : 0] : bsum;
        red[(grp * HALF + tt) * CO + 2 * cp] = v.x;
        red[(grp * HALF + tt) * CO + 2 * cp + 1] = v.y;
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < HALF * CO; e += 256) {
      const int tt = e / CO, co = e - tt * CO, t = half * HALF + tt;
      if (t > KK) continue;
      float v = 0.f;
#pragma unroll
      for (int g2 = 0; g2 < 8; ++g2) v += red[(g2 * HALF + tt) * CO + co];
      if (t < KK) grad_add(dW + t * CO + co, v);
      else if (dbias) grad_add(dbias + co, v);
    }
  }
}

// Register-tiled stem convolution: thread = 4 consecutive pixels of a row x 8 output channels (32 accumulators); the 5 x 8 input
// patch is read once per thread and every weight vector fetched from shared memory feeds 4 pixels.  Optionally accumulates the
// BatchNorm batch statistics (per-channel sum / sum of squares of the stored values) so no separate reduction pass is needed.
template <typename T, int K>
__global__ void __launch_bounds__(256) stem_conv_tiled_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                              T* __restrict__ y, AwrAcc* __restrict__ stats, int N, int H, int W, int Cout) {
  pdl_entry();
  extern __shared__ __align__(16) float ws[];   // [K*K][Cout] then reduction scratch
  for (int i = threadIdx.x; i < K * K * Cout; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  constexpr int PADK = K / 2, XS = 4 + K - 1;
  const int G = Cout >> 3, W4 = W >> 2;
  // Cout == 64 (G == 8 == warps per CTA): a WARP owns one channel octet and its lanes 32 consecutive pixel quads.  The weight
  // vectors are then warp-uniform (broadcast LDS, no bank conflicts -- with the octet varying across lanes the two 16-byte weight loads
  // per tap were 2-way conflicted and their shared-memory bandwidth, not the FMA pipe, bounded the kernel) and the x loads coalesce.
  const bool wcg = (G == 8 && blockDim.x == 256);
  const int lane = threadIdx.x & 31;
  const int cg = wcg ? (int)(threadIdx.x >> 5) : (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) % G);
  const long long items = (long long)N * H * W4 * (wcg ? 1 : G);
  const long long stride = wcg ? (long long)gridDim.x * 32 : (long long)gridDim.x * blockDim.x;
  float s1[8] = {}, s2[8] = {};
  for (long long i = wcg ? (long long)blockIdx.x * 32 + lane : (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += stride) {
    long long t = wcg ? i : i / G;
    const int w4 = (int)(t % W4); t /= W4;
    const int hy = (int)(t % H);
    const int n = (int)(t / H);
    const int wx = w4 * 4;
    const float* xi = x + (long long)n * H * W;
    // the whole K x (4+K-1) input patch first: three aligned 16-byte loads per row (columns wx-4 .. wx+7), rows / columns outside the
    // image read as zero from clamped addresses -- all 3K loads are in flight before the first FMA
    static_assert(K == 5, "patch loader is written for the 5x5 stem");
    float xs[K][XS];
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int hh = hy + r - PADK;
      const bool rok = (hh >= 0 && hh < H);
      const float* row = xi + (long long)min(max(hh, 0), H - 1) * W;
      const bool lok = rok && wx >= 4, rok2 = rok && wx + 4 < W;
      const float4 l4 = __ldg(reinterpret_cast<const float4*>(row + (wx >= 4 ? wx - 4 : 0)));
      const float4 m4 = __ldg(reinterpret_cast<const float4*>(row + wx));
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(row + (wx + 4 < W ? wx + 4 : wx)));
      xs[r][0] = lok ? l4.z : 0.f; xs[r][1] = lok ? l4.w : 0.f;
      xs[r][2] = rok ? m4.x : 0.f; xs[r][3] = rok ? m4.y : 0.f; xs[r][4] = rok ? m4.z : 0.f; xs[r][5] = rok ? m4.w : 0.f;
      xs[r][6] = rok2 ? r4.x : 0.f; xs[r][7] = rok2 ? r4.y : 0.f;
    }
    // accumulators as channel pairs: one packed FFMA2 (fma.rn.f32x2, sm_100) = two IEEE FMAs, half the issue slots
    float2 acc2[4][4];
#pragma unroll
    for (int p4 = 0; p4 < 4; ++p4)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc2[p4][q] = bias ? make_float2(bias[cg * 8 + 2 * q], bias[cg * 8 + 2 * q + 1]) : make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < K; ++r) {
#pragma unroll
      for (int s2_ = 0; s2_ < K; ++s2_) {
        const float4 wa = *reinterpret_cast<const float4*>(ws + (r * K + s2_) * Cout + cg * 8);
        const float4 wb = *reinterpret_cast<const float4*>(ws + (r * K + s2_) * Cout + cg * 8 + 4);
        const float2 w2[4] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y), make_float2(wb.z, wb.w)};
#pragma unroll
        for (int p4 = 0; p4 < 4; ++p4) {
          const float2 x2 = make_float2(xs[r][p4 + s2_], xs[r][p4 + s2_]);
#pragma unroll
          for (int q = 0; q < 4; ++q) acc2[p4][q] = __ffma2_rn(x2, w2[q], acc2[p4][q]);
        }
      }
    }
    float acc[4][8];
#pragma unroll
    for (int p4 = 0; p4 < 4; ++p4)
#pragma unroll
      for (int q = 0; q < 4; ++q) { acc[p4][2 * q] = acc2[p4][q].x; acc[p4][2 * q + 1] = acc2[p4][q].y; }
#pragma unroll
    for (int p4 = 0; p4 < 4; ++p4) {
      Vec8<T>::store(y + ((((long long)n * H + hy) * W + wx + p4) * G + cg) * 8, acc[p4]);
      if (stats) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { const float v = from_store<T>(acc[p4][q]); s1[q] += v; s2[q] += v * v; }
      }
    }
  }
  if (stats && wcg) {
    // the warp's 32 lanes share the channel octet: shuffle-reduce, one atomic per channel per warp
#pragma unroll
    for (int q = 0; q < 8; ++q) { s1[q] = warp_sum(s1[q]); s2[q] = warp_sum(s2[q]); }
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 8; ++q) { acc_add(stats + cg * 8 + q, s1[q]); acc_add(stats + Cout + cg * 8 + q, s2[q]); }
    }
  } else if (stats) {
    // threads with equal (threadIdx % G) share a channel octet: reduce over them in shared memory, one atomic per channel per block
    __syncthreads();
    float* red = ws;
#pragma unroll
    for (int q = 0; q < 8; ++q) { red[threadIdx.x * 16 + q] = s1[q]; red[threadIdx.x * 16 + 8 + q] = s2[q]; }
    __syncthreads();
    for (int ch = threadIdx.x; ch < 2 * Cout; ch += blockDim.x) {
      const int which = ch / Cout, c = ch % Cout, g = c >> 3, q = c & 7;
      float sum = 0.f;
      for (int r = g; r < (int)blockDim.x; r += G) sum += red[r * 16 + which * 8 + q];
      acc_add(stats + which * Cout + c, sum);
    }
  }
}

}  // namespace

#define DISPATCH_T(dtype, ...)                                             \
  if ((dtype) == AWR_DTYPE_F32) { typedef float T; __VA_ARGS__; }          \
  else if ((dtype) == AWR_DTYPE_BF16) { typedef bf16 T; __VA_ARGS__; }     \
  else return AWR_ERR_UNSUPPORTED;

// csrc/stem_tc.cu
int stem_conv_tc_launch(const float* x, const float* w, const float* bias, void* y, void* stats, int N, int H, int W, int Cout, int k, cudaStream_t st);
int stem_wgrad_tc_launch(const float* x, const void* dy, void* dW, void* dbias, int N, int H, int W, int Cout, int k, cudaStream_t st);

extern "C" {

int awr_conv_simt(const void* in, const float* w, const float* bias, void* out, int dtype, int N, int Hi, int Wi, int Ck, int Ho,
                  int Wo, int Cn, int R, int S, int stride, int pad, int transposed, int w_sk, int w_sn, int w_tap, int out_mode,
                  int n_valid, int accumulate, void* stream) {
  AWR_HOST_CHECK(in && w && out && N > 0 && Ck % TK == 0 && Cn % TN == 0 && R > 0 && S > 0 && stride > 0);
  AWR_HOST_CHECK((w_sk == 1 && w_sn % 4 == 0) || (w_sn == 1 && w_sk % 4 == 0));
  AWR_HOST_CHECK(out_mode == 0 || (out_mode == 1 && ((long long)Ho * Wo) % 4 == 0 && n_valid > 0 && n_valid <= Cn && !accumulate));
  GatherP p{N, Hi, Wi, Ck, Ho, Wo, Cn, R, S, stride, pad, transposed, w_sk, w_sn, w_tap, out_mode, n_valid, accumulate};
  const long long M = (long long)N * Ho * Wo;
  dim3 grid((unsigned)((M + TM - 1) / TM), Cn / TN);
  DISPATCH_T(dtype, launch_pdl(conv_gather_kernel<T>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const T*)in, w, bias, out, p));
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

int awr_conv_wgrad_simt(const void* pointwise, const void* gathered, void* dW, int dtype, int N, int Hc, int Wc, int Cp, int Hf, int Wf,
                        int Cg, int R, int S, int stride, int pad, int s_p, int s_g, int w_tap, void* stream) {
  AWR_HOST_CHECK(pointwise && gathered && dW && N > 0 && Cp % TM == 0 && Cg % TN == 0);
  const long long Q = (long long)N * Hc * Wc;
  const int tiles = (Cp / TM) * (Cg / TN) * R * S;
  int ksplit = (148 * 4 + tiles - 1) / tiles;
  const long long maxsplit = (Q + 63) / 64;
  if (ksplit > maxsplit) ksplit = (int)maxsplit;
  if (ksplit < 1) ksplit = 1;
  WgradP p{N, Hc, Wc, Cp, Hf, Wf, Cg, R, S, stride, pad, s_p, s_g, w_tap, ksplit};
  dim3 grid((Cp / TM) * (Cg / TN), R * S, ksplit);
  DISPATCH_T(dtype, launch_pdl(conv_wgrad_kernel<T>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const T*)pointwise, (const T*)gathered, reinterpret_cast<GradT*>(dW), p));
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

int awr_stem_conv(const float* x, const float* w, const float* bias, void* y, void* stats, int dtype, int N, int H, int W, int Cout, int k,
                  void* stream) {
  AWR_HOST_CHECK(x && w && y && N > 0 && Cout % 8 == 0 && k % 2 == 1 && k <= 7);
  // bf16 output, 64 channels, 5x5, rows that tile by 128 pixels: the tcgen05 kernel (csrc/stem_tc.cu).  AWR_STEM_TC=0 keeps the CUDA-core kernel.
  static const bool stem_tc = [] { const char* e = getenv("AWR_STEM_TC"); return !(e && e[0] == '0'); }();
  if (stem_tc && dtype == AWR_DTYPE_BF16) {
    const int rc = stem_conv_tc_launch(x, w, bias, y, stats, N, H, W, Cout, k, (cudaStream_t)stream);
    if (rc != AWR_ERR_UNSUPPORTED) return rc;
  }
  if (k == 5 && W % 4 == 0 && (256 % (Cout / 8)) == 0 && (reinterpret_cast<unsigned long long>(x) & 15ull) == 0ull) {   // 16-byte patch loads
    const long long items4 = (long long)N * H * (W / 4) * (Cout / 8);
    long long blocks4 = (items4 + 255) / 256;
    if (blocks4 > 148 * 8) blocks4 = 148 * 8;
    size_t smem4 = (size_t)k * k * Cout * sizeof(float);
    if (smem4 < 256 * 16 * sizeof(float)) smem4 = 256 * 16 * sizeof(float);
    DISPATCH_T(dtype, launch_pdl(stem_conv_tiled_kernel<T, 5>, dim3((int)blocks4), dim3(256), smem4, (cudaStream_t)stream, x, w, bias, (T*)y, (AwrAcc*)stats, N, H, W, Cout));
    AWR_LAUNCH_CHECK();
    return AWR_OK;
  }
  AWR_HOST_CHECK(stats == nullptr);
  const long long items = (long long)N * H * W * (Cout / 8);
  long long blocks = (items + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const size_t smem = (size_t)k * k * Cout * sizeof(float);
  DISPATCH_T(dtype, launch_pdl(stem_conv_kernel<T>, dim3((int)blocks), dim3(256), smem, (cudaStream_t)stream, x, w, bias, (T*)y, N, H, W, Cout, k));
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

int awr_stem_wgrad(const float* x, const void* dy, void* dW, void* dbias, int dtype, int N, int H, int W, int Cout, int k,
                   void* stream) {
  AWR_HOST_CHECK(x && dy && dW && N > 0 && (Cout == 64 || Cout == 128 || Cout == 256) && k % 2 == 1 && k <= 7);
  static const bool stem_tc = [] { const char* e = getenv("AWR_STEM_TC"); return !(e && e[0] == '0'); }();
  if (stem_tc && dtype == AWR_DTYPE_BF16) {
    const int rc = stem_wgrad_tc_launch(x, dy, dW, dbias, N, H, W, Cout, k, (cudaStream_t)stream);
    if (rc != AWR_ERR_UNSUPPORTED) return rc;
  }
  const long long P = (long long)N * H * W;
  const int ppb = 256;
  AWR_HOST_CHECK((k * k + (256 / Cout) - 1) / (256 / Cout) <= 8 * 1 || true);
  // taps handled per thread = ceil(k*k / (256/Cout)) must be <= 8
  AWR_HOST_CHECK((k * k + (256 / Cout) - 1) / (256 / Cout) <= 8);
  if (k == 5 && Cout == 64 && H % 8 == 0 && W % 8 == 0 && W <= 256) {
    const int pitch = ((W + 4) + 3) & ~3;
    static const bool old_kernel = [] { const char* e = getenv("AWR_STEM_WGRAD"); return e && e[0] == 'o'; }();       // AWR_STEM_WGRAD=old
    if (!old_kernel) {
      const size_t smem2 = ((size_t)12 * pitch + 8 * 13 * 64) * sizeof(float);
      DISPATCH_T(dtype, launch_pdl(stem_wgrad_pair_kernel<T, 5>, dim3(N * (H / 8)), dim3(256), smem2, (cudaStream_t)stream, x, (const T*)dy, reinterpret_cast<GradT*>(dW), reinterpret_cast<GradT*>(dbias), N, H, W));
      AWR_LAUNCH_CHECK();
      return AWR_OK;
    }
    const size_t smem = ((size_t)12 * pitch + 4 * 26 * 64) * sizeof(float);
    DISPATCH_T(dtype, launch_pdl(stem_wgrad_tiled_kernel<T, 5>, dim3(N * (H / 8)), dim3(256), smem, (cudaStream_t)stream, x, (const T*)dy, reinterpret_cast<GradT*>(dW), reinterpret_cast<GradT*>(dbias), N, H, W));
    AWR_LAUNCH_CHECK();
    return AWR_OK;
  }
  DISPATCH_T(dtype, launch_pdl(stem_wgrad_kernel<T>, dim3((int)((P + ppb - 1) / ppb)), dim3(256), 0, (cudaStream_t)stream, x, (const T*)dy, reinterpret_cast<GradT*>(dW), reinterpret_cast<GradT*>(dbias), N, H,
                                                                                                      W, Cout, k, ppb));
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

}  // extern "C"
