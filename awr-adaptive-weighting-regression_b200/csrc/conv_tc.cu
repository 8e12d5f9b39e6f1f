// NHWC bf16 implicit-GEMM convolution on tcgen05 tensor cores (sm_100a): fprop / dgrad of Conv2d and ConvTranspose2d.
//
//   D[128 pixels, Ntile channels] (fp32, TMEM) = sum over taps t, 64-channel blocks kc of  A_t,kc [128 x 64] * B_t,kc [64 x Ntile]
//
// A tile = one TMA box (64 ch, Wt, Ht, Nt) of the NHWC activation, shifted by the tap offset; padding is TMA out-of-bounds
// zero fill; strided gathers (stride-2 convs, ConvTranspose2d dgrad) use the tensor map's element strides.  No im2col buffer.
// B tile = TMA box of the bf16 weight matrix [tap][Cout][Cin] -- K-major when the contraction runs over Cin (fprop), MN-major
// (transposed by the UMMA descriptor, no transposed weight copy) when it runs over Cout (dgrad).
// ConvTranspose2d fprop / stride-2 Conv2d dgrad are decomposed into stride^2 output-parity classes, each a dense
// small-tap convolution whose epilogue stores to the strided output positions (no zero insertion).
//
// Warp roles (192 threads): warp 4 = TMA producer, warp 5 = TMEM allocator + MMA issuer (one elected lane), warps 0..3 =
// epilogue (TMEM -> registers -> bias/accumulate -> global).  smem ring of kStages {A,B} tiles with full/empty mbarriers,
// two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1; persistent CTAs, static tile schedule.
//
// Replaces nn.Conv2d / nn.ConvTranspose2d (+autograd) of model/resnet_deconv.py:31-53,78-86,141-142,182-188 and
// model/hourglass.py:10 in the bf16 precision mode.
#include "tc_common.cuh"
#include "conv_tc_shared.h"
#include "awr_b200.h"

namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int kMaxTaps = kConvMaxTaps;
constexpr int kMaxClasses = kConvMaxClasses;
constexpr int kABytes = 128 * 128;          // 128 pixels x 64 bf16

typedef ConvTapClass TapClass;

struct ConvTcParams {
  int N, Hc, Wc;                 // coarse (tile) grid
  int lgWt, lgHt, Wt, Ht, Nt;    // tile geometry: 128 = Wt*Ht*Nt
  int tiles_w, tiles_h, tiles_n_img, tiles_m, tiles_c;   // tile counts
  int kblocks, Ntile, stages;
  int a_stride;                  // element stride of the gathered tensor per coarse pixel
  int b_mn;                      // B operand MN-major (dgrad)
  int Ho, Wo, Cn, out_s;         // output tensor (NHWC) and its stride per coarse pixel
  int out_mode, n_valid, accumulate;
  int nclasses;
  TapClass cls[kMaxClasses];
};

#ifdef AWR_CONV_PROFILE
// debug-only role timers (cycles, summed per CTA): [0] producer wait-empty, [1] mma wait-full, [2] mma wait-tmem-empty, [3] epilogue
// wait-tmem-full, [4] epilogue busy, [5] kernel total, [6] tiles, [7] k-iterations
__device__ unsigned long long g_conv_prof[148 * 8];
#define PROF_T0() const long long t0__ = clock64()
#define PROF_ADD(slot) atomicAdd(&g_conv_prof[blockIdx.x * 8 + (slot)], (unsigned long long)(clock64() - t0__))
#else
#define PROF_T0()
#define PROF_ADD(slot)
#endif

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const float* __restrict__ bias,
               void* __restrict__ outp, AwrAcc* __restrict__ stats, const __grid_constant__ ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();                 // the next kernel may start its prologue while this grid runs
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef AWR_CONV_PROFILE
  const long long k_t0 = clock64();
#endif
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_bytes = p.Ntile * 128;
  const int stage_bytes = kABytes + b_bytes;
  const int total_tiles = p.nclasses * p.tiles_m * p.tiles_c;
  // per-CTA BatchNorm partial sums behind the pipeline stages: one private [2*Cn] fp32 slice per epilogue warp (a single writer per
  // element, so no shared-memory atomics and a fixed summation order), combined in warp order at the end and flushed through the
  // order-independent global accumulators (AwrAcc) -- the statistics are bit-reproducible
  float* s_stats = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + (size_t)p.stages * stage_bytes);
  if (stats) {
    for (int i = threadIdx.x; i < 4 * 2 * p.Cn; i += kThreads) s_stats[i] = 0.f;
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();                    // prologue done; from here on the previous kernel's outputs are visible

  if (warp == 4) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int ct = tile % p.tiles_c; int r = tile / p.tiles_c;
        const int mt = r % p.tiles_m; const int c = r / p.tiles_m;
        const int tw = mt % p.tiles_w; int r2 = mt / p.tiles_w;
        const int th = r2 % p.tiles_h; const int tn = r2 / p.tiles_h;
        const int w0 = tw * p.Wt, h0 = th * p.Ht, n0 = tn * p.Nt, c0 = ct * p.Ntile;
        const TapClass& tc_ = p.cls[c];
        for (int t = 0; t < tc_.ntaps; ++t) {
          const int ax = w0 * p.a_stride + tc_.ox[t], ay = h0 * p.a_stride + tc_.oy[t], wi = tc_.widx[t];
          for (int kc = 0; kc < p.kblocks; ++kc) {
            { PROF_T0(); mbar_wait(&empty_bar[stage], phase ^ 1u); PROF_ADD(0); }
            mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
            uint8_t* sa = smem_raw + (smem_base - smem_u32(smem_raw)) + (size_t)stage * stage_bytes;
            tma_load_4d(sa, &tmA, &full_bar[stage], kc * 64, ax, ay, n0);
            uint8_t* sb = sa + kABytes;
            if (!p.b_mn) {
              tma_load_3d(sb, &tmB, &full_bar[stage], kc * 64, c0, wi);
            } else {
              for (int j = 0; j < p.Ntile / 64; ++j) tma_load_3d(sb + j * 8192, &tmB, &full_bar[stage], c0 + 64 * j, kc * 64, wi);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ======================================= MMA issuer =======================================
    // whole warp walks the barriers; one elected lane issues.  Descriptors are built once: per k-iteration only the 14-bit
    // start-address field moves (stage offset, +32 B per UMMA_K for K-major / +2048 B for MN-major operands).
    const uint32_t idesc = umma_idesc_bf16(128, p.Ntile, 0, p.b_mn);
    const uint64_t adesc0 = umma_desc_sw128(smem_base, 16, 1024);
    const uint64_t bdesc0 = p.b_mn ? umma_desc_sw128(smem_base + kABytes, 8192, 1024) : umma_desc_sw128(smem_base + kABytes, 16, 1024);
    const uint32_t bstep = p.b_mn ? (2048u >> 4) : (32u >> 4);
    const uint32_t sstep = (uint32_t)stage_bytes >> 4;
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int c = (tile / p.tiles_c) / p.tiles_m;
      const int iters = p.cls[c].ntaps * p.kblocks;
      { PROF_T0(); mbar_wait(&tempty_bar[as], aphase ^ 1u); if (lane == 0) PROF_ADD(2); }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.Ntile);
      for (int it = 0; it < iters; ++it) {
        { PROF_T0(); mbar_wait(&full_bar[stage], phase); if (lane == 0) PROF_ADD(1); }
        tc_loop_fence();
        if (elect_one()) {
          const uint64_t ad = adesc0 + (uint64_t)((uint32_t)stage * sstep);
          const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)stage * sstep);
          umma_bf16(d_tmem, ad, bd, idesc, it ? 1u : 0u);
          umma_bf16(d_tmem, ad + 2, bd + bstep, idesc, 1u);
          umma_bf16(d_tmem, ad + 4, bd + 2 * bstep, idesc, 1u);
          umma_bf16(d_tmem, ad + 6, bd + 3 * bstep, idesc, 1u);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) {
        if (iters > 0) umma_commit(&tfull_bar[as]);
        else mbar_arrive(&tfull_bar[as]);
      }
      __syncwarp();
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else {
    // ======================================= epilogue =======================================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int dw = row & (p.Wt - 1), dh = (row >> p.lgWt) & (p.Ht - 1), dn = row >> (p.lgWt + p.lgHt);
    int as = 0; uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int ct = tile % p.tiles_c; int r = tile / p.tiles_c;
      const int mt = r % p.tiles_m; const int c = r / p.tiles_m;
      const int tw = mt % p.tiles_w; int r2 = mt / p.tiles_w;
      const int th = r2 % p.tiles_h; const int tn = r2 / p.tiles_h;
      const int n = tn * p.Nt + dn, hc = th * p.Ht + dh, wc = tw * p.Wt + dw, c0 = ct * p.Ntile;
      const bool valid = n < p.N && hc < p.Hc && wc < p.Wc;
      const int ho = hc * p.out_s + p.cls[c].py, wo = wc * p.out_s + p.cls[c].px;
      const bool has_acc = p.cls[c].ntaps > 0;
      if (warp == 0 && lane == 0) { PROF_T0(); mbar_wait(&tfull_bar[as], aphase); PROF_ADD(3); }
      else mbar_wait(&tfull_bar[as], aphase);
#ifdef AWR_CONV_PROFILE
      const long long e_t0 = clock64();
#endif
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.Ntile);
      for (int ch = 0; ch < p.Ntile; ch += 32) {
        uint32_t v[32];
        if (has_acc) { tmem_ld32(t_addr + ch, v); tmem_ld_wait(); }
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        if (!valid && !stats) continue;
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if (bias) {
          const float4* b4 = reinterpret_cast<const float4*>(bias + c0 + ch);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = __ldg(b4 + i);
            f[4 * i] += bb.x; f[4 * i + 1] += bb.y; f[4 * i + 2] += bb.z; f[4 * i + 3] += bb.w;
          }
        }
        if (stats) {
          // per-channel sum / sum of squares of the bf16-rounded outputs (what BatchNorm will normalise): 32x32 transpose-
          // reduce over the warp's rows with 31 shuffles per quantity; lane L ends up owning channel c0+ch+L.
          float s1[32], s2[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float r = valid ? __bfloat162float(__float2bfloat16_rn(f[i])) : 0.f;
            s1[i] = r; s2[i] = r * r;
          }
#pragma unroll
          for (int st = 16; st > 0; st >>= 1) {
            const bool up = (lane & st) != 0;
#pragma unroll
            for (int i = 0; i < st; ++i) {
              const float k1 = up ? s1[i + st] : s1[i], d1 = up ? s1[i] : s1[i + st];
              const float k2 = up ? s2[i + st] : s2[i], d2 = up ? s2[i] : s2[i + st];
              s1[i] = k1 + __shfl_xor_sync(0xffffffffu, d1, st);
              s2[i] = k2 + __shfl_xor_sync(0xffffffffu, d2, st);
            }
          }
          float* mine = s_stats + q * 2 * p.Cn;          // lane L owns channel c0+ch+L of this warp's slice
          mine[c0 + ch + lane] += s1[0];
          mine[p.Cn + c0 + ch + lane] += s2[0];
          if (!valid) continue;
        }
        if (p.out_mode == 0) {
          bf16* dst = reinterpret_cast<bf16*>(outp) + (((size_t)n * p.Ho + ho) * p.Wo + wo) * p.Cn + c0 + ch;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = f[g * 8 + i];
            if (p.accumulate) {
              float e[8];
              Vec8<bf16>::load(dst + g * 8, e);
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] += e[i];
            }
            Vec8<bf16>::store(dst + g * 8, o);
          }
        } else {   // NCHW fp32, first n_valid channels (the (B,4J,F,F) prediction volume the AWR head reads)
          float* dst = reinterpret_cast<float*>(outp);
          const size_t P = (size_t)p.Ho * p.Wo, pix = (size_t)ho * p.Wo + wo;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int cc = c0 + ch + i;
            if (cc < p.n_valid) dst[((size_t)n * p.n_valid + cc) * P + pix] = f[i];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
#ifdef AWR_CONV_PROFILE
      if (warp == 0 && lane == 0) {
        atomicAdd(&g_conv_prof[blockIdx.x * 8 + 4], (unsigned long long)(clock64() - e_t0));
        atomicAdd(&g_conv_prof[blockIdx.x * 8 + 6], 1ull);
        atomicAdd(&g_conv_prof[blockIdx.x * 8 + 7], (unsigned long long)(p.cls[c].ntaps * p.kblocks));
      }
#endif
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (stats) {
    for (int i = threadIdx.x; i < 2 * p.Cn; i += kThreads) {
      const float v = ((s_stats[i] + s_stats[2 * p.Cn + i]) + s_stats[4 * p.Cn + i]) + s_stats[6 * p.Cn + i];
      if (v != 0.f) acc_add(stats + i, v);
    }
  }
  if (warp == 5) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
#ifdef AWR_CONV_PROFILE
  if (threadIdx.x == 0) atomicAdd(&g_conv_prof[blockIdx.x * 8 + 5], (unsigned long long)(clock64() - k_t0));
#endif
}

}  // namespace

#ifdef AWR_CONV_PROFILE
extern "C" int awr_debug_conv_profile(unsigned long long* out_host, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out_host, g_conv_prof, sizeof(g_conv_prof));
  if (reset) { static unsigned long long z[148 * 8]; cudaMemcpyToSymbol(g_conv_prof, z, sizeof(z)); }
  return 0;
}
#endif

// bf16 weight matrix [tap][Cout][Cin] as a 3-D tensor map: K-major boxes (64 k, Ntile n) for fprop, MN-major boxes (64 n, 64 k) for dgrad
bool conv_make_weight_map(CUtensorMap* m, const ConvGeom& g, const void* w, int Ntile) {
  const int T = g.R * g.S;
  if (!g.b_mn) {
    const long long dims[3] = {g.Ck, g.Cn, T};
    const long long str[3] = {1, g.w_sn, T > 1 ? g.w_tap : (long long)g.Cn * g.w_sn};
    const int box[3] = {64, Ntile, 1};
    return tc::make_tmap_bf16(m, w, 3, dims, str, box, nullptr);
  }
  const long long dims[3] = {g.Cn, g.Ck, T};
  const long long str[3] = {1, g.w_sk, T > 1 ? g.w_tap : (long long)g.Ck * g.w_sk};
  const int box[3] = {64, 64, 1};
  return tc::make_tmap_bf16(m, w, 3, dims, str, box, nullptr);
}

extern "C" {

int awr_conv_tc(const void* in, const void* w, const float* bias, void* out, void* stats, int N, int Hi, int Wi, int Ck, int Ho, int Wo,
                int Cn, int R, int S, int stride, int pad, int transposed, int w_sk, int w_sn, int w_tap, int out_mode, int n_valid,
                int accumulate, void* stream) {
  AWR_HOST_CHECK(in && w && out && N > 0 && Ck % 64 == 0 && Cn % 64 == 0 && R > 0 && S > 0 && R * S <= kMaxTaps);
  AWR_HOST_CHECK(stride == 1 || stride == 2);
  AWR_HOST_CHECK((w_sk == 1 && w_sn % 8 == 0) || (w_sn == 1 && w_sk % 8 == 0));
  AWR_HOST_CHECK(w_tap % 8 == 0 || R * S == 1);
  AWR_HOST_CHECK(out_mode == 0 || (out_mode == 1 && n_valid > 0 && n_valid <= Cn && !accumulate));
  AWR_HOST_CHECK(stats == nullptr || (out_mode == 0 && !accumulate));
  // geometry shared by both kernels
  ConvGeom g;
  memset(&g, 0, sizeof(g));
  const int cs = (transposed && stride > 1) ? stride : 1;     // output pixels per coarse pixel (parity classes)
  AWR_HOST_CHECK(Ho % cs == 0 && Wo % cs == 0);
  g.N = N; g.Hi = Hi; g.Wi = Wi; g.Ck = Ck; g.Ho = Ho; g.Wo = Wo; g.Cn = Cn;
  g.Hc = Ho / cs; g.Wc = Wo / cs; g.out_s = cs;
  g.a_stride = (!transposed) ? stride : 1;
  g.R = R; g.S = S; g.w_sk = w_sk; g.w_sn = w_sn; g.w_tap = w_tap; g.b_mn = (w_sn == 1 && w_sk != 1) ? 1 : 0;
  g.out_mode = out_mode; g.n_valid = n_valid; g.accumulate = accumulate;
  AWR_HOST_CHECK(is_pow2(g.Wc) && is_pow2(g.Hc) && g.Wc <= 256 && g.Hc <= 256);
  if (!transposed) {            // in = out*stride - pad + tap
    g.nclasses = 1;
    TapClass& c = g.cls[0];
    for (int r = 0; r < R; ++r)
      for (int s = 0; s < S; ++s) { c.oy[c.ntaps] = (short)(r - pad); c.ox[c.ntaps] = (short)(s - pad); c.widx[c.ntaps] = (short)(r * S + s); ++c.ntaps; }
  } else {                      // in = (out + pad - tap) / stride, only where divisible
    g.nclasses = cs * cs;
    for (int py = 0; py < cs; ++py)
      for (int px = 0; px < cs; ++px) {
        TapClass& c = g.cls[py * cs + px];
        c.py = py; c.px = px;
        for (int r = 0; r < R; ++r) {
          if ((py + pad - r) % cs != 0) continue;
          for (int s = 0; s < S; ++s) {
            if ((px + pad - s) % cs != 0) continue;
            c.oy[c.ntaps] = (short)((py + pad - r) / cs); c.ox[c.ntaps] = (short)((px + pad - s) / cs); c.widx[c.ntaps] = (short)(r * S + s);
            ++c.ntaps;
          }
        }
      }
  }
  static const bool no_halo = getenv("AWR_B200_NO_HALO") != nullptr;          // debugging switch, read once
  if (conv_halo_supported(g) && !no_halo) return conv_halo_launch(g, in, w, bias, out, stats, (cudaStream_t)stream);

  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.Hc = g.Hc; p.Wc = g.Wc;
  p.Wt = p.Wc < 128 ? p.Wc : 128;
  p.Ht = (128 / p.Wt) < p.Hc ? (128 / p.Wt) : p.Hc;
  p.Nt = 128 / (p.Wt * p.Ht);
  p.lgWt = ilog2(p.Wt); p.lgHt = ilog2(p.Ht);
  p.tiles_w = p.Wc / p.Wt; p.tiles_h = p.Hc / p.Ht; p.tiles_n_img = (N + p.Nt - 1) / p.Nt;
  p.tiles_m = p.tiles_w * p.tiles_h * p.tiles_n_img;
  p.kblocks = Ck / 64;
  p.a_stride = g.a_stride;
  AWR_HOST_CHECK(p.Wt * p.a_stride <= 256 && p.Ht * p.a_stride <= 256);
  p.b_mn = g.b_mn;
  p.Ho = Ho; p.Wo = Wo; p.Cn = Cn; p.out_s = cs;
  p.out_mode = out_mode; p.n_valid = n_valid; p.accumulate = accumulate;
  p.nclasses = g.nclasses;
  for (int c = 0; c < g.nclasses; ++c) p.cls[c] = g.cls[c];
  // N tile: the largest of {256,128,64} dividing Cn that still yields >= 148 tiles, else the smallest
  const int cand[3] = {256, 128, 64};
  p.Ntile = 64;
  for (int i = 0; i < 3; ++i) {
    if (Cn % cand[i]) continue;
    if ((long long)p.nclasses * p.tiles_m * (Cn / cand[i]) >= 148 || cand[i] == 64) { p.Ntile = cand[i]; break; }
  }
  p.tiles_c = Cn / p.Ntile;
  const int stage_bytes = kABytes + p.Ntile * 128;
  const int stats_bytes = stats ? 4 * 2 * Cn * (int)sizeof(float) : 0;
  p.stages = (200 * 1024 - stats_bytes) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 + stats_bytes;

  CUtensorMap tmA, tmB;
  {
    const long long dims[4] = {Ck, Wi, Hi, N};
    const long long str[4] = {1, Ck, (long long)Wi * Ck, (long long)Hi * Wi * Ck};
    const int box[4] = {64, p.Wt * p.a_stride, p.Ht * p.a_stride, p.Nt};
    const int es[4] = {1, p.a_stride, p.a_stride, 1};
    if (!make_tmap_bf16(&tmA, in, 4, dims, str, box, es)) return AWR_ERR_DRIVER;
  }
  if (!conv_make_weight_map(&tmB, g, w, p.Ntile)) return AWR_ERR_DRIVER;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int total_tiles = p.nclasses * p.tiles_m * p.tiles_c;
  int sms = awr_sm_budget();
  int grid = total_tiles < sms ? total_tiles : sms;
  if (launch_pdl(conv_tc_kernel, dim3(grid), dim3(kThreads), smem, (cudaStream_t)stream, tmA, tmB, bias, out, reinterpret_cast<AwrAcc*>(stats), p) != cudaSuccess) return (int)cudaGetLastError();
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

}  // extern "C"
