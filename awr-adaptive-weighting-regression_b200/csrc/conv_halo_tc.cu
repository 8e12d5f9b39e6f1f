// Halo-tile implicit-GEMM convolution on tcgen05 (sm_100a): all taps of a convolution read ONE activation tile.
//
// Hardware fact this kernel is built on (probe: awr_debug_umma_window, tools/dbg_umma_window.py): UMMA applies the
// SWIZZLE_128B pattern on absolute shared-memory address bits, so a K-major operand descriptor may start at ANY 128-byte
// row of a TMA-written tile and use ANY stride between its 8-row groups.  A CTA therefore loads, per 64-channel block, one
// halo patch (16+ey rows x 16+ex cols of pixels, one TMA box, zero OOB fill = padding) and feeds every tap (dy,dx) of the
// convolution as a row-shifted window of it:  start = halo + ((dy-min_dy)*PW + (dx-min_dx) + 8*sub) * 128 B, SBO = PW*128 B.
// Activation traffic per output tile drops by taps*256/((16+ey)(16+ex))  (3x3: 7.1x; 2x2 deconv classes: 3.5x) versus one
// shifted TMA box per tap (csrc/conv_tc.cu), which is bound by the ~64 B/clk/SM L2 port.
//
// Super-tile = 16 cols x 16 rows of one image = two M=128 sub-tiles (cols 0-7 / 8-15, 16 groups of 8 pixels each) sharing every
// weight tile: per (tap, 64-channel block) ONE B tile feeds 8 MMAs (2 sub-tiles x K=64), halving weight traffic per FLOP.
// Accumulators: 2 TMEM stages x 2 sub-tiles x Ntile (<=128) fp32 columns.  Epilogue: TMEM -> registers -> (+bias) -> bf16 ->
// swizzled smem staging tile -> (a) per-channel sum / sum-of-squares for BatchNorm read column-wise from smem, (b) fully
// coalesced NHWC stores (optionally accumulating).  Roles: warps 0-7 epilogue (two groups of four, one per sub-tile), warp 8 TMA producer, warp 9 MMA issuer (the SM arbiter favours
// higher warp ids, so the two latency-critical single-lane roles sit above the ALU-heavy epilogue warps).
//
// Handles every unit-stride gather: stride-1 Conv2d fprop/dgrad, ConvTranspose2d(k4,s2,p1) fprop parity classes, stride-2 Conv2d
// dgrad parity classes, on feature maps >= 16x16 with tap extents <= 2.  Other cases stay on conv_tc_kernel.
#include "tc_common.cuh"
#include "conv_tc_shared.h"
#include "awr_b200.h"

namespace {

using namespace tc;

constexpr int kThreads = 320;                 // 8 epilogue warps (two groups, one per sub-tile) + TMA producer warp + MMA issuer warp
constexpr int kHaloStageBytes = 44032;       // >= 18*18 rows * 128 B, multiple of 1024
constexpr int kHaloStages = 2;

struct HaloClass {
  int ntaps, py, px, min_ox, min_oy, pw, halo_bytes, pad_;
  short aoff[kConvMaxTaps];        // tap window start inside the halo, in 16-byte units
  short widx[kConvMaxTaps];
};

struct HaloParams {
  int N, Hc, Wc, st_w, st_h, items_m;
  int tiles_c, Ntile, kblocks, b_mn, b_stages;
  int Ho, Wo, Cn, out_s, out_mode, n_valid, accumulate;
  int nclasses;
  HaloClass cls[kConvMaxClasses];
};

struct HaloMaps { CUtensorMap a[kConvMaxClasses]; };

#ifdef AWR_CONV_PROFILE
// debug-only role timers (cycles per CTA): [0] producer wait halo-empty, [1] producer wait b-empty, [2] mma wait halo-full, [3] mma wait b-full,
// [4] mma wait tmem-empty, [5] epilogue wait tmem-full, [6] epilogue busy, [7] kernel total, [8] items
__device__ unsigned long long g_halo_prof[148 * 16];
#define HPROF_T0() const long long t0__ = clock64()
#define HPROF_ADD(slot) atomicAdd(&g_halo_prof[blockIdx.x * 16 + (slot)], (unsigned long long)(clock64() - t0__))
#else
#define HPROF_T0()
#define HPROF_ADD(slot)
#endif

__device__ __forceinline__ void epi_bar(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }   // one named barrier per epilogue group

__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ HaloMaps tmA, const __grid_constant__ CUtensorMap tmB, const float* __restrict__ bias,
                 void* __restrict__ outp, float* __restrict__ stats, const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();                 // the next kernel may start its prologue while this grid runs
  __shared__ __align__(8) uint64_t hfull[kHaloStages], hempty[kHaloStages], bfull[8], bempty[8], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef AWR_CONV_PROFILE
  const long long k_t0 = clock64();
#endif
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const int b_bytes = p.Ntile * 128;
  const uint32_t b_base = smem_base + kHaloStages * kHaloStageBytes;
  const uint32_t stg_base = b_base + (uint32_t)(p.b_stages * b_bytes);            // staging tile [128][Ntile] bf16 (1024-aligned)
  float* s_stats = reinterpret_cast<float*>(smem_al + (stg_base - smem_base) + 2 * 128 * p.Ntile * 2);      // behind the two staging tiles
  const int total_items = p.nclasses * p.tiles_c * p.items_m;
  if (stats) {
    for (int i = threadIdx.x; i < 2 * p.Cn; i += kThreads) s_stats[i] = 0.f;
  }
  if (threadIdx.x == 0) {
    for (int c = 0; c < p.nclasses; ++c) tma_prefetch_desc(&tmA.a[c]);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kHaloStages; ++i) { mbar_init(&hfull[i], 1); mbar_init(&hempty[i], 1); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();                    // prologue done; from here on the previous kernel's outputs are visible

  if (warp == 8) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      int hs = 0; uint32_t hph = 0; int bs = 0; uint32_t bph = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int mi = item % p.items_m; int r = item / p.items_m;
        const int ct = r % p.tiles_c; const int c = r / p.tiles_c;
        const int sw = mi % p.st_w; int r2 = mi / p.st_w;
        const int sh = r2 % p.st_h; const int n = r2 / p.st_h;
        const HaloClass& hc = p.cls[c];
        const int c0 = ct * p.Ntile;
        for (int kc = 0; kc < p.kblocks; ++kc) {
          if (hc.ntaps == 0) break;
          { HPROF_T0(); mbar_wait(&hempty[hs], hph ^ 1u); HPROF_ADD(0); }
          mbar_expect_tx(&hfull[hs], (uint32_t)hc.halo_bytes);
          tma_load_4d(smem_al + hs * kHaloStageBytes, &tmA.a[c], &hfull[hs], kc * 64, sw * 16 + hc.min_ox, sh * 16 + hc.min_oy, n);
          if (++hs == kHaloStages) { hs = 0; hph ^= 1u; }
          for (int t = 0; t < hc.ntaps; ++t) {
            { HPROF_T0(); mbar_wait(&bempty[bs], bph ^ 1u); HPROF_ADD(1); }
            mbar_expect_tx(&bfull[bs], (uint32_t)b_bytes);
            uint8_t* sb = smem_al + (b_base - smem_base) + (size_t)bs * b_bytes;
            const int wi = hc.widx[t];
            if (!p.b_mn) {
              tma_load_3d(sb, &tmB, &bfull[bs], kc * 64, c0, wi);
            } else {
              for (int j = 0; j < p.Ntile / 64; ++j) tma_load_3d(sb + j * 8192, &tmB, &bfull[bs], c0 + 64 * j, kc * 64, wi);
            }
            if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ======================================= MMA issuer =======================================
    const uint32_t idesc = umma_idesc_bf16(128, p.Ntile, 0, p.b_mn);
    const uint64_t bdesc0 = p.b_mn ? umma_desc_sw128(b_base, 8192, 1024) : umma_desc_sw128(b_base, 16, 1024);
    const uint32_t bstep = p.b_mn ? (2048u >> 4) : (32u >> 4);
    const uint32_t bsstep = (uint32_t)b_bytes >> 4;
    int hs = 0; uint32_t hph = 0; int bs = 0; uint32_t bph = 0;
    int as = 0; uint32_t aphase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int c = (item / p.items_m) / p.tiles_c;
      const HaloClass& hc = p.cls[c];
      const int ntaps = hc.ntaps;
      // A descriptor template of this class: SBO = halo row pitch (pw pixels * 128 B)
      const uint64_t adesc0 = umma_desc_sw128(smem_base, 16, (uint32_t)hc.pw * 128u);
      { HPROF_T0(); mbar_wait(&tempty[as], aphase ^ 1u); if (lane == 0) HPROF_ADD(4); }
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(as * 2 * p.Ntile), d1 = d0 + (uint32_t)p.Ntile;
      uint32_t first = 1;
      for (int kc = 0; kc < p.kblocks && ntaps > 0; ++kc) {
        { HPROF_T0(); mbar_wait(&hfull[hs], hph); if (lane == 0) HPROF_ADD(2); }
        tc_fence_after();
        const uint64_t ah = adesc0 + (uint64_t)((uint32_t)(hs * kHaloStageBytes) >> 4);
        // taps are issued in pairs (16 MMAs per elected-lane block): the ~130-cycle issue bubble per block (probe:
        // tools/dbg_umma_pipe.py) is then paid once per two weight tiles
        for (int t = 0; t < ntaps; t += 2) {
          const bool two = (t + 1 < ntaps);
          int bs2 = bs + 1; uint32_t bph2 = bph;
          if (bs2 == p.b_stages) { bs2 = 0; bph2 ^= 1u; }
          { HPROF_T0(); mbar_wait(&bfull[bs], bph); if (two) mbar_wait(&bfull[bs2], bph2); if (lane == 0) HPROF_ADD(3); }
          tc_fence_after();
          if (elect_one()) {
            HPROF_T0();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (u == 1 && !two) break;
              const int tt = t + u, sb_ = u ? bs2 : bs;
              const uint64_t a0 = ah + (uint64_t)(uint32_t)hc.aoff[tt];
              const uint64_t a1 = a0 + 64;                       // sub-tile 1 = 8 pixels (1024 B) to the right
              const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)sb_ * bsstep);
              umma_bf16(d0, a0, bd, idesc, first ? 0u : 1u);
              umma_bf16(d1, a1, bd, idesc, first ? 0u : 1u);
              first = 0;
#pragma unroll
              for (int k = 1; k < 4; ++k) {
                umma_bf16(d0, a0 + 2 * k, bd + k * bstep, idesc, 1u);
                umma_bf16(d1, a1 + 2 * k, bd + k * bstep, idesc, 1u);
              }
              umma_commit(&bempty[sb_]);
              if (tt == ntaps - 1) umma_commit(&hempty[hs]);
            }
            HPROF_ADD(9);
          }
          __syncwarp();
          first = 0;
          bs = bs2; bph = bph2;
          if (two) { if (++bs == p.b_stages) { bs = 0; bph ^= 1u; } }
        }
        if (++hs == kHaloStages) { hs = 0; hph ^= 1u; }
      }
      if (elect_one()) {
        if (ntaps > 0) umma_commit(&tfull[as]);
        else mbar_arrive(&tfull[as]);
      }
      __syncwarp();
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else {
    // ======================================= epilogue =======================================
    // two epilogue groups of four warps: group `grp` drains sub-tile `grp` (its own staging tile and named barrier), so the two
    // sub-tiles of a super-tile are converted, reduced and stored concurrently -- the epilogue, not the MMA stream, bounds these kernels
    const int q = warp & 3, grp = warp >> 2;
    const int et = q * 32 + lane;                 // 0..127: epilogue thread id within the group == accumulator row (TMEM lane)
    const int g = et >> 3, j = et & 7;            // row = group g (image row h0+g), pixel j inside the group
    uint8_t* const stg = smem_al + (stg_base - smem_base) + grp * (128 * p.Ntile * 2);
    const int row_bytes = p.Ntile * 2, chunks16 = row_bytes >> 4;        // 16-byte chunks per staged row (8 or 16)
    int as = 0; uint32_t aphase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int mi = item % p.items_m; int r = item / p.items_m;
      const int ct = r % p.tiles_c; const int c = r / p.tiles_c;
      const int sw = mi % p.st_w; int r2 = mi / p.st_w;
      const int sh = r2 % p.st_h; const int n = r2 / p.st_h;
      const int c0 = ct * p.Ntile;
      const bool has_acc = p.cls[c].ntaps > 0;
      const int py = p.cls[c].py, px = p.cls[c].px;
      { HPROF_T0(); mbar_wait(&tfull[as], aphase); if (et == 0 && grp == 0) HPROF_ADD(5); }
#ifdef AWR_CONV_PROFILE
      const long long e_t0 = clock64();
#endif
      tc_fence_after();
      {
        const int sub = grp;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * 2 + sub) * p.Ntile);
        const int hc_ = sh * 16 + g, wc_ = sw * 16 + sub * 8 + j;
        const int ho = hc_ * p.out_s + py, wo = wc_ * p.out_s + px;
        for (int ch = 0; ch < p.Ntile; ch += 32) {
          uint32_t v[32];
          if (has_acc) { tmem_ld32(t_addr + ch, v); tmem_ld_wait(); }
          else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0u;
          }
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
          if (p.out_mode == 1) {      // fp32 NCHW planes (prediction volume): straight from registers
            float* dst = reinterpret_cast<float*>(outp);
            const size_t P = (size_t)p.Ho * p.Wo, pix = (size_t)ho * p.Wo + wo;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int cc = c0 + ch + i;
              if (cc < p.n_valid) dst[((size_t)n * p.n_valid + cc) * P + pix] = f[i];
            }
          } else {                    // bf16 -> swizzled staging row `et`: 16-B chunk index XOR (row & (chunks16-1))
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              uint4 pk;
              __nv_bfloat162 h0 = __floats2bfloat162_rn(f[8 * k4], f[8 * k4 + 1]), h1 = __floats2bfloat162_rn(f[8 * k4 + 2], f[8 * k4 + 3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(f[8 * k4 + 4], f[8 * k4 + 5]), h3 = __floats2bfloat162_rn(f[8 * k4 + 6], f[8 * k4 + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
              const int lc = (ch >> 3) + k4;                                  // logical 16-B chunk in the row
              const int pc = lc ^ (et & (chunks16 - 1));
              *reinterpret_cast<uint4*>(stg + et * row_bytes + pc * 16) = pk;
            }
          }
        }
        {                             // this warp's share of the accumulator stage is out of TMEM: hand it back to the MMA warp (8 arrivals)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[as]);
        }
        if (p.out_mode != 1) {
        epi_bar(grp);
        // (a) BatchNorm statistics: thread -> one column pair, a slice of rows; reads are bank-conflict free
        if (stats) {
          const int pairs = p.Ntile >> 1, slices = 128 / pairs, rows_per = 128 / slices;
          const int cp = et % pairs, sl = et / pairs;
          float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
          const int lcs = cp >> 2, wofs = (cp & 3) * 4;
          for (int rr = sl * rows_per; rr < (sl + 1) * rows_per; rr += 8) {      // rows_per is a multiple of 8: 8 independent loads in flight
            uint32_t u[8];
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) {
              const int r8 = rr + k8;
              u[k8] = *reinterpret_cast<const uint32_t*>(stg + r8 * row_bytes + ((lcs ^ (r8 & (chunks16 - 1))) << 4) + wofs);
            }
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) {
              const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[k8]));
              s1a += x.x; s1b += x.y; s2a += x.x * x.x; s2b += x.y * x.y;
            }
          }
          atomicAdd(s_stats + c0 + 2 * cp, s1a); atomicAdd(s_stats + c0 + 2 * cp + 1, s1b);
          atomicAdd(s_stats + p.Cn + c0 + 2 * cp, s2a); atomicAdd(s_stats + p.Cn + c0 + 2 * cp + 1, s2b);
        }
        // (b) coalesced stores: `chunks16` consecutive threads cover one pixel's Ntile channels (256 / 128 contiguous bytes)
        {
          const int tpp = chunks16, ppp = 128 / tpp;                          // threads per pixel, pixels per pass
          const int lc = et % tpp, pr = et / tpp;
          bf16* const outb = reinterpret_cast<bf16*>(outp);
          if (!p.accumulate) {
            for (int pass = 0; pass < 128 / ppp; pass += 4) {                  // 128/ppp is 8 or 16: four loads in flight, then four stores
              uint4 val[4]; bf16* dst[4];
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const int rr = (pass + k4) * ppp + pr;
                const int gg = rr >> 3, jj = rr & 7;
                const int ho2 = (sh * 16 + gg) * p.out_s + py, wo2 = (sw * 16 + sub * 8 + jj) * p.out_s + px;
                val[k4] = *reinterpret_cast<const uint4*>(stg + rr * row_bytes + ((lc ^ (rr & (chunks16 - 1))) << 4));
                dst[k4] = outb + (((size_t)n * p.Ho + ho2) * p.Wo + wo2) * p.Cn + c0 + lc * 8;
              }
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) *reinterpret_cast<uint4*>(dst[k4]) = val[k4];
            }
          } else {
            // accumulate into the existing gradient: four read-modify-writes in flight per thread (the global loads dominate)
            for (int pass = 0; pass < 128 / ppp; pass += 4) {
              uint4 val[4], old[4]; bf16* dst[4];
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const int rr = (pass + k4) * ppp + pr;
                const int gg = rr >> 3, jj = rr & 7;
                const int ho2 = (sh * 16 + gg) * p.out_s + py, wo2 = (sw * 16 + sub * 8 + jj) * p.out_s + px;
                dst[k4] = outb + (((size_t)n * p.Ho + ho2) * p.Wo + wo2) * p.Cn + c0 + lc * 8;
                old[k4] = *reinterpret_cast<const uint4*>(dst[k4]);
                val[k4] = *reinterpret_cast<const uint4*>(stg + rr * row_bytes + ((lc ^ (rr & (chunks16 - 1))) << 4));
              }
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&val[k4]);
                const __nv_bfloat162* ho_ = reinterpret_cast<const __nv_bfloat162*>(&old[k4]);
                uint4 r;
                __nv_bfloat162* hr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 a = __bfloat1622float2(hv[i]), b = __bfloat1622float2(ho_[i]);
                  hr[i] = __floats2bfloat162_rn(a.x + b.x, a.y + b.y);
                }
                *reinterpret_cast<uint4*>(dst[k4]) = r;
              }
            }
          }
        }
        epi_bar(grp);                 // staging tile free for the next item
        }
      }
#ifdef AWR_CONV_PROFILE
      if (et == 0 && grp == 0) { atomicAdd(&g_halo_prof[blockIdx.x * 16 + 6], (unsigned long long)(clock64() - e_t0)); atomicAdd(&g_halo_prof[blockIdx.x * 16 + 8], 1ull); }
#endif
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (stats) {
    for (int i = threadIdx.x; i < 2 * p.Cn; i += kThreads) {
      const float v = s_stats[i];
      if (v != 0.f) atomicAdd(stats + i, v);
    }
  }
  if (warp == 9) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
#ifdef AWR_CONV_PROFILE
  if (threadIdx.x == 0) atomicAdd(&g_halo_prof[blockIdx.x * 16 + 7], (unsigned long long)(clock64() - k_t0));
#endif
}

}  // namespace

#ifdef AWR_CONV_PROFILE
extern "C" int awr_debug_halo_profile(unsigned long long* out_host, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out_host, g_halo_prof, sizeof(g_halo_prof));
  if (reset) { static unsigned long long z[148 * 16]; cudaMemcpyToSymbol(g_halo_prof, z, sizeof(z)); }
  return 0;
}
#endif

bool conv_halo_supported(const ConvGeom& g) {
  if (g.a_stride != 1 || g.Hc < 16 || g.Wc < 16 || (g.Hc % 16) || (g.Wc % 16)) return false;
  for (int c = 0; c < g.nclasses; ++c) {
    const ConvTapClass& t = g.cls[c];
    if (t.ntaps == 0) continue;
    int mnx = 99, mxx = -99, mny = 99, mxy = -99;
    for (int i = 0; i < t.ntaps; ++i) { mnx = min(mnx, (int)t.ox[i]); mxx = max(mxx, (int)t.ox[i]); mny = min(mny, (int)t.oy[i]); mxy = max(mxy, (int)t.oy[i]); }
    if (mxx - mnx > 2 || mxy - mny > 2) return false;
  }
  return true;
}

int conv_halo_launch(const ConvGeom& g, const void* in, const void* w, const float* bias, void* out, float* stats, cudaStream_t stream) {
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.N = g.N; p.Hc = g.Hc; p.Wc = g.Wc; p.st_w = g.Wc / 16; p.st_h = g.Hc / 16; p.items_m = g.N * p.st_w * p.st_h;
  p.kblocks = g.Ck / 64; p.b_mn = g.b_mn;
  p.Ho = g.Ho; p.Wo = g.Wo; p.Cn = g.Cn; p.out_s = g.out_s; p.out_mode = g.out_mode; p.n_valid = g.n_valid; p.accumulate = g.accumulate;
  p.nclasses = g.nclasses;
  // N tile <= 128 (two sub-tiles x two TMEM stages); prefer 128 unless that leaves most SMs without a work item
  p.Ntile = (g.Cn % 128 == 0 && (long long)g.nclasses * p.items_m * (g.Cn / 128) >= 120) ? 128 : 64;
  if (g.Cn % 128 != 0) p.Ntile = 64;
  p.tiles_c = g.Cn / p.Ntile;
  HaloMaps maps;
  for (int c = 0; c < g.nclasses; ++c) {
    const ConvTapClass& t = g.cls[c];
    HaloClass& h = p.cls[c];
    h.ntaps = t.ntaps; h.py = t.py; h.px = t.px;
    int mnx = 0, mxx = 0, mny = 0, mxy = 0;
    if (t.ntaps > 0) {
      mnx = mny = 99; mxx = mxy = -99;
      for (int i = 0; i < t.ntaps; ++i) { mnx = min(mnx, (int)t.ox[i]); mxx = max(mxx, (int)t.ox[i]); mny = min(mny, (int)t.oy[i]); mxy = max(mxy, (int)t.oy[i]); }
    }
    h.min_ox = mnx; h.min_oy = mny; h.pw = 16 + (mxx - mnx);
    const int hrows = 16 + (mxy - mny);
    h.halo_bytes = hrows * h.pw * 128;
    for (int i = 0; i < t.ntaps; ++i) { h.aoff[i] = (short)((((int)t.oy[i] - mny) * h.pw + ((int)t.ox[i] - mnx)) * 8); h.widx[i] = t.widx[i]; }
    const long long dims[4] = {g.Ck, g.Wi, g.Hi, g.N};
    const long long str[4] = {1, g.Ck, (long long)g.Wi * g.Ck, (long long)g.Hi * g.Wi * g.Ck};
    const int box[4] = {64, h.pw, hrows, 1};
    if (!make_tmap_bf16(&maps.a[c], in, 4, dims, str, box, nullptr)) return AWR_ERR_DRIVER;
  }
  for (int c = g.nclasses; c < kConvMaxClasses; ++c) maps.a[c] = maps.a[0];
  CUtensorMap tmB;
  if (!conv_make_weight_map(&tmB, g, w, p.Ntile)) return AWR_ERR_DRIVER;
  const int b_bytes = p.Ntile * 128, stg_bytes = 2 * 128 * p.Ntile * 2, stats_bytes = stats ? 2 * g.Cn * (int)sizeof(float) : 0;
  int bst = (205 * 1024 - kHaloStages * kHaloStageBytes - stg_bytes - stats_bytes) / b_bytes;
  if (bst > 8) bst = 8;
  if (bst < 2) return AWR_ERR_UNSUPPORTED;
  p.b_stages = bst;
  const size_t smem = (size_t)kHaloStages * kHaloStageBytes + (size_t)bst * b_bytes + stg_bytes + stats_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int total = p.nclasses * p.tiles_c * p.items_m;
  const int grid = total < 148 ? total : 148;
  if (launch_pdl(conv_halo_kernel, dim3(grid), dim3(kThreads), smem, stream, maps, tmB, bias, out, stats, p) != cudaSuccess) return (int)cudaGetLastError();
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}
