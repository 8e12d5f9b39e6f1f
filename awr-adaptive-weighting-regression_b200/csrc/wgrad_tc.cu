// Weight-gradient GEMM of Conv2d / ConvTranspose2d on tcgen05 tensor cores (sm_100a), NHWC bf16 activations, fp32 dW.
//
//   dW[tap][i*s_p + j*s_g] += sum over coarse pixels q of  pointwise[q, i] * gathered[q*stride - pad + tap, j]
//
// as a GEMM whose contraction dimension is the PIXEL index:  D[m = (tap, gathered channel j)][n = pointwise channel i].
// Both operands are consumed as MN-major UMMA operands straight from TMA boxes of the NHWC tensors (channels contiguous,
// pixels stride 128 B in smem) -- no transposed copy of any activation exists.
//
// Operand reuse (the per-SM L2 port, ~64 B/clk, is the binding resource -- DESIGN.md section 8): a CTA owns a GROUP of taps
// that share one halo patch of the gathered tensor per 64-pixel block (8x8 coarse pixels): UMMA applies the 128-B swizzle on
// absolute smem address bits, so each tap is just a row-shifted window (start = halo + ((dy*PW + dx) * 128 B), K-group stride
// SBO = PW * 128 B) of the same TMA-written tile, and ONE pointwise tile feeds every tap of the group.  Groups: all taps of a
// stride-1 conv when they fit TMEM (Cin = 64: five M tiles of two taps each), one tap row otherwise; for stride-2 layers the
// taps of one parity class (their gathers are unit shifts of the same stride-2 sub-grid, loaded with TMA element strides).
// M tile = 128 rows = two 64-channel blocks (two taps of a 64-channel tensor, or two channel blocks of one tap) stitched by the
// descriptor's leading-dimension offset.  Accumulators: n_mtiles x Ntile fp32 TMEM columns (<= 512), one pass per CTA; the
// pixel range is split across CTAs (split-K) and partial tiles are combined with fp32 red.global.add.
//
// Replaces the weight part of autograd's convolution_backward for model/resnet_deconv.py and model/hourglass.py layers
// in the bf16 precision mode.  Same argument meaning as awr_conv_wgrad_simt.
#include "tc_common.cuh"
#include "awr_b200.h"

namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int kMaxMTiles = 5;
constexpr int kMaxGroups = 9;

// 16-byte vector reduction (sm_90+): one L2 atomic transaction for four consecutive floats
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// deterministic build: four order-independent accumulator slots instead of one vector reduction
__device__ __forceinline__ void red_add_v4(AwrAcc* addr, float4 v) { acc_add(addr, v.x); acc_add(addr + 1, v.y); acc_add(addr + 2, v.z); acc_add(addr + 3, v.w); }

struct MTile {
  int tap[2];            // weight tap of each 64-row block
  int ch[2];             // gathered-channel offset (within the CTA's channel block) of each 64-row block
  int aoff[2];           // byte offset of each block's window inside the stage's halo region
  int valid1;            // second block carries real data
};
struct TapGroup {
  int n_mtiles, halo_ox, halo_oy, pw, hrows;
  MTile mt[kMaxMTiles];
};

struct WgradParams {
  int N, Hc, Wc, tiles_w, tiles_h, pix_blocks, Wt, Ht, Nt;   // coarse grid in 64-pixel blocks (8x8x1, or whole small maps x Nt images)
  int Cp, Cg, Ntile, tiles_n, cg_blocks, halo_blocks; // halo_blocks: 64-channel TMA boxes of the gathered tensor per k-block (1 or 2)
  int stride, s_p, s_g, w_tap;
  int ksplit, stages, halo_block_bytes, stage_bytes;
  int ngroups;
  TapGroup grp[kMaxGroups];
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmG0, const __grid_constant__ CUtensorMap tmG1, const __grid_constant__ CUtensorMap tmG2,
                const __grid_constant__ CUtensorMap tmG3, const __grid_constant__ CUtensorMap tmP, GradT* __restrict__ dW,
                const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();                 // the next kernel may start its prologue while this grid runs
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tfull_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_al = smem_raw + (smem_base - smem_u32(smem_raw));

  // work item of this CTA: (tap group, gathered-channel block, pointwise-channel tile, pixel range)
  int item = blockIdx.x;
  const int ks = item % p.ksplit; item /= p.ksplit;
  const int nt = item % p.tiles_n; item /= p.tiles_n;
  const int cgb = item % p.cg_blocks; const int gi = item / p.cg_blocks;
  const TapGroup& G = p.grp[gi];
  const int cg0 = cgb * 64 * p.halo_blocks;
  const int per = (p.pix_blocks + p.ksplit - 1) / p.ksplit;
  const int pb0 = ks * per, pb1 = min(p.pix_blocks, pb0 + per);
  const int iters = max(pb1 - pb0, 0);
  const int hbytes = G.pw * G.hrows * 128;
  // the gathered tensor map of this group (box shape depends on the group's halo extents); slot = gi & 3, the host checks that
  // groups beyond the fourth repeat the box shape of group gi - 4
  const CUtensorMap* tmG = (gi & 3) == 0 ? &tmG0 : ((gi & 3) == 1 ? &tmG1 : ((gi & 3) == 2 ? &tmG2 : &tmG3));

  if (threadIdx.x == 0) {
    tma_prefetch_desc(tmG);
    tma_prefetch_desc(&tmP);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tfull_bar, 1);
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
      const uint32_t tx = (uint32_t)(p.halo_blocks * hbytes + p.Ntile * 128);
      for (int pb = pb0; pb < pb1; ++pb) {
        const int tw = pb % p.tiles_w; int r = pb / p.tiles_w;
        const int th = r % p.tiles_h; const int n = (r / p.tiles_h) * p.Nt;
        const int w0 = tw * p.Wt, h0 = th * p.Ht;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_expect_tx(&full_bar[stage], tx);
        uint8_t* sa = smem_al + (size_t)stage * p.stage_bytes;
        for (int b = 0; b < p.halo_blocks; ++b)
          tma_load_4d(sa + b * p.halo_block_bytes, tmG, &full_bar[stage], cg0 + 64 * b, w0 * p.stride + G.halo_ox, h0 * p.stride + G.halo_oy, n);
        uint8_t* sb = sa + p.halo_blocks * p.halo_block_bytes;
        for (int j = 0; j < p.Ntile / 64; ++j) tma_load_4d(sb + j * 8192, &tmP, &full_bar[stage], nt * p.Ntile + 64 * j, w0, h0, n);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 5) {
    // ======================================= MMA issuer =======================================
    const uint32_t idesc = umma_idesc_bf16(128, p.Ntile, 1, 1);
    const uint32_t sbo = (uint32_t)G.pw * 128u;                         // stride between 8-pixel K groups (one halo row)
    const uint32_t kstep = (2u * sbo) >> 4;                             // UMMA_K = 16 pixels = two halo rows
    const uint32_t b_off = (uint32_t)(p.halo_blocks * p.halo_block_bytes);
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_loop_fence();
      if (elect_one()) {
        const uint32_t sa = smem_base + (uint32_t)(stage * p.stage_bytes);
        const uint64_t bd = umma_desc_sw128(sa + b_off, 8192, 1024);
        for (int m = 0; m < G.n_mtiles; ++m) {
          const MTile& T = G.mt[m];
          const uint64_t ad = umma_desc_sw128(sa + (uint32_t)T.aoff[0], (uint32_t)(T.aoff[1] - T.aoff[0]), sbo);
          const uint32_t d = tmem_base + (uint32_t)(m * p.Ntile);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, ad + (uint64_t)(k * kstep), bd + (uint64_t)(k * 128), idesc, (it | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) {
      if (iters > 0) umma_commit(&tfull_bar);
      else mbar_arrive(&tfull_bar);
    }
    __syncwarp();
  } else {
    // ======================================= epilogue =======================================
    const int q = warp & 3;
    mbar_wait(&tfull_bar, 0);
    tc_fence_after();
    if (iters > 0) {
      for (int m = 0; m < G.n_mtiles; ++m) {
        const MTile& T = G.mt[m];
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * p.Ntile);
        for (int ch = 0; ch < p.Ntile; ch += 32) {
          uint32_t v[32];
          tmem_ld32(t_addr + ch, v);
          tmem_ld_wait();
          // The warp's 32x32 fp32 block goes through shared memory (pipeline stages are idle by now) so that each lane owns FOUR
          // consecutive elements along the contiguous dimension of dW and one red.global.add.v4.f32 covers 16 bytes: a quarter of the
          // reduction instructions, every one on whole 128-B lines.  tb[row][col], row = gathered channel (this lane), col = pointwise ch.
          float* tb = reinterpret_cast<float*>(smem_al) + q * (32 * 36);
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(tb + lane * 36 + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                                                                         __uint_as_float(v[i + 3]));
          __syncwarp();
          const int r0 = q * 32, bk = r0 >> 6;                    // the warp's 32 rows lie inside one 64-row block
          if (bk == 0 || T.valid1) {
            GradT* wbase = dW + (size_t)T.tap[bk] * p.w_tap + (size_t)(cg0 + T.ch[bk] + (r0 & 63)) * p.s_g + (size_t)(nt * p.Ntile + ch) * p.s_p;
            if (p.s_p == 1) {
              // columns contiguous (ConvTranspose2d layout): lane -> (row rr0 + lane/8, cols 4*(lane%8)..+3), 8 iterations of 4 rows
              const int cq = (lane & 7) * 4, rsub = lane >> 3;
#pragma unroll
              for (int it8 = 0; it8 < 8; ++it8) {
                const int rr = it8 * 4 + rsub;
                const float4 x = *reinterpret_cast<const float4*>(tb + rr * 36 + cq);
                red_add_v4(wbase + (size_t)rr * p.s_g + cq, x);
              }
            } else {
              // rows contiguous (Conv2d layout, s_g == 1): lane -> (col cc0 + lane/8, rows 4*(lane%8)..+3)
              const int rq = (lane & 7) * 4, csub = lane >> 3;
#pragma unroll
              for (int it8 = 0; it8 < 8; ++it8) {
                const int cc = it8 * 4 + csub;
                const float4 x = make_float4(tb[rq * 36 + cc], tb[(rq + 1) * 36 + cc], tb[(rq + 2) * 36 + cc], tb[(rq + 3) * 36 + cc]);
                red_add_v4(wbase + (size_t)cc * p.s_p + rq, x);
              }
            }
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

extern "C" {

int awr_conv_wgrad_tc(const void* pointwise, const void* gathered, void* dW, int N, int Hc, int Wc, int Cp, int Hf, int Wf, int Cg, int R,
                      int S, int stride, int pad, int s_p, int s_g, int w_tap, void* stream) {
  AWR_HOST_CHECK(pointwise && gathered && dW && N > 0 && Cp % 64 == 0 && Cg % 64 == 0 && (Cg == 64 || Cg % 128 == 0));
  AWR_HOST_CHECK(R > 0 && S > 0 && R * S <= 16 && (stride == 1 || stride == 2));
  AWR_HOST_CHECK((s_p == 1 && s_g % 4 == 0) || (s_g == 1 && s_p % 4 == 0));      // 16-byte vector reductions need one contiguous dimension
  AWR_HOST_CHECK(is_pow2(Wc) && is_pow2(Hc) && Wc <= 256 && Hc <= 256);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.Hc = Hc; p.Wc = Wc;
  p.Cp = Cp; p.Cg = Cg; p.stride = stride; p.s_p = s_p; p.s_g = s_g; p.w_tap = w_tap;
  p.halo_blocks = (Cg == 64) ? 1 : 2;
  p.cg_blocks = Cg / (64 * p.halo_blocks);
  p.Ntile = (Cp % 128 == 0) ? 128 : 64;
  p.tiles_n = Cp / p.Ntile;
  const int max_mt = 512 / p.Ntile < kMaxMTiles ? 512 / p.Ntile : kMaxMTiles;

  // ---- pixel blocks: 8x8 coarse pixels of one image; smaller maps use the whole map of several images -------------------
  // (maps below 8x8 keep the 64-pixel K block by taking Wt x Ht x Nt = 64 with Nt images; no halo sharing across images is
  //  needed because each image's halo rows are separate TMA box slices -> handled by treating them as 1-tap groups)
  const bool small = (Wc < 8 || Hc < 8);
  const int Wt = small ? Wc : 8, Ht = small ? Hc : 8, Nt = 64 / (Wt * Ht);
  p.Wt = Wt; p.Ht = Ht; p.Nt = Nt;
  p.tiles_w = Wc / Wt; p.tiles_h = Hc / Ht;
  p.pix_blocks = p.tiles_w * p.tiles_h * ((N + Nt - 1) / Nt);

  // ---- tap groups ------------------------------------------------------------------------------------------------------
  struct TapI { int r, s, oy, ox; };
  TapI cls[4][16]; int ncls[4] = {0, 0, 0, 0};
  const int nclass = stride * stride;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      // gathered pixel = q*stride - pad + (r,s): parity class of the tap and its unit shift inside the stride-sub-grid
      const int ty = r - pad, tx = s - pad;
      const int cy = ((ty % stride) + stride) % stride, cx = ((tx % stride) + stride) % stride;
      const int c = cy * stride + cx;
      cls[c][ncls[c]++] = TapI{r, s, (ty - cy) / stride, (tx - cx) / stride};
    }
  int ng = 0;
  for (int c = 0; c < nclass; ++c) {
    if (ncls[c] == 0) continue;
    // taps per group: as many as fit TMEM; groups are cut at tap-row boundaries when a whole class does not fit
    const int taps_per_mt = (Cg == 64) ? 2 : 1;
    int start = 0;
    while (start < ncls[c]) {
      int end = ncls[c];
      if (small) end = start + 1;
      else if ((end - start + taps_per_mt - 1) / taps_per_mt > max_mt) {
        end = start;                              // take whole tap rows while they fit
        while (end < ncls[c]) {
          int e2 = end; const int row_r = cls[c][end].r;
          while (e2 < ncls[c] && cls[c][e2].r == row_r) ++e2;
          if ((e2 - start + taps_per_mt - 1) / taps_per_mt > max_mt) break;
          end = e2;
        }
        if (end == start) end = start + taps_per_mt * max_mt < ncls[c] ? start + taps_per_mt * max_mt : ncls[c];
      }
      AWR_HOST_CHECK(ng < kMaxGroups);
      TapGroup& G = p.grp[ng];
      int mnx = 99, mxx = -99, mny = 99, mxy = -99;
      for (int i = start; i < end; ++i) { mnx = min(mnx, cls[c][i].ox); mxx = max(mxx, cls[c][i].ox); mny = min(mny, cls[c][i].oy); mxy = max(mxy, cls[c][i].oy); }
      G.pw = Wt + (mxx - mnx); G.hrows = Ht * Nt + (mxy - mny);   // small maps: Nt > 1 only with a single tap (no extents)
      const int cy = c / stride, cx = c % stride;
      G.halo_ox = mnx * stride + cx; G.halo_oy = mny * stride + cy;
      const int hb = G.pw * G.hrows * 128;
      if (hb > p.halo_block_bytes) p.halo_block_bytes = hb;
      int m = 0;
      for (int i = start; i < end; i += taps_per_mt) {
        MTile& T = G.mt[m++];
        const int o0 = ((cls[c][i].oy - mny) * G.pw + (cls[c][i].ox - mnx)) * 128;
        T.tap[0] = cls[c][i].r * S + cls[c][i].s; T.aoff[0] = o0; T.ch[0] = 0;
        if (taps_per_mt == 2) {
          if (i + 1 < end) {
            T.tap[1] = cls[c][i + 1].r * S + cls[c][i + 1].s;
            T.aoff[1] = ((cls[c][i + 1].oy - mny) * G.pw + (cls[c][i + 1].ox - mnx)) * 128; T.ch[1] = 0; T.valid1 = 1;
          } else { T.tap[1] = T.tap[0]; T.aoff[1] = o0; T.ch[1] = 0; T.valid1 = 0; }
        } else { T.tap[1] = T.tap[0]; T.aoff[1] = -1; T.ch[1] = 64; T.valid1 = 1; }   // second channel block: fixed up below
      }
      G.n_mtiles = m;
      ++ng;
      start = end;
    }
  }
  p.ngroups = ng;
  p.halo_block_bytes = (p.halo_block_bytes + 1023) & ~1023;
  for (int g = 0; g < ng; ++g)
    for (int m = 0; m < p.grp[g].n_mtiles; ++m)
      if (p.grp[g].mt[m].aoff[1] < 0) p.grp[g].mt[m].aoff[1] = p.grp[g].mt[m].aoff[0] + p.halo_block_bytes;

  // ---- tensor maps: one per group slot (gi & 3); groups beyond 4 must repeat the box shape of group gi-4 ---------------
  CUtensorMap tmG[4], tmP;
  for (int g = 0; g < 4; ++g) {
    const TapGroup& G = p.grp[g < ng ? g : 0];
    const long long dims[4] = {Cg, Wf, Hf, N};
    const long long str[4] = {1, Cg, (long long)Wf * Cg, (long long)Hf * Wf * Cg};
    int box[4], es[4] = {1, stride, stride, 1};
    if (!small) { box[0] = 64; box[1] = G.pw * stride; box[2] = G.hrows * stride; box[3] = 1; }
    else { box[0] = 64; box[1] = Wt * stride; box[2] = Ht * stride; box[3] = Nt; }
    AWR_HOST_CHECK(box[1] <= 256 && box[2] <= 256);
    if (!make_tmap_bf16(&tmG[g], gathered, 4, dims, str, box, es)) return AWR_ERR_DRIVER;
  }
  for (int g = 4; g < ng; ++g) AWR_HOST_CHECK(p.grp[g].pw == p.grp[g - 4].pw && p.grp[g].hrows == p.grp[g - 4].hrows);
  if (small) for (int g = 0; g < ng; ++g) { p.grp[g].pw = 8; p.grp[g].hrows = 8; }   // dense 64-row block: K groups are consecutive 8-row atoms
  {
    const long long dims[4] = {Cp, Wc, Hc, N};
    const long long str[4] = {1, Cp, (long long)Wc * Cp, (long long)Hc * Wc * Cp};
    const int box[4] = {64, Wt, Ht, Nt};
    if (!make_tmap_bf16(&tmP, pointwise, 4, dims, str, box, nullptr)) return AWR_ERR_DRIVER;
  }
  p.stage_bytes = p.halo_blocks * p.halo_block_bytes + p.Ntile * 128;
  p.stages = (200 * 1024) / p.stage_bytes;
  if (p.stages > 8) p.stages = 8;
  AWR_HOST_CHECK(p.stages >= 2);
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;

  const int base_items = ng * p.cg_blocks * p.tiles_n;
  int ksplit = awr_sm_budget() / base_items;                       // whole waves: never more CTAs than SMs unless the base grid already exceeds them
  if (ksplit < 1) ksplit = 1;
  if (ksplit > p.pix_blocks) ksplit = p.pix_blocks;
  p.ksplit = ksplit;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  if (launch_pdl(wgrad_tc_kernel, dim3(base_items * ksplit), dim3(kThreads), smem, (cudaStream_t)stream, tmG[0], tmG[1], tmG[2], tmG[3], tmP, reinterpret_cast<GradT*>(dW), p) != cudaSuccess) return (int)cudaGetLastError();
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

}  // extern "C"
