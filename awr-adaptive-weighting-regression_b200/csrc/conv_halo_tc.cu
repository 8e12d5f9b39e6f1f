// Halo-tile implicit-GEMM convolution on tcgen05 (sm_100a): all taps of a convolution read ONE activation tile.
//
// Hardware fact this kernel is built on (probe: awr_debug_umma_window, tools/dbg_umma_window.py): UMMA applies the
// SWIZZLE_128B pattern on absolute shared-memory address bits, so a K-major operand descriptor may start at ANY 128-byte
// row of a TMA-written tile and use ANY stride between its 8-row groups.  A CTA therefore loads, per 64-channel block, one
// halo patch (16+ey rows x 16+ex cols of pixels, one TMA box, zero OOB fill = padding) and feeds every tap (dy,dx) of the
// convolution as a row-shifted window of it:  start = halo + ((dy-min_dy)*PW + (dx-min_dx) + 8*sub) * 128 B, SBO = PW*128 B.
// Activation traffic per output tile drops by taps*256/((16+ey)(16+ex))  (3x3: 7.1x; 2x2 deconv classes: 3.5x) versus one
// shifted TMA box per tap (csrc/conv_tc.cu), which is bound by the ~42 B/clk/SM L2 port.
//
// Super-tile = 16 cols x 16 rows of one image = two M=128 sub-tiles (cols 0-7 / 8-15, 16 groups of 8 pixels each) sharing every
// weight tile: per (tap, 64-channel block) ONE B tile feeds 8 MMAs (2 sub-tiles x K=64), halving weight traffic per FLOP.
// Accumulators: 2 TMEM stages x 2 sub-tiles x Ntile (<=128) fp32 columns (exactly 4*Ntile columns are allocated).
//
// Round-2 structure (profiles/r02_probes.md: the round-1 kernel was bound by bytes in flight, not by the tensor pipe):
//  * halo patches and weight tiles have their own producer warps and rings (one thread issues one TMA per ~570 cycles whatever the box
//    size, tools/dbg_tma_rate.py): halo patches run 2-3 k-steps ahead instead of waiting behind the weight ring, and weight tiles travel
//    as pair stages (two taps per barrier round);
//  * layers whose whole weight set fits (Cin = Cout = 64, 3x3: 72 KB) keep it RESIDENT in shared memory for the CTA's lifetime:
//    per item only the 41 KB halo patch moves, and no weight-ring barrier traffic remains;
//  * a CTA owns ONE output-channel tile for all its items, so the BatchNorm statistics live in registers across items: the
//    coalesced-store pass (each thread re-reads one 16-byte chunk of the staged tile) also accumulates sum / sum-of-squares of its
//    8 channels -- no separate statistics pass, no shared-memory atomics -- and one fixed-order reduction per CTA feeds the
//    order-independent global accumulators (AwrAcc, common.cuh): the statistics are bit-reproducible;
//  * the epilogue works in 64-column halves through a 2 x 16 KB staging area (was 2 x 32 KB), which pays for the deeper rings.
// Roles: warps 0-7 epilogue (two groups of four, one per sub-tile), warp 8 halo producer, warp 10 weight producer, warps 9 / 11 MMA issuers
// (one per sub-tile: a lone thread needs ~40 cycles per tcgen05.mma it issues plus ~450 per barrier round, tools/dbg_umma_queue.py -- as
// long as a 48-cycle N=64 MMA takes to execute -- so each sub-tile's accumulator gets its own issuing warp).
//
// Handles every unit-stride gather: stride-1 Conv2d fprop/dgrad, ConvTranspose2d(k4,s2,p1) fprop parity classes, stride-2 Conv2d
// dgrad parity classes, on feature maps >= 16x16 with tap extents <= 2.  Other cases stay on conv_tc_kernel.
#include "tc_common.cuh"
#include "conv_tc_shared.h"
#include "awr_b200.h"

namespace {

using namespace tc;

constexpr int kThreads = 384;                 // 8 epilogue warps (two groups, one per sub-tile) + halo producer + MMA issuer 0 + weight producer + MMA issuer 1
constexpr int kMaxHaloStages = 3;
constexpr int kMaxBStages = 10;               // ring depth, or the resident tap count
constexpr int kStagingBytes = 2 * 128 * 128;  // two epilogue groups x 128 rows x 64 bf16
constexpr int kSmemBudget = 226 * 1024;       // dynamic shared memory incl. 1 KB alignment slack (227 KB per CTA on sm_100)

// Tap grid: every tap class of a launch is described on ONE (E+1) x (E+1) grid of unit shifts (E = the largest window extent of the launch:
// 3x3 stride-1 -> 2, ConvTranspose2d / stride-2 dgrad parity classes -> 1, 1x1 -> 0) anchored at the class's smallest offsets; grid position
// g = gy*(E+1)+gx is present when bit g of `mask` is set and then uses weight tap widx[g].  All classes share the halo box (16+E)^2 pixels,
// so the A-window offset of a grid position, ((gy*(16+E) + gx) * 128 B, is a COMPILE-TIME constant: the MMA issue loop is straight-line
// code whose descriptors differ by immediates.  (Round 1 looked the offsets up in a parameter table per tap: ~30 dependent R2UR / LDC
// instructions in front of every 16 MMAs drained the tensor pipe's short queue -- 79 instead of 48 cycles per N=64 MMA,
// profiles/r02_probes.md.)
struct HaloClass {
  int mask, last_g, py, px, min_ox, min_oy, ntaps, pad_;
  short widx[9];
  short pad2_[7];
};

struct HaloParams {
  int N, Hc, Wc, st_w, st_h, items_m;
  int tiles_c, Ntile, kblocks, b_mn, b_stages, resident;
  int halo_stages, halo_stage_bytes, halo_bytes;
  int Ho, Wo, Cn, out_s, out_mode, n_valid, accumulate;
  int nclasses;
  HaloClass cls[kConvMaxClasses];
};

#ifdef AWR_CONV_PROFILE
// timeline of CTA 0 (events: (tag << 48) | cycles since kernel entry; per role, plain stores)
__device__ unsigned long long g_halo_tl[6 * 128];
#define HTL(tag) do { if (blockIdx.x == 0 && tl_n < 128) { g_halo_tl[tl_role * 128 + tl_n++] = ((unsigned long long)(tag) << 48) | (unsigned long long)(clock64() - k_t0); } } while (0)
#else
#define HTL(tag)
#endif

__device__ __forceinline__ void epi_bar(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }   // one named barrier per epilogue group

// poll without suspending: the producer multiplexes two queues and must never sleep on one of them
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"       // NON-blocking (try_wait suspends the thread for up to ~1 us when not ready)
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// the CTA's work items -> (tap class, m item).  A CTA keeps ONE channel tile for its whole life.  CL = CTA pairs (clusters of 2): the
// two CTAs of a pair take neighbouring items (2q, 2q+1: same class, same channel tile), so they need the SAME weight tiles in the same
// order and each tile is fetched from L2 once per pair (TMA multicast).
template <bool CL>
struct ItemIter {
  int j, step, total;
  __device__ __forceinline__ ItemIter(const HaloParams& p, int rank) {
    if (CL) { const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1; j = 2 * (cid / p.tiles_c) + rank; step = 2 * (ncl / p.tiles_c); }
    else { j = blockIdx.x / p.tiles_c; step = gridDim.x / p.tiles_c; }
    total = p.nclasses * p.items_m;
  }
  __device__ __forceinline__ bool valid() const { return j < total; }
  __device__ __forceinline__ void next() { j += step; }
};

// the 4 MMAs of one tap for ONE sub-tile: four K=16 steps on the weight tile `bd` (both MMA warps read the same weight tile; sub-tile 1's
// A window is 8 pixels = 1024 B to the right of sub-tile 0's)
__device__ __forceinline__ void issue_tap(uint32_t d, uint64_t a0, uint64_t bd, uint32_t bstep, uint32_t idesc, uint32_t acc0) {
  umma_bf16(d, a0, bd, idesc, acc0);
#pragma unroll
  for (int k = 1; k < 4; ++k) umma_bf16(d, a0 + 2 * k, bd + k * bstep, idesc, 1u);
}

// NH = Ntile / 64 (64-column halves of the accumulator tile); E = window extent of the tap grid; CL = CTA pairs sharing weight tiles
template <int NH, int E, bool CL>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const float* __restrict__ bias,
                 void* __restrict__ outp, AwrAcc* __restrict__ stats, const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();                 // the next kernel may start its prologue while this grid runs
  __shared__ __align__(8) uint64_t hfull[kMaxHaloStages], hempty[kMaxHaloStages], bfull[kMaxBStages], bempty[kMaxBStages], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  constexpr int Ntile = NH * 64;
  constexpr uint32_t kTmemCols = 4 * Ntile;      // 2 stages x 2 sub-tiles x Ntile: 256 or 512 (powers of two)
  constexpr int GW = E + 1, NT = GW * GW, PW = 16 + E;          // tap grid side / positions, halo row pitch in pixels

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef AWR_CONV_PROFILE
  const long long k_t0 = clock64();
  int tl_n = 0;
  const int tl_role = threadIdx.x == 0 ? 0 : (warp == 8 ? 1 : (warp == 10 ? 5 : (warp == 9 ? 2 : (warp < 4 ? 3 : 4))));
#endif
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int b_bytes = Ntile * 128;
  const uint32_t b_base = smem_base + (uint32_t)(p.halo_stages * p.halo_stage_bytes);
  const uint32_t stg_base = b_base + (uint32_t)(p.b_stages * 2 * b_bytes);        // staging: [2 groups][128 rows][64 bf16] (1024-aligned)
  const int rank = CL ? (int)cluster_ctarank() : 0;
  const int ct = (CL ? (blockIdx.x >> 1) : blockIdx.x) % p.tiles_c, c0 = ct * Ntile;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    // every consumer-side barrier counts BOTH MMA warps (one commit each); in a CTA pair a weight stage is released by all four
    for (int i = 0; i < p.halo_stages; ++i) { mbar_init(&hfull[i], 1); mbar_init(&hempty[i], 2); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], CL ? 4 : 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 2); mbar_init(&tempty[i], 8); }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&tmem_base_s, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (CL) cluster_sync_all();    // the peer's barriers exist before any multicast tile / commit can reach them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) HTL(1);
  pdl_wait();                    // prologue done; from here on the previous kernel's outputs are visible
  if (threadIdx.x == 0) HTL(2);

  // BatchNorm statistics of this thread's 8 channels (per 64-column half), carried across all items of the CTA
  float s1[NH][8], s2[NH][8];
#pragma unroll
  for (int h = 0; h < NH; ++h)
#pragma unroll
    for (int k = 0; k < 8; ++k) { s1[h][k] = 0.f; s2[h][k] = 0.f; }

  if (warp == 8) {
    // ======================================= TMA producer: halo patches =======================================
    // The two operand streams have their own single-thread producers (this warp: one halo patch per (item, 64-channel block); warp 10: the
    // weight tiles), so the halo patches run halo_stages-1 k-steps ahead of the MMAs instead of queueing behind the weight ring, and neither
    // loop pays for the other's bookkeeping: a lone thread retires ~1 instruction per 5 cycles, so everything an issue needs is decoded once
    // per item and the per-TMA path is poll / expect_tx / issue on 32-bit shared addresses (profiles/r02_probes.md).
    if (lane == 0) {
      const uint32_t hfull0 = smem_u32(&hfull[0]);
      int hs = 0; uint32_t hph = 0;
      for (ItemIter<CL> ih(p, rank); ih.valid(); ih.next()) {
        const int c = ih.j / p.items_m, mi = ih.j - c * p.items_m;
        if (p.cls[c].mask == 0) continue;                    // classes without taps produce zeros: nothing to load
        const int sw = mi % p.st_w; const int r2 = mi / p.st_w;
        const int sh = r2 % p.st_h, n = r2 / p.st_h;
        const int x = sw * 16 + p.cls[c].min_ox, y = sh * 16 + p.cls[c].min_oy;
        for (int kc = 0; kc < p.kblocks; ++kc) {
          mbar_wait(&hempty[hs], hph ^ 1u);
          mbar_expect_tx_s(hfull0 + 8u * hs, (uint32_t)p.halo_bytes);
          HTL(10);
          tma_load_4d_s(smem_base + (uint32_t)(hs * p.halo_stage_bytes), &tmA, hfull0 + 8u * hs, kc * 64, x, y, n);
          if (++hs == p.halo_stages) { hs = 0; hph ^= 1u; }
        }
      }
    }
  } else if (warp == 10) {
    // ======================================= TMA producer: weight tiles =======================================
    // A stage of the weight ring holds the (up to) TWO tiles of a grid-position pair (g, g+1) -- what one MMA issue block consumes -- behind
    // ONE full/empty barrier pair: every mbarrier operation costs the issuing thread 100-150 cycles (test_wait 149, B300_MICROARCH.md), and with
    // a barrier round trip per 8 KB tile a lone thread sustained one tile per ~500 cycles, slower than the 384 cycles the MMAs need for it.
    if (lane == 0) {
      const uint32_t bfull0 = smem_u32(&bfull[0]);
      int bs = 0; uint32_t bph = 0;
      int bi = 0;                                             // pair counter: in a CTA pair, CTA (bi & 1) fetches pair bi for both
      for (ItemIter<CL> ib(p, rank); ib.valid(); ib.next()) {
        const HaloClass& hc = p.cls[ib.j / p.items_m];
        const uint32_t mask = (uint32_t)hc.mask;
        if (mask == 0) continue;
        int wv[NT + 1];                                       // weight tap index per grid position (registers: every use is unrolled)
#pragma unroll
        for (int g = 0; g < NT; ++g) wv[g] = (int)hc.widx[g];
        wv[NT] = 0;
        for (int kc = 0; kc < p.kblocks; ++kc) {
#pragma unroll
          for (int g = 0; g < NT; g += 2) {
            const bool t0 = (mask >> g) & 1u, t1 = (g + 1 < NT) && ((mask >> (g + 1)) & 1u);
            if (!(t0 || t1)) continue;
            if (!p.resident) mbar_wait(&bempty[bs], bph ^ 1u);
            mbar_expect_tx_s(bfull0 + 8u * bs, (uint32_t)(((t0 ? 1 : 0) + (t1 ? 1 : 0)) * b_bytes));
            HTL(11);
            if (!CL || (bi & 1) == rank) {                    // CTA pair: this CTA's turn -- one L2 read, delivered to both CTAs
              const uint32_t sb0 = b_base + (uint32_t)(bs * 2 * b_bytes), bar = bfull0 + 8u * bs;
              const int wi0 = wv[g], wi1 = wv[g + 1];
              auto load_tile = [&](uint32_t sb, int wi) {
                if (!p.b_mn) {
                  if (CL) tma_load_3d_mc_s(sb, &tmB, bar, kc * 64, c0, wi, (uint16_t)3);
                  else tma_load_3d_s(sb, &tmB, bar, kc * 64, c0, wi);
                } else {
#pragma unroll
                  for (int j = 0; j < NH; ++j) {
                    if (CL) tma_load_3d_mc_s(sb + j * 8192, &tmB, bar, c0 + 64 * j, kc * 64, wi, (uint16_t)3);
                    else tma_load_3d_s(sb + j * 8192, &tmB, bar, c0 + 64 * j, kc * 64, wi);
                  }
                }
              };
              if (t0) load_tile(sb0, wi0);
              if (t1) load_tile(sb0 + (t0 ? (uint32_t)b_bytes : 0u), wi1);
            }
            ++bi;
            if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
          }
        }
        if (p.resident) break;                                // weights stay in shared memory: nothing more to load
      }
    }
  } else if (warp == 9 || warp == 11) {
    // ======================================= MMA issuers (warp 9: sub-tile 0, warp 11: sub-tile 1) =======================================
    const int sub = warp == 9 ? 0 : 1;
    const uint32_t idesc = umma_idesc_bf16(128, Ntile, 0, p.b_mn);
    const uint64_t bdesc0 = p.b_mn ? umma_desc_sw128(b_base, 8192, 1024) : umma_desc_sw128(b_base, 16, 1024);
    const uint32_t bstep = p.b_mn ? (2048u >> 4) : (32u >> 4);
    constexpr uint32_t bsstep = (uint32_t)b_bytes >> 4;          // one tile; a ring stage holds a pair (2 * bsstep)
    const uint64_t adesc0 = umma_desc_sw128(smem_base + (uint32_t)sub * 1024u, 16, (uint32_t)PW * 128u);        // SBO = halo row pitch
    const uint32_t hstep = (uint32_t)p.halo_stage_bytes >> 4;
    int hs = 0; uint32_t hph = 0; int bs = 0; uint32_t bph = 0;
    int as = 0; uint32_t aphase = 0;
    if (p.resident) {
      // whole weight set resident (one class, one 64-channel block, all NT grid positions present, stage == grid position):
      // per item ONE elected block issues NT*8 MMAs back to back
      bool first_item = true;
      for (ItemIter<CL> it(p, rank); it.valid(); it.next()) {
        mbar_wait(&tempty[as], aphase ^ 1u);
        if (lane == 0 && sub == 0) HTL(20);
        tc_fence_after();
        mbar_wait(&hfull[hs], hph);
        if (lane == 0 && sub == 0) HTL(21);
        const uint32_t d = tmem_base + (uint32_t)((as * 2 + sub) * Ntile);
        const uint64_t ah = adesc0 + (uint64_t)((uint32_t)hs * hstep);
        if (first_item) {               // the weight tiles are still arriving: tap by tap, as they land
          first_item = false;
#pragma unroll
          for (int g = 0; g < NT; ++g) {
            if ((g & 1) == 0) mbar_wait(&bfull[g >> 1], 0);
            if (elect_one()) {
              issue_tap(d, ah + (uint64_t)(((g / GW) * PW + (g % GW)) * 8), bdesc0 + (uint64_t)(g * bsstep), bstep, idesc, g ? 1u : 0u);
              if (g == NT - 1) { umma_commit(&hempty[hs]); umma_commit(&tfull[as]); }
            }
            __syncwarp();
          }
        } else if (elect_one()) {
#pragma unroll
          for (int g = 0; g < NT; ++g)
            issue_tap(d, ah + (uint64_t)(((g / GW) * PW + (g % GW)) * 8), bdesc0 + (uint64_t)(g * bsstep), bstep, idesc, g ? 1u : 0u);
          umma_commit(&hempty[hs]);
          umma_commit(&tfull[as]);
        }
        __syncwarp();
        if (lane == 0 && sub == 0) HTL(22);
        if (++hs == p.halo_stages) { hs = 0; hph ^= 1u; }
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    } else {
      for (ItemIter<CL> it(p, rank); it.valid(); it.next()) {
        const int c = it.j / p.items_m;
        const uint32_t mask = (uint32_t)p.cls[c].mask, last_g = (uint32_t)p.cls[c].last_g;
        mbar_wait(&tempty[as], aphase ^ 1u);
        if (lane == 0 && sub == 0) HTL(20);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)((as * 2 + sub) * Ntile);
        uint32_t acc = 0;                              // 0 for the first tap of the item, 1 afterwards
        for (int kc = 0; kc < p.kblocks && mask != 0; ++kc) {
          mbar_wait(&hfull[hs], hph);
          if (lane == 0 && sub == 0) HTL(21);
          const uint64_t ah = adesc0 + (uint64_t)((uint32_t)hs * hstep);
          // grid positions in pairs: up to 8 MMAs per warp and elected block, A-window offsets are immediates
#pragma unroll
          for (int g = 0; g < NT; g += 2) {
            const bool t0 = (mask >> g) & 1u, t1 = (g + 1 < NT) && ((mask >> (g + 1)) & 1u);
            if (!(t0 || t1)) continue;
            if (lane == 0 && sub == 0) HTL(23);
            mbar_wait(&bfull[bs], bph);
            if (lane == 0 && sub == 0) HTL(24);
            if (elect_one()) {
              const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)bs * 2u * bsstep);
              if (t0) issue_tap(d, ah + (uint64_t)(((g / GW) * PW + (g % GW)) * 8), bd, bstep, idesc, acc);
              if (t1) issue_tap(d, ah + (uint64_t)((((g + 1) / GW) * PW + ((g + 1) % GW)) * 8), bd + (uint64_t)(t0 ? bsstep : 0u), bstep, idesc,
                                t0 ? 1u : acc);
              if (CL) umma_commit_mc(&bempty[bs], (uint16_t)3); else umma_commit(&bempty[bs]);
              if ((uint32_t)g == last_g || (uint32_t)(g + 1) == last_g) umma_commit(&hempty[hs]);
            }
            __syncwarp();
            acc = 1;
            if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
          }
          if (++hs == p.halo_stages) { hs = 0; hph ^= 1u; }
        }
        if (elect_one()) {
          if (mask != 0) umma_commit(&tfull[as]);
          else mbar_arrive(&tfull[as]);
        }
        __syncwarp();
        if (lane == 0 && sub == 0) HTL(22);
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ======================================= epilogue =======================================
    // two epilogue groups of four warps: group `grp` drains sub-tile `grp` in 64-column halves through its own 16 KB staging tile
    // (TMEM -> registers -> (+bias) -> bf16 -> XOR-swizzled rows), then every thread re-reads ONE 16-byte chunk column of 16 rows:
    // fully coalesced NHWC stores (8 threads cover a pixel's 128 contiguous bytes) and, from the same registers, the BatchNorm
    // partial sums of the thread's 8 channels.  The head GEMM (out_mode 1) writes fp32 NCHW planes straight from registers.
    const int q = warp & 3, grp = warp >> 2;
    const int et = q * 32 + lane;                 // 0..127: epilogue thread id within the group == accumulator row (TMEM lane)
    const int g = et >> 3, j = et & 7;            // row = group g (image row h0+g), pixel j inside the group
    uint8_t* const stg = smem_al + (stg_base - smem_base) + grp * (128 * 128);
    const int lc = et & 7, pr = et >> 3;          // store pass: 16-byte chunk column, first row
    int as = 0; uint32_t aphase = 0;
    for (ItemIter<CL> it(p, rank); it.valid(); it.next()) {
      const int c = it.j / p.items_m, mi = it.j % p.items_m;
      const int sw = mi % p.st_w; const int r2 = mi / p.st_w;
      const int sh = r2 % p.st_h, n = r2 / p.st_h;
      const bool has_acc = p.cls[c].mask != 0;
      const int py = p.cls[c].py, px = p.cls[c].px;
      mbar_wait(&tfull[as], aphase);
      if (et == 1) HTL(30 + grp);
      tc_fence_after();
      const int sub = grp;
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * 2 + sub) * Ntile);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        uint32_t v[64];
        if (has_acc) {
          tmem_ld32(t_addr + h * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld32(t_addr + h * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) v[i] = 0u;
        }
        if (h == NH - 1) {            // this warp's share of the accumulator stage is out of TMEM: hand it back to the MMA warp (8 arrivals)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[as]);
          if (et == 1) HTL(32 + grp);
        }
        if (bias) {
          const float4* b4 = reinterpret_cast<const float4*>(bias + c0 + h * 64);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 bb = __ldg(b4 + i);
            v[4 * i] = __float_as_uint(__uint_as_float(v[4 * i]) + bb.x); v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + bb.y);
            v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + bb.z); v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + bb.w);
          }
        }
        if (p.out_mode == 1) {      // fp32 NCHW planes (prediction volume): straight from registers
          float* dst = reinterpret_cast<float*>(outp);
          const int ho = (sh * 16 + g) * p.out_s + py, wo = (sw * 16 + sub * 8 + j) * p.out_s + px;
          const size_t P = (size_t)p.Ho * p.Wo, pix = (size_t)ho * p.Wo + wo;
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const int cc = c0 + h * 64 + i;
            if (cc < p.n_valid) dst[((size_t)n * p.n_valid + cc) * P + pix] = __uint_as_float(v[i]);
          }
          continue;
        }
        // bf16 -> staging row `et` (128 B): 16-byte chunk index XOR (row & 7)
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8) {
          uint4 pk;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(v[8 * k8]), __uint_as_float(v[8 * k8 + 1]));
          __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(v[8 * k8 + 2]), __uint_as_float(v[8 * k8 + 3]));
          __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[8 * k8 + 4]), __uint_as_float(v[8 * k8 + 5]));
          __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(v[8 * k8 + 6]), __uint_as_float(v[8 * k8 + 7]));
          pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(stg + et * 128 + ((k8 ^ (et & 7)) << 4)) = pk;
        }
        epi_bar(grp);
        if (et == 1) HTL(34 + grp);
        // store pass: rows rr = pass*16 + pr, chunk column lc; four rows in flight
        bf16* const outb = reinterpret_cast<bf16*>(outp);
#pragma unroll
        for (int pass = 0; pass < 8; pass += 4) {
          uint4 val[4], old[4]; bf16* dst[4];
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int rr = (pass + k4) * 16 + pr;
            const int gg = rr >> 3, jj = rr & 7;
            const int ho2 = (sh * 16 + gg) * p.out_s + py, wo2 = (sw * 16 + sub * 8 + jj) * p.out_s + px;
            dst[k4] = outb + (((size_t)n * p.Ho + ho2) * p.Wo + wo2) * p.Cn + c0 + h * 64 + lc * 8;
            if (p.accumulate) old[k4] = *reinterpret_cast<const uint4*>(dst[k4]);
            val[k4] = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((lc ^ (rr & 7)) << 4));
          }
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&val[k4]);
            if (stats != nullptr) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 x = __bfloat1622float2(hv[i]);
                s1[h][2 * i] += x.x; s1[h][2 * i + 1] += x.y;
                s2[h][2 * i] = fmaf(x.x, x.x, s2[h][2 * i]); s2[h][2 * i + 1] = fmaf(x.y, x.y, s2[h][2 * i + 1]);
              }
            }
            if (!p.accumulate) {
              *reinterpret_cast<uint4*>(dst[k4]) = val[k4];
            } else {                // accumulate into the existing gradient (read-modify-write; the global loads were issued above)
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
        epi_bar(grp);                 // staging tile free for the next half / item
        if (et == 1) HTL(36 + grp);
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    // per-CTA statistics, fixed order: lanes sharing a chunk column (lane, lane^8, lane^16, lane^24) -> warp partial in the group's staging tile
    if (stats != nullptr) {
      float* part = reinterpret_cast<float*>(stg) + q * (NH * 128);     // the group's own staging tile (free after its last barrier): [warp][half][stat][64 ch]
#pragma unroll
      for (int h = 0; h < NH; ++h)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float a = s1[h][k], b = s2[h][k];
          a += __shfl_xor_sync(0xffffffffu, a, 8); b += __shfl_xor_sync(0xffffffffu, b, 8);
          a += __shfl_xor_sync(0xffffffffu, a, 16); b += __shfl_xor_sync(0xffffffffu, b, 16);
          if (lane < 8) { part[h * 128 + lane * 8 + k] = a; part[h * 128 + 64 + lane * 8 + k] = b; }
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL) cluster_sync_all();    // neither CTA may leave while its peer can still multicast into it or signal its barriers
  if (threadIdx.x == 0) HTL(3);
  if (stats != nullptr && threadIdx.x < NH * 128) {
    // thread -> (half, stat, channel): sum the 8 warp partials in fixed order, then one order-independent global accumulation
    const float* part = reinterpret_cast<const float*>(smem_al + (stg_base - smem_base));
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += part[(w >> 2) * (128 * 128 / 4) + (w & 3) * (NH * 128) + threadIdx.x];
    const int h = threadIdx.x >> 7, st = (threadIdx.x >> 6) & 1, ch = threadIdx.x & 63;
    acc_add(stats + st * p.Cn + c0 + h * 64 + ch, v);
  }
  if (warp == 9) { __syncwarp(); tmem_dealloc(tmem_base, kTmemCols); }
#ifdef AWR_CONV_PROFILE
  if (threadIdx.x == 0) HTL(4);
#endif
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

#ifdef AWR_CONV_PROFILE
extern "C" int awr_debug_halo_timeline(unsigned long long* out_host /*[769]: count, events (0 = unused slot)*/, int flags) {
  (void)flags;
  cudaDeviceSynchronize();
  out_host[0] = 6 * 128;
  cudaMemcpyFromSymbol(out_host + 1, g_halo_tl, sizeof(g_halo_tl));
  static unsigned long long z[6 * 128];
  cudaMemcpyToSymbol(g_halo_tl, z, sizeof(z));
  return 0;
}
#endif

// window extents of one tap class
static void class_extents(const ConvTapClass& t, int& mnx, int& mxx, int& mny, int& mxy) {
  mnx = mny = 99; mxx = mxy = -99;
  for (int i = 0; i < t.ntaps; ++i) { mnx = min(mnx, (int)t.ox[i]); mxx = max(mxx, (int)t.ox[i]); mny = min(mny, (int)t.oy[i]); mxy = max(mxy, (int)t.oy[i]); }
}

bool conv_halo_supported(const ConvGeom& g) {
  if (g.a_stride != 1 || g.Hc < 16 || g.Wc < 16 || (g.Hc % 16) || (g.Wc % 16)) return false;
  for (int c = 0; c < g.nclasses; ++c) {
    const ConvTapClass& t = g.cls[c];
    if (t.ntaps == 0) continue;
    int mnx, mxx, mny, mxy;
    class_extents(t, mnx, mxx, mny, mxy);
    if (mxx - mnx > 2 || mxy - mny > 2) return false;
  }
  return true;
}

template <int NH, int E, bool CL>
static cudaError_t halo_launch_t(int grid, size_t smem, cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias, void* out,
                                 AwrAcc* st, const HaloParams& p) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<NH, E, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (CL) return launch_pdl_cluster(conv_halo_kernel<NH, E, CL>, dim3(grid), dim3(kThreads), smem, stream, 2, tmA, tmB, bias, out, st, p);
  return launch_pdl(conv_halo_kernel<NH, E, CL>, dim3(grid), dim3(kThreads), smem, stream, tmA, tmB, bias, out, st, p);
}
template <int NH, bool CL>
static cudaError_t halo_launch_e(int E, int grid, size_t smem, cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias,
                                 void* out, AwrAcc* st, const HaloParams& p) {
  return E == 0 ? halo_launch_t<NH, 0, CL>(grid, smem, stream, tmA, tmB, bias, out, st, p)
       : E == 1 ? halo_launch_t<NH, 1, CL>(grid, smem, stream, tmA, tmB, bias, out, st, p)
                : halo_launch_t<NH, 2, CL>(grid, smem, stream, tmA, tmB, bias, out, st, p);
}

int conv_halo_launch(const ConvGeom& g, const void* in, const void* w, const float* bias, void* out, void* stats, cudaStream_t stream) {
  // tuning overrides, read once (experiments only): AWR_HALO_STAGES = 2|3, AWR_HALO_RESIDENT = 0|1, AWR_HALO_NTILE = 64|128
  static const int env_hs = env_int("AWR_HALO_STAGES", 0), env_res = env_int("AWR_HALO_RESIDENT", 1), env_nt = env_int("AWR_HALO_NTILE", 0),
                   env_pair = env_int("AWR_HALO_PAIR", 0);       // CTA pairs + TMA multicast of the weight tiles: built, tested, NOT faster here
                   // (profiles/r02_probes.md: L2 is not the limiter; 2.104 vs 2.088 ms per step), so off by default
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.N = g.N; p.Hc = g.Hc; p.Wc = g.Wc; p.st_w = g.Wc / 16; p.st_h = g.Hc / 16; p.items_m = g.N * p.st_w * p.st_h;
  p.kblocks = g.Ck / 64; p.b_mn = g.b_mn;
  p.Ho = g.Ho; p.Wo = g.Wo; p.Cn = g.Cn; p.out_s = g.out_s; p.out_mode = g.out_mode; p.n_valid = g.n_valid; p.accumulate = g.accumulate;
  p.nclasses = g.nclasses;
  // N tile <= 128 (two sub-tiles x two TMEM stages); prefer 128 unless that leaves most SMs without a work item
  p.Ntile = (g.Cn % 128 == 0 && (long long)g.nclasses * p.items_m * (g.Cn / 128) >= 120) ? 128 : 64;
  if (env_nt == 64 || (env_nt == 128 && g.Cn % 128 == 0)) p.Ntile = env_nt;
  p.tiles_c = g.Cn / p.Ntile;

  // ---- tap grid: E = largest extent over the classes; every class is laid out on the (E+1)^2 grid anchored at its smallest offsets ----
  int E = 0, max_taps = 0;
  for (int c = 0; c < g.nclasses; ++c) {
    const ConvTapClass& t = g.cls[c];
    if (t.ntaps == 0) continue;
    int mnx, mxx, mny, mxy;
    class_extents(t, mnx, mxx, mny, mxy);
    E = max(E, max(mxx - mnx, mxy - mny));
    max_taps = max(max_taps, t.ntaps);
  }
  const int GW = E + 1, PW = 16 + E;
  bool full_grid = true;
  for (int c = 0; c < g.nclasses; ++c) {
    const ConvTapClass& t = g.cls[c];
    HaloClass& h = p.cls[c];
    h.ntaps = t.ntaps; h.py = t.py; h.px = t.px;
    if (t.ntaps == 0) { full_grid = false; continue; }
    int mnx, mxx, mny, mxy;
    class_extents(t, mnx, mxx, mny, mxy);
    h.min_ox = mnx; h.min_oy = mny;
    for (int i = 0; i < t.ntaps; ++i) {
      const int gpos = ((int)t.oy[i] - mny) * GW + ((int)t.ox[i] - mnx);
      if (h.mask & (1 << gpos)) return AWR_ERR_UNSUPPORTED;              // two taps on one shift: not a convolution we build
      h.mask |= 1 << gpos; h.widx[gpos] = t.widx[i];
      h.last_g = max(h.last_g, gpos);
    }
    if (h.mask != (1 << (GW * GW)) - 1) full_grid = false;
  }
  p.halo_bytes = PW * PW * 128;
  CUtensorMap tmA, tmB;
  {
    const long long dims[4] = {g.Ck, g.Wi, g.Hi, g.N};
    const long long str[4] = {1, g.Ck, (long long)g.Wi * g.Ck, (long long)g.Hi * g.Wi * g.Ck};
    const int box[4] = {64, PW, PW, 1};
    if (!make_tmap_bf16(&tmA, in, 4, dims, str, box, nullptr)) return AWR_ERR_DRIVER;
  }
  if (!conv_make_weight_map(&tmB, g, w, p.Ntile)) return AWR_ERR_DRIVER;

  // ---- shared-memory budget: halo stages | weight ring in PAIR stages of two tiles (or the resident weight set) | 32 KB staging ---------
  const int b_bytes = p.Ntile * 128, pair_bytes = 2 * b_bytes;
  p.halo_stage_bytes = (p.halo_bytes + 1023) & ~1023;
  const int res_pairs = (max_taps + 1) / 2;
  p.resident = (env_res && g.nclasses == 1 && full_grid && p.kblocks == 1 && res_pairs <= kMaxBStages &&
                2 * p.halo_stage_bytes + res_pairs * pair_bytes + kStagingBytes + 1024 <= kSmemBudget) ? 1 : 0;
  if (p.resident) {
    p.b_stages = res_pairs;
    p.halo_stages = (3 * p.halo_stage_bytes + res_pairs * pair_bytes + kStagingBytes + 1024 <= kSmemBudget) ? 3 : 2;
  } else {
    // three halo stages when at least 3 pair stages still fit, else two
    p.halo_stages = 3;
    int bst = (kSmemBudget - 1024 - kStagingBytes - 3 * p.halo_stage_bytes) / pair_bytes;
    if (bst < 3) { p.halo_stages = 2; bst = (kSmemBudget - 1024 - kStagingBytes - 2 * p.halo_stage_bytes) / pair_bytes; }
    if (env_hs == 2 || env_hs == 3) { p.halo_stages = env_hs; bst = (kSmemBudget - 1024 - kStagingBytes - env_hs * p.halo_stage_bytes) / pair_bytes; }
    if (bst > 5) bst = 5;
    if (bst < 2) return AWR_ERR_UNSUPPORTED;
    p.b_stages = bst;
  }
  const size_t smem = (size_t)p.halo_stages * p.halo_stage_bytes + (size_t)p.b_stages * pair_bytes + kStagingBytes + 1024;
  // grid: at most one CTA per SM.  CTA pairs (clusters of 2 sharing every weight tile through TMA multicast) when the items pair up inside
  // a class; then the number of PAIRS is a multiple of tiles_c, else the number of CTAs (every CTA keeps one channel tile).
  const long long total = (long long)p.nclasses * p.tiles_c * p.items_m;
  const bool pair = env_pair && (p.items_m % 2 == 0) && total >= 2 * p.tiles_c;
  int grid;
  if (pair) {
    const int cap = awr_sm_budget() / 2;
    int ncl = (int)(total / 2 < cap ? total / 2 : cap);
    ncl -= ncl % p.tiles_c;
    grid = 2 * ncl;
  } else {
    grid = (int)(total < awr_sm_budget() ? total : awr_sm_budget());
    grid -= grid % p.tiles_c;
  }
  if (grid < p.tiles_c) return AWR_ERR_UNSUPPORTED;
  AwrAcc* st = reinterpret_cast<AwrAcc*>(stats);
  cudaError_t e;
  const int nh = p.Ntile / 64;
  if (pair) e = nh == 1 ? halo_launch_e<1, true>(E, grid, smem, stream, tmA, tmB, bias, out, st, p) : halo_launch_e<2, true>(E, grid, smem, stream, tmA, tmB, bias, out, st, p);
  else      e = nh == 1 ? halo_launch_e<1, false>(E, grid, smem, stream, tmA, tmB, bias, out, st, p) : halo_launch_e<2, false>(E, grid, smem, stream, tmA, tmB, bias, out, st, p);
  if (e != cudaSuccess) return (int)cudaGetLastError();
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}
