// Stem convolution (1 input channel, 5x5, pad 2, 64 output channels; model/resnet_deconv.py:32, model/hourglass.py:112) on tcgen05.
//
// K = 25 is not a tensor-core shape and the input is fp32, so the CUDA-core kernel (conv_simt.cu: stem_conv_tiled_kernel) spends ~1 500 issue
// slots per 4 pixels x 8 channels and runs at 70 us for the headline batch, six times its 67 MB output write.  Here the CTA builds the im2col
// operand itself: a tile is 128 consecutive pixels of one image row; thread r writes pixel r's 25 taps as ONE 128-byte K-major row of a
// SWIZZLE_128B operand tile -- k 0..24 = bf16(x) ("hi"), k 32..56 = bf16(x - hi) ("lo"), the rest zero -- exactly where a TMA load would have
// put it (16-byte chunk c of row r at chunk c ^ (r & 7)).  The weights are split the same way, so
//     D = A[hi|lo] * B0[w_hi|w_hi]^T  (K = 64, four MMAs)  +  A[hi] * B1[w_lo]^T  (K = 32, two MMAs)
// carries x*w to ~2^-16 relative (x_lo*w_lo is dropped): the result matches the fp32 CUDA-core kernel to output rounding.  Six M=128, N=64
// MMAs per tile are nothing; the kernel is bound by building A (~100 instructions per pixel) and by the output write.
//
// 128 threads, 3 CTAs per SM (67 KB smem, 128 TMEM columns each), persistent over tiles with two operand / accumulator stages: while the
// MMAs of tile t run, the CTA drains tile t-1: TMEM -> registers -> (+bias) -> bf16 -> XOR-swizzled per-warp staging -> coalesced 16-byte
// NHWC stores; the BatchNorm batch statistics (sum / sum of squares of the STORED bf16 values) are accumulated in registers during that store
// pass (a thread owns the same 8 channels for the CTA's lifetime), reduced once per CTA in a fixed order and added to the order-independent
// global accumulators (AwrAcc).
#include <cuda_bf16.h>

#include "tc_common.cuh"
#include "common.cuh"
#include "awr_b200.h"

namespace {

using namespace tc;

constexpr int kStemThreads = 128;
constexpr int kTile = 128;                       // pixels per tile = GEMM rows
constexpr int kCout = 64;
constexpr int kABytes = kTile * 128;             // 128 rows x 64 bf16
constexpr int kBBytes = kCout * 128;             // 64 rows x 64 bf16
constexpr int kPatchW = kTile + 4;               // 5 x 132 fp32 input patch
constexpr int kSmemBytes = 2 * kABytes + 2 * kBBytes + 4 * 32 * 128 + 2 * 5 * kPatchW * 4 + 1024;

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

__global__ void __launch_bounds__(kStemThreads, 3)
stem_conv_tc_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                    AwrAcc* __restrict__ stats, int N, int H, int W) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  __shared__ __align__(8) uint64_t full_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[kCout];
  __shared__ float s_red[4][8][16];              // [warp][channel octet][sum 0..7 | sumsq 0..7]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* base = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  uint8_t* sA = base;                            // 2 stages
  uint8_t* sB = base + 2 * kABytes;              // B0 (w_hi | w_hi), B1 (w_lo | 0)
  uint8_t* sStage = sB + 2 * kBBytes;            // 4 warps x 32 rows x 128 B
  float* sPatch = reinterpret_cast<float*>(sStage + 4 * 32 * 128);

  if (tid == 0) {
    mbar_init(&full_bar[0], 1); mbar_init(&full_bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();                                    // the optimizer of the previous step wrote `w`; earlier readers of `y` are done

  // ---- weights: w is [25][64] fp32 (tap-major).  Row c of B0 / B1 = output channel c, K-major, SWIZZLE_128B ---------------------------
  if (tid < kCout) {
    const int c = tid;
    float wv[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) wv[t] = t < 25 ? w[t * kCout + c] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a = wv[8 * j + 2 * q], b = wv[8 * j + 2 * q + 1];
        const float ah = __bfloat162float(__float2bfloat16_rn(a)), bh = __bfloat162float(__float2bfloat16_rn(b));
        hi[q] = pack_bf16(a, b);
        lo[q] = pack_bf16(a - ah, b - bh);
      }
      const uint4 h4 = make_uint4(hi[0], hi[1], hi[2], hi[3]), l4 = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(sB + c * 128 + ((j ^ (c & 7)) << 4)) = h4;
      *reinterpret_cast<uint4*>(sB + c * 128 + (((4 + j) ^ (c & 7)) << 4)) = h4;
      *reinterpret_cast<uint4*>(sB + kBBytes + c * 128 + ((j ^ (c & 7)) << 4)) = l4;
      *reinterpret_cast<uint4*>(sB + kBBytes + c * 128 + (((4 + j) ^ (c & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    }
    s_bias[c] = bias ? bias[c] : 0.f;
  }

  const int tpr = W / kTile;                     // tiles per image row
  const int total = N * H * tpr;
  const uint32_t idesc = umma_idesc_bf16(128, kCout, 0, 0);
  const uint64_t b0desc = umma_desc_sw128(smem_u32(sB), 16, 1024), b1desc = umma_desc_sw128(smem_u32(sB) + kBBytes, 16, 1024);
  float s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
  uint8_t* myStage = sStage + warp * 32 * 128;
  const int cc = lane & 7;                       // the channel octet this thread stores and accumulates statistics for

  // drain accumulator stage `s` of tile `tile`: TMEM -> bf16 staging -> coalesced stores + statistics
  auto epilogue = [&](int tile, int s, uint32_t parity) {
    mbar_wait(&full_bar[s], parity);
    tc_fence_after();
    uint32_t v0[32], v1[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * kCout);
    tmem_ld32(taddr, v0);
    tmem_ld32(taddr + 32, v1);
    tmem_ld_wait();
    tc_fence_before();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t pk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = 8 * j + 2 * q;
        const float a = __uint_as_float(c < 32 ? v0[c] : v1[c - 32]) + s_bias[c];
        const float b = __uint_as_float(c + 1 < 32 ? v0[c + 1] : v1[c + 1 - 32]) + s_bias[c + 1];
        pk[q] = pack_bf16(a, b);
      }
      *reinterpret_cast<uint4*>(myStage + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    __syncwarp();
    const int xb = tile % tpr, row = tile / tpr;                         // row = n * H + y
    __nv_bfloat16* dst = y + ((size_t)row * W + (size_t)xb * kTile + warp * 32) * kCout;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = (lane >> 3) + 4 * i;
      const uint4 q4 = *reinterpret_cast<const uint4*>(myStage + rr * 128 + ((cc ^ (rr & 7)) << 4));
      *reinterpret_cast<uint4*>(dst + (size_t)rr * kCout + cc * 8) = q4;
      const uint32_t u[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[q]));
        s1[2 * q] += f.x; s2[2 * q] = fmaf(f.x, f.x, s2[2 * q]);
        s1[2 * q + 1] += f.y; s2[2 * q + 1] = fmaf(f.y, f.y, s2[2 * q + 1]);
      }
    }
    __syncwarp();                                // the staging slice is free for the next tile
  };

  // input patch of a tile: 5 x 132 fp32 (zero outside the image), element i = tid + 128 j; fetched into registers one tile ahead so
  // that the global-load latency hides behind the A build and the epilogue of the tiles in flight
  constexpr int kPatchPerThread = (5 * kPatchW + kStemThreads - 1) / kStemThreads;      // 6
  int p_r[kPatchPerThread], p_c[kPatchPerThread];
#pragma unroll
  for (int j = 0; j < kPatchPerThread; ++j) { const int i = tid + j * kStemThreads; p_r[j] = i / kPatchW - 2; p_c[j] = i % kPatchW - 2; }
  float pv[kPatchPerThread];
  auto fetch_patch = [&](int tile) {
    const int xb = tile % tpr; const int r2 = tile / tpr;
    const int yy = r2 % H, n = r2 / H;
    const float* xi = x + (size_t)n * H * W;
#pragma unroll
    for (int j = 0; j < kPatchPerThread; ++j) {
      const int hh = yy + p_r[j], ww = xb * kTile + p_c[j];
      pv[j] = (hh >= 0 && hh < H && ww >= 0 && ww < W && p_r[j] < 3) ? __ldg(xi + (size_t)hh * W + ww) : 0.f;
    }
  };
  auto store_patch = [&](int buf) {
#pragma unroll
    for (int j = 0; j < kPatchPerThread; ++j) {
      const int i = tid + j * kStemThreads;
      if (i < 5 * kPatchW) sPatch[buf * 5 * kPatchW + i] = pv[j];
    }
  };
  int it = 0, prev_tile = -1;
  if ((int)blockIdx.x < total) { fetch_patch(blockIdx.x); store_patch(0); }
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
    const int s = it & 1;
    __syncthreads();                             // patch buffer s is complete; every warp is past the previous iteration's epilogue
    const int next = tile + gridDim.x;
    if (next < total) fetch_patch(next);
    // ---- im2col row of pixel `tid`: hi | lo, K-major, swizzled ---------------------------------------------------------------------
    {
      const float* pt = sPatch + s * 5 * kPatchW;
      float tv[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) tv[t] = t < 25 ? pt[(t / 5) * kPatchW + tid + (t % 5)] : 0.f;
      uint8_t* arow = sA + s * kABytes + tid * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float a = tv[8 * j + 2 * q], b = tv[8 * j + 2 * q + 1];
          const float ah = __bfloat162float(__float2bfloat16_rn(a)), bh = __bfloat162float(__float2bfloat16_rn(b));
          hi[q] = pack_bf16(a, b);
          lo[q] = pack_bf16(a - ah, b - bh);
        }
        *reinterpret_cast<uint4*>(arow + ((j ^ (tid & 7)) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(arow + (((4 + j) ^ (tid & 7)) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
    fence_proxy_async();                         // generic-proxy writes of A (and, first time, B) -> visible to the tensor core's async proxy
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      const uint64_t ad = umma_desc_sw128(smem_u32(sA) + s * kABytes, 16, 1024);
      const uint32_t d = tmem_base + (uint32_t)(s * kCout);
      umma_bf16(d, ad, b0desc, idesc, 0u);
      umma_bf16(d, ad + 2, b0desc + 2, idesc, 1u);
      umma_bf16(d, ad + 4, b0desc + 4, idesc, 1u);
      umma_bf16(d, ad + 6, b0desc + 6, idesc, 1u);
      umma_bf16(d, ad, b1desc, idesc, 1u);
      umma_bf16(d, ad + 2, b1desc + 2, idesc, 1u);
      umma_commit(&full_bar[s]);
    }
    __syncwarp();
    // ---- while those run: drain the previous tile (stage s^1).  Its A stage and TMEM stage are reused at iteration it+1, after the
    //      __syncthreads at the top of the loop, i.e. after every warp has finished this epilogue.
    if (prev_tile >= 0) epilogue(prev_tile, s ^ 1, (uint32_t)(((it - 1) >> 1) & 1));
    prev_tile = tile;
    if (next < total) store_patch(s ^ 1);        // buffer s^1 was last read before this iteration's second __syncthreads
  }
  if (prev_tile >= 0) epilogue(prev_tile, (it - 1) & 1, (uint32_t)(((it - 1) >> 1) & 1));

  // ---- statistics: lanes with the same channel octet (lane & 7) first, then the four warps in order ------------------------------------
  if (stats) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 8); s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 16);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], 8); s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], 16);
    }
    if (lane < 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { s_red[warp][lane][k] = s1[k]; s_red[warp][lane][8 + k] = s2[k]; }
    }
    __syncthreads();
    {                                            // thread t: t < 64 -> sum of channel t; else sum of squares of channel t - 64
      const int c = tid & 63, which = tid >> 6;
      const float tot = ((s_red[0][c >> 3][which * 8 + (c & 7)] + s_red[1][c >> 3][which * 8 + (c & 7)]) + s_red[2][c >> 3][which * 8 + (c & 7)]) +
                        s_red[3][c >> 3][which * 8 + (c & 7)];
      acc_add(stats + which * kCout + c, tot);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, 128); }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Weight gradient of the stem on tcgen05:  dW[tap][c] = sum over pixels of patch[pixel][tap] * dy[pixel][c]  (+ dbias[c] = sum of dy).
// The contraction runs over PIXELS, so both operands are MN-major: the hand-built im2col tile [pixel][hi | lo taps] of the forward kernel,
// read through an MN-major descriptor, is A^T (M = 64 tap rows, padded to M = 128 by a second, all-zero 64-row atom reached through the
// descriptor's leading-dimension offset), and dy's NHWC tile [pixel][64 channels] arrives by TMA as B.  A constant 1.0 in tap slot 25 makes
// row 25 of D the bias gradient.  One accumulator D[128 x 64] in TMEM per CTA for its whole pixel range; the epilogue adds the hi and lo
// rows of each tap and issues one atomic per (tap, channel) and CTA.
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int kWgSmemBytes = 2 * kABytes + kABytes + 2 * kABytes + 2 * 5 * kPatchW * 4 + 1024;      // A x2, zero atom, dy x2, patches

__global__ void __launch_bounds__(kStemThreads, 2)
stem_wgrad_tc_kernel(const float* __restrict__ x, const __grid_constant__ CUtensorMap tmDy, GradT* __restrict__ dW, GradT* __restrict__ dbias,
                     int N, int H, int W) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2], done_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* base = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  uint8_t* sA = base;                            // 2 stages x [128 pixels][hi | lo taps]
  uint8_t* sZ = base + 2 * kABytes;              // the all-zero second M atom
  uint8_t* sDy = sZ + kABytes;                   // 2 stages x [128 pixels][64 channels] (TMA, SWIZZLE_128B)
  float* sPatch = reinterpret_cast<float*>(sDy + 2 * kABytes);

  if (tid == 0) {
    tma_prefetch_desc(&tmDy);
    for (int i = 0; i < 2; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < kABytes / 16; i += kStemThreads) reinterpret_cast<uint4*>(sZ)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();

  const int tpr = W / kTile;
  const int total = N * H * tpr;
  // contiguous tile range per CTA
  const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t0 = (int)blockIdx.x * per, t1 = min(total, t0 + per);
  const uint32_t idesc = umma_idesc_bf16(128, kCout, 1, 1);

  constexpr int kPatchPerThread = (5 * kPatchW + kStemThreads - 1) / kStemThreads;
  int p_r[kPatchPerThread], p_c[kPatchPerThread];
#pragma unroll
  for (int j = 0; j < kPatchPerThread; ++j) { const int i = tid + j * kStemThreads; p_r[j] = i / kPatchW - 2; p_c[j] = i % kPatchW - 2; }
  float pv[kPatchPerThread];
  auto fetch_patch = [&](int tile) {
    const int xb = tile % tpr; const int r2 = tile / tpr;
    const int yy = r2 % H, n = r2 / H;
    const float* xi = x + (size_t)n * H * W;
#pragma unroll
    for (int j = 0; j < kPatchPerThread; ++j) {
      const int hh = yy + p_r[j], ww = xb * kTile + p_c[j];
      pv[j] = (hh >= 0 && hh < H && ww >= 0 && ww < W && p_r[j] < 3) ? __ldg(xi + (size_t)hh * W + ww) : 0.f;
    }
  };
  auto store_patch = [&](int buf) {
#pragma unroll
    for (int j = 0; j < kPatchPerThread; ++j) {
      const int i = tid + j * kStemThreads;
      if (i < 5 * kPatchW) sPatch[buf * 5 * kPatchW + i] = pv[j];
    }
  };

  if (t0 < t1) {
    fetch_patch(t0); store_patch(0);
    if (tid == 0) {
      mbar_expect_tx(&full_bar[0], (uint32_t)kABytes);
      tma_load_3d(sDy, &tmDy, &full_bar[0], 0, 0, t0);
    }
  }
  int it = 0;
  for (int tile = t0; tile < t1; ++tile, ++it) {
    const int s = it & 1;
    if (it >= 2) mbar_wait(&empty_bar[s], (uint32_t)(((it >> 1) - 1) & 1));    // the MMAs that read A[s] / dy[s] two tiles ago are done
    __syncthreads();                             // patch buffer s complete
    const int next = tile + 1;
    if (next < t1) fetch_patch(next);
    {
      const float* pt = sPatch + s * 5 * kPatchW;
      float tv[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) tv[t] = t < 25 ? pt[(t / 5) * kPatchW + tid + (t % 5)] : (t == 25 ? 1.f : 0.f);       // slot 25: the bias row
      uint8_t* arow = sA + s * kABytes + tid * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float a = tv[8 * j + 2 * q], b = tv[8 * j + 2 * q + 1];
          const float ah = __bfloat162float(__float2bfloat16_rn(a)), bh = __bfloat162float(__float2bfloat16_rn(b));
          hi[q] = pack_bf16(a, b);
          lo[q] = pack_bf16(a - ah, b - bh);
        }
        *reinterpret_cast<uint4*>(arow + ((j ^ (tid & 7)) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(arow + (((4 + j) ^ (tid & 7)) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        mbar_wait(&full_bar[s], (uint32_t)((it >> 1) & 1));            // dy tile landed
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA) + s * kABytes;
        const uint32_t lbo = (smem_u32(sZ)) - a0;                       // second M atom: the zero block
        const uint64_t ad = umma_desc_sw128(a0, lbo, 1024);
        const uint64_t bd = umma_desc_sw128(smem_u32(sDy) + s * kABytes, 8192, 1024);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tmem_base, ad + (uint64_t)(k * 128), bd + (uint64_t)(k * 128), idesc, (it | k) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
        if (next < t1) {                                                // dy of the next tile into the other stage (its last readers: tile it-1)
          if (it >= 1) mbar_wait(&empty_bar[s ^ 1], (uint32_t)((((it - 1) >> 1)) & 1));
          mbar_expect_tx(&full_bar[s ^ 1], (uint32_t)kABytes);
          tma_load_3d(sDy + (s ^ 1) * kABytes, &tmDy, &full_bar[s ^ 1], 0, 0, next);
        }
      }
      __syncwarp();
    }
    if (next < t1) store_patch(s ^ 1);
  }
  if (warp == 0 && elect_one()) {
    if (it > 0) umma_commit(&done_bar); else mbar_arrive(&done_bar);
  }
  __syncwarp();
  mbar_wait(&done_bar, 0u);
  tc_fence_after();

  // ---- epilogue: D rows 0..25 (hi taps, bias row) sit in warp 0's TMEM lanes, rows 32..56 (lo taps) in warp 1's -------------------------
  float* sHi = reinterpret_cast<float*>(sA);     // [32][64] fp32 each (the operand stages are idle now)
  float* sLo = sHi + 32 * 64;
  if (warp < 2 && it > 0) {
    uint32_t v0[32], v1[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    tmem_ld32(taddr, v0);
    tmem_ld32(taddr + 32, v1);
    tmem_ld_wait();
    float* dst = (warp == 0 ? sHi : sLo) + (tid & 31) * 64;
#pragma unroll
    for (int c = 0; c < 32; ++c) { dst[c] = __uint_as_float(v0[c]); dst[32 + c] = __uint_as_float(v1[c]); }
  }
  tc_fence_before();
  __syncthreads();
  if (it > 0) {
    for (int i = tid; i < 25 * kCout; i += kStemThreads) grad_add(dW + i, sHi[i] + sLo[i]);
    if (dbias && tid < kCout) grad_add(dbias + tid, sHi[25 * kCout + tid] + sLo[25 * kCout + tid]);
  }
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, 64); }
}

}  // namespace

// bf16 NHWC output, Cout = 64, k = 5, W a multiple of 128.  Returns AWR_ERR_UNSUPPORTED otherwise (the caller falls back to the CUDA-core kernel).
int stem_conv_tc_launch(const float* x, const float* w, const float* bias, void* y, void* stats, int N, int H, int W, int Cout, int k, cudaStream_t st) {
  if (!(Cout == kCout && k == 5 && W % kTile == 0 && N > 0 && H > 0)) return AWR_ERR_UNSUPPORTED;
  static const bool configured = [] {
    return cudaFuncSetAttribute(stem_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) == cudaSuccess;
  }();
  if (!configured) return AWR_ERR_DRIVER;
  const int total = N * H * (W / kTile);
  const int grid = total < 3 * awr_sm_budget() ? total : 3 * awr_sm_budget();
  launch_pdl(stem_conv_tc_kernel, dim3(grid), dim3(kStemThreads), (size_t)kSmemBytes, st, x, w, bias, (__nv_bfloat16*)y, (AwrAcc*)stats, N, H, W);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

// dy bf16 NHWC (N,H,W,64), k = 5, W a multiple of 128; dW [25][64] and dbias [64] (optional) are accumulated into (GradT: fp32, or the
// accumulator slots of the bit-reproducible build).  Returns AWR_ERR_UNSUPPORTED for other shapes.
int stem_wgrad_tc_launch(const float* x, const void* dy, void* dW, void* dbias, int N, int H, int W, int Cout, int k, cudaStream_t st) {
  if (!(Cout == kCout && k == 5 && W % kTile == 0 && N > 0 && H > 0)) return AWR_ERR_UNSUPPORTED;
  static const bool configured = [] {
    return cudaFuncSetAttribute(stem_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes) == cudaSuccess;
  }();
  if (!configured) return AWR_ERR_DRIVER;
  const int total = N * H * (W / kTile);
  CUtensorMap tm;
  {
    const long long dims[3] = {kCout, kTile, total};
    const long long str[3] = {1, kCout, (long long)kTile * kCout};
    const int box[3] = {kCout, kTile, 1};
    if (!make_tmap_bf16(&tm, dy, 3, dims, str, box, nullptr)) return AWR_ERR_DRIVER;
  }
  const int grid = total < 2 * awr_sm_budget() ? total : 2 * awr_sm_budget();
  launch_pdl(stem_wgrad_tc_kernel, dim3(grid), dim3(kStemThreads), (size_t)kWgSmemBytes, st, x, tm, reinterpret_cast<GradT*>(dW),
             reinterpret_cast<GradT*>(dbias), N, H, W);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}
