// Hardware probe (debug): can a SWIZZLE_128B K-major UMMA operand be a ROW-SHIFTED window of a larger TMA-written smem tile?
//
// G (rows x 64 bf16) is loaded by one TMA box into smem (SWIZZLE_128B).  A = the 128 rows selected by the descriptor
//   row(m) = r0 + (m / 8) * (sbo_rows) + (m % 8),   start address = tile + r0 * 128 B, SBO = sbo_rows * 128 B,
// B = Bm (64 x 64, K-major).  D[128 x 64] = A * Bm^T is written as fp32.  base_mode: 0 -> descriptor base_offset 0,
// 1 -> base_offset = (start_address >> 7) & 7 (PTX matrix-descriptor rule for non-1024-B-aligned starts).
// This decides whether conv taps can be fed from ONE halo tile (design note in DESIGN.md, "operand reuse").
#include "../tc_common.cuh"
#include "awr_b200_debug.h"

namespace {
using namespace tc;

__global__ void __launch_bounds__(128, 1)
umma_window_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmB, float* __restrict__ D, int rows, int r0,
                   int sbo_rows, int base_mode) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sg = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sb = sg + rows * 128;
  if (threadIdx.x == 0) { mbar_init(&full_bar, 1); mbar_init(&done_bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full_bar, (uint32_t)(rows * 128 + 64 * 128));
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(sg)),
                 "l"(reinterpret_cast<uint64_t>(&tmG)), "r"(smem_u32(&full_bar)), "r"(0), "r"(0)
                 : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(sb)),
                 "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(&full_bar)), "r"(0), "r"(0)
                 : "memory");
    mbar_wait(&full_bar, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    const uint32_t a_addr = base + (uint32_t)r0 * 128u;
    uint64_t ad = umma_desc_sw128(a_addr, 16, (uint32_t)sbo_rows * 128u);
    if (base_mode == 1) ad |= (uint64_t)((a_addr >> 7) & 7u) << 49;
    const uint64_t bd = umma_desc_sw128(base + (uint32_t)rows * 128u, 16, 1024);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, k ? 1u : 0u);
    umma_commit(&done_bar);
  }
  __syncthreads();
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int ch = 0; ch < 64; ch += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ch, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[row * 64 + ch + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem, 64); }
}

// MMA-rate probe: one CTA per SM issues `iters` x 4 back-to-back tcgen05.mma (M=128, N, K=16) from fixed smem operands, cycling over
// `nacc` independent accumulators; no TMA traffic, no epilogue.  out[blockIdx] = cycles from first issue to completion of the last.
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(unsigned long long* out, int N, int nacc, int iters, int a_rows_shift) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t* z = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < (48 * 1024) / 4; i += 128) z[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(&done_bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint64_t ad = umma_desc_sw128(base + a_rows_shift * 128, 16, 1024), bd = umma_desc_sw128(base + 16384, 16, 1024);
    long long t0 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tmem + (uint32_t)((it % nacc) * N);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, 1u);
      }
      umma_commit(&done_bar);
    }
    __syncwarp();
    mbar_wait(&done_bar, 0);
    if (elect_one()) out[blockIdx.x] = (unsigned long long)(clock64() - t0);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem, 512); }
}


// MMA-rate probe with the operand pattern of the halo-tile convolution kernel: per "tap" one B tile feeds 8 MMAs over two accumulators
// (sub-tile 1 = 1024 B to the right of sub-tile 0), the A window starts `a_start` bytes into the tile (any multiple of 128) and its
// 8-row groups are `sbo` bytes apart (the halo row pitch; 1024 = dense canonical tile).  out[blockIdx] = cycles for iters*8 MMAs.
__global__ void __launch_bounds__(128, 1) umma_rate2_kernel(unsigned long long* out, int N, int sbo, int a_start, int tap_step, int iters) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t* z = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += 128) z[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(&done_bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint64_t ad0 = umma_desc_sw128(base + (uint32_t)a_start, 16, (uint32_t)sbo), bd = umma_desc_sw128(base + 64 * 1024, 16, 1024);
    long long t0 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint64_t a0 = ad0 + (uint64_t)(((uint32_t)(it % 3) * (uint32_t)tap_step) >> 4), a1 = a0 + 64;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_bf16(tmem, a0 + 2 * k, bd + 2 * k, idesc, 1u);
          umma_bf16(tmem + (uint32_t)N, a1 + 2 * k, bd + 2 * k, idesc, 1u);
        }
      }
      const long long t1 = clock64();                    // issue done (the MMAs may still be queued / executing)
      umma_commit(&done_bar);
      out[gridDim.x + blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    __syncwarp();
    mbar_wait(&done_bar, 0);
    if (elect_one()) out[blockIdx.x] = (unsigned long long)(clock64() - t0);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem, 512); }
}

// TMA throughput probe: every CTA streams `groups` boxes of `rows` x 128 B (2-D map over G[g_rows][64] bf16, SWIZZLE_128B) through a ring of
// `stages` buffers; a consumer thread only waits for each box and hands the buffer back.  cta_stride_rows = 0: all CTAs fetch the SAME
// rows at the same time (how the conv kernels fetch weight tiles); > 0: CTA b starts b*cta_stride_rows rows further (private data).
// out[blockIdx] = cycles from the first issue to the arrival of the last box.
// mode 0: one producer thread; 1: two producer threads (warps 0 and 2) issue alternate boxes; 2: alternate between two tensor maps;
// 3: 1-D bulk copies (cp.async.bulk, no tensor map) of rows*128 contiguous bytes
__global__ void __launch_bounds__(96, 1)
tma_rate_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmG2, const uint8_t* __restrict__ Graw, unsigned long long* out,
                int g_rows, int rows, int stages, int groups, int cta_stride_rows, int span_rows, int mode) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int stage_bytes = (rows * 128 + 1023) & ~1023;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int row0 = (int)(((long long)blockIdx.x * cta_stride_rows) % g_rows);
  if ((warp == 0 || (warp == 2 && mode == 1)) && lane == 0) {
    int st = 0; uint32_t ph = 0;
    for (int g = 0; g < groups; ++g) {
      if (mode != 1 || (g & 1) == (warp >> 1)) {
        mbar_wait(&empty_bar[st], ph ^ 1u);
        mbar_expect_tx(&full_bar[st], (uint32_t)(rows * 128));
        const int r = (row0 + (g % span_rows) * rows) % (g_rows - rows);
        if (mode == 3) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + st * stage_bytes)),
                       "l"(reinterpret_cast<uint64_t>(Graw + (size_t)r * 128)), "r"((uint32_t)(rows * 128)), "r"(smem_u32(&full_bar[st]))
                       : "memory");
        } else {
          const CUtensorMap* tm = (mode == 2 && (g & 1)) ? &tmG2 : &tmG;
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                           smem_u32(sm + st * stage_bytes)),
                       "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(&full_bar[st])), "r"(0), "r"(r)
                       : "memory");
        }
      }
      if (++st == stages) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 1 && lane == 0) {
    const long long t0 = clock64();
    int st = 0; uint32_t ph = 0;
    for (int g = 0; g < groups; ++g) {
      mbar_wait(&full_bar[st], ph);
      mbar_arrive(&empty_bar[st]);
      if (++st == stages) { st = 0; ph ^= 1u; }
    }
    out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
}

// Pipeline probe: the MMA-issuer loop of the conv kernels without epilogue.  mode bit0: tcgen05.commit to an mbarrier per group of
// `per` MMAs; bit1: mbarrier wait (on a barrier completed by that commit chain = "empty->full" ping) per group; bit2: a producer
// warp refills a `stages`-deep ring with 2-D TMA boxes of `tma_rows` x 128 B per group and the issuer waits for them (full pipeline).
__global__ void __launch_bounds__(128, 1)
umma_pipe_kernel(const __grid_constant__ CUtensorMap tmG, unsigned long long* out, int N, int per, int groups, int mode, int stages,
                 int tma_rows) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], done_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int stage_bytes = 32768;
  for (int i = threadIdx.x; i < (stages * stage_bytes) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 2 && (mode & 4)) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int g = 0; g < groups; ++g) {
        mbar_wait(&empty_bar[st], ph ^ 1u);
        mbar_expect_tx(&full_bar[st], (uint32_t)(tma_rows * 128));
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                         smem_u32(sm + st * stage_bytes)),
                     "l"(reinterpret_cast<uint64_t>(&tmG)), "r"(smem_u32(&full_bar[st])), "r"(0), "r"((g * tma_rows) & 0xFFFF)
                     : "memory");
        if (++st == stages) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint64_t ad0 = umma_desc_sw128(base, 16, 1024), bd0 = umma_desc_sw128(base + 16384, 16, 1024);
    long long t0 = clock64();
    int st = 0; uint32_t ph = 0;
    for (int g = 0; g < groups; ++g) {
      if (mode & 4) { mbar_wait(&full_bar[st], ph); tc_fence_after(); }
      else if ((mode & 2) && g >= stages) { mbar_wait(&empty_bar[st], ph ^ 1u); tc_fence_after(); }   // completed by the commit `stages` groups ago
      if (elect_one()) {
        const uint64_t ad = ad0 + (uint64_t)((uint32_t)(st * stage_bytes) >> 4), bd = bd0 + (uint64_t)((uint32_t)(st * stage_bytes) >> 4);
        if (per == 8) {
#pragma unroll
          for (int m = 0; m < 8; ++m) umma_bf16(tmem + (uint32_t)((m & 1) * N), ad + 2 * (m & 3), bd + 2 * (m & 3), idesc, 1u);
        } else {
#pragma unroll
          for (int m = 0; m < 4; ++m) umma_bf16(tmem + (uint32_t)((m & 1) * N), ad + 2 * (m & 3), bd + 2 * (m & 3), idesc, 1u);
        }
        if (mode & 1) umma_commit(&empty_bar[st]);
      }
      __syncwarp();
      if (++st == stages) { st = 0; ph ^= 1u; }
    }
    if (elect_one()) umma_commit(&done_bar);
    __syncwarp();
    mbar_wait(&done_bar, 0);
    if (lane == 0) out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem, 512); }
}

}  // namespace

extern "C" int awr_debug_umma_pipe(const void* G, int g_rows, unsigned long long* out_dev, int N, int per, int groups, int mode, int stages,
                                   int tma_rows, int grid, void* stream) {
  AWR_HOST_CHECK(G && out_dev && N >= 16 && 2 * N <= 512 && per > 0 && groups > 0 && stages >= 1 && stages <= 6 && tma_rows <= 256 && tma_rows > 0);
  CUtensorMap tmG;
  const long long dims[2] = {64, g_rows}, str[2] = {1, 64};
  const int box[2] = {64, tma_rows};
  if (!make_tmap_bf16(&tmG, G, 2, dims, str, box, nullptr)) return AWR_ERR_DRIVER;
  const size_t smem = (size_t)stages * 32768 + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  umma_pipe_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(tmG, out_dev, N, per, groups, mode, stages, tma_rows);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

extern "C" int awr_debug_umma_rate(unsigned long long* out_dev, int N, int nacc, int iters, int a_rows_shift, int grid, void* stream) {
  AWR_HOST_CHECK(out_dev && N >= 16 && N <= 256 && nacc >= 1 && nacc * N <= 512 && iters > 0 && grid > 0);
  const size_t smem = 49 * 1024 + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  umma_rate_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(out_dev, N, nacc, iters, a_rows_shift);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

extern "C" int awr_debug_umma_rate2(unsigned long long* out_dev, int N, int sbo, int a_start, int tap_step, int iters, int grid, void* stream) {
  AWR_HOST_CHECK(out_dev && N >= 16 && 2 * N <= 512 && sbo >= 1024 && sbo % 128 == 0 && a_start % 128 == 0 && tap_step % 128 == 0 && iters > 0 && grid > 0);
  AWR_HOST_CHECK(a_start + 2 * tap_step + 15 * sbo + 2048 <= 64 * 1024);
  const size_t smem = 97 * 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  umma_rate2_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(out_dev, N, sbo, a_start, tap_step, iters);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

extern "C" int awr_debug_tma_rate(const void* G, int g_rows, unsigned long long* out_dev, int rows, int stages, int groups, int cta_stride_rows,
                                  int span_rows, int mode, int grid, void* stream) {
  AWR_HOST_CHECK(G && out_dev && rows > 0 && rows <= 256 && stages >= 1 && stages <= 8 && groups > 0 && g_rows > rows && span_rows > 0 && grid > 0);
  CUtensorMap tmG;
  const long long dims[2] = {64, g_rows}, str[2] = {1, 64};
  const int box[2] = {64, rows};
  if (!make_tmap_bf16(&tmG, G, 2, dims, str, box, nullptr)) return AWR_ERR_DRIVER;
  CUtensorMap tmG2;
  const long long dims2[2] = {64, g_rows - 8};
  if (!make_tmap_bf16(&tmG2, G, 2, dims2, str, box, nullptr)) return AWR_ERR_DRIVER;
  const size_t smem = (size_t)stages * ((rows * 128 + 1023) & ~1023) + 1024;
  cudaError_t e = cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  tma_rate_kernel<<<grid, 96, smem, (cudaStream_t)stream>>>(tmG, tmG2, (const uint8_t*)G, out_dev, g_rows, rows, stages, groups, cta_stride_rows, span_rows, mode);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

extern "C" int awr_debug_umma_window(const void* G, const void* Bm, float* D, int rows, int r0, int sbo_rows, int base_mode, void* stream) {
  AWR_HOST_CHECK(G && Bm && D && rows >= 128 && rows <= 256 && r0 >= 0 && sbo_rows >= 8);
  AWR_HOST_CHECK(r0 + 15 * sbo_rows + 8 <= rows);
  CUtensorMap tmG, tmB;
  {
    const long long dims[2] = {64, rows}, str[2] = {1, 64};
    const int box[2] = {64, rows};
    if (!make_tmap_bf16(&tmG, G, 2, dims, str, box, nullptr)) return AWR_ERR_DRIVER;
  }
  {
    const long long dims[2] = {64, 64}, str[2] = {1, 64};
    const int box[2] = {64, 64};
    if (!make_tmap_bf16(&tmB, Bm, 2, dims, str, box, nullptr)) return AWR_ERR_DRIVER;
  }
  const size_t smem = (size_t)rows * 128 + 64 * 128 + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  umma_window_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmG, tmB, D, rows, r0, sbo_rows, base_mode);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}
