// Weight-gradient GEMM of Conv2d / ConvTranspose2d on tcgen05 tensor cores (sm_100a), NHWC bf16 activations, fp32 dW.
//
//   dW[tap][i*s_p + j*s_g] += sum over coarse pixels q of  pointwise[q, i] * gathered[q*stride - pad + tap, j]
//
// as a GEMM whose contraction dimension is the PIXEL index:  D[m = gathered channel j (of tap t)][n = pointwise channel i].
// Both operands are read straight from the NHWC tensors by TMA boxes of (64 channels x 64 pixels) and consumed as
// MN-major UMMA operands (the channel dimension is contiguous in memory, the pixel dimension strides by 128 B in smem),
// so no transposed copy of the activations is ever made.  The gathered operand's box is shifted by the tap offset
// (zero OOB fill = padding) and strided by the tensor map's element strides for stride-2 layers.
// M tile = 128 = two 64-channel blocks: two different taps when the gathered tensor has 64 channels, else two channel
// blocks of one tap.  The pixel range is split across CTAs (split-K); partial tiles are combined with fp32 red.global.add.
//
// Replaces the weight part of autograd's convolution_backward for model/resnet_deconv.py and model/hourglass.py layers
// in the bf16 precision mode.  Same argument meaning as awr_conv_wgrad_simt.
#include "tc_common.cuh"
#include "awr_b200.h"

namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int kABytes = 128 * 128;   // 2 blocks x (64 pixels x 64 ch bf16)

struct WgradTcParams {
  int N, Hc, Wc;                     // coarse grid (pointwise tensor)
  int Wt, Ht, Nt, tiles_w, tiles_h, tiles_nb, pix_blocks;   // 64-pixel K blocks
  int Cp, Cg, Ntile, tiles_n;        // pointwise channels (GEMM N), gathered channels (GEMM M)
  int m_items;                       // tap slots (Cg == 64: pairs of taps) or taps x Cg/128
  int R, S, stride, pad;
  int s_p, s_g, w_tap;
  int ksplit, stages;
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmP, float* __restrict__ dW,
                const __grid_constant__ WgradTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tfull_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stage_bytes = kABytes + p.Ntile * 128;

  // work item of this CTA
  int item = blockIdx.x;
  const int ks = item % p.ksplit; item /= p.ksplit;
  const int nt = item % p.tiles_n; const int mi = item / p.tiles_n;
  const int T = p.R * p.S;
  int tap[2], ch0[2];
  if (p.Cg == 64) { tap[0] = 2 * mi; tap[1] = 2 * mi + 1; ch0[0] = ch0[1] = 0; }
  else { const int cb = p.Cg / 128; tap[0] = tap[1] = mi / cb; ch0[0] = (mi % cb) * 128; ch0[1] = ch0[0] + 64; }
  const bool blk1_valid = tap[1] < T;
  if (!blk1_valid) tap[1] = tap[0];
  const int per = (p.pix_blocks + p.ksplit - 1) / p.ksplit;
  const int pb0 = ks * per, pb1 = min(p.pix_blocks, pb0 + per);
  const int iters = max(pb1 - pb0, 0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmP);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int pb = pb0; pb < pb1; ++pb) {
        const int tw = pb % p.tiles_w; int r = pb / p.tiles_w;
        const int th = r % p.tiles_h; const int tn = r / p.tiles_h;
        const int w0 = tw * p.Wt, h0 = th * p.Ht, n0 = tn * p.Nt;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
        uint8_t* sa = smem_raw + (smem_base - smem_u32(smem_raw)) + (size_t)stage * stage_bytes;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int tr = tap[b] / p.S, ts = tap[b] % p.S;
          tma_load_4d(sa + b * 8192, &tmG, &full_bar[stage], ch0[b], w0 * p.stride - p.pad + ts, h0 * p.stride - p.pad + tr, n0);
        }
        uint8_t* sb = sa + kABytes;
        for (int j = 0; j < p.Ntile / 64; ++j) tma_load_4d(sb + j * 8192, &tmP, &full_bar[stage], nt * p.Ntile + 64 * j, w0, h0, n0);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.Ntile, 1, 1);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_base + (uint32_t)(stage * stage_bytes), sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = umma_desc_sw128(sa + k * 2048, 8192, 1024);
          const uint64_t bd = umma_desc_sw128(sb + k * 2048, 8192, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (it | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (iters > 0) umma_commit(&tfull_bar);
      else mbar_arrive(&tfull_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane, blk = row >> 6, j = row & 63;
    const bool valid = (blk == 0 || blk1_valid) && iters > 0;
    float* base = dW + (size_t)tap[blk] * p.w_tap + (size_t)(ch0[blk] + j) * p.s_g + (size_t)(nt * p.Ntile) * p.s_p;
    mbar_wait(&tfull_bar, 0);
    tc_fence_after();
    if (iters > 0) {
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int ch = 0; ch < p.Ntile; ch += 32) {
        uint32_t v[32];
        tmem_ld32(t_addr + ch, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(base + (size_t)(ch + i) * p.s_p, __uint_as_float(v[i]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tmem_dealloc(tmem_base, 256); }
}

}  // namespace

extern "C" {

int awr_conv_wgrad_tc(const void* pointwise, const void* gathered, float* dW, int N, int Hc, int Wc, int Cp, int Hf, int Wf, int Cg, int R,
                      int S, int stride, int pad, int s_p, int s_g, int w_tap, void* stream) {
  AWR_HOST_CHECK(pointwise && gathered && dW && N > 0 && Cp % 64 == 0 && Cg % 64 == 0 && (Cg == 64 || Cg % 128 == 0));
  AWR_HOST_CHECK(R > 0 && S > 0 && (stride == 1 || stride == 2));
  AWR_HOST_CHECK(is_pow2(Wc) && is_pow2(Hc) && Wc <= 256 && Hc <= 256);
  WgradTcParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.Hc = Hc; p.Wc = Wc;
  p.Wt = Wc < 64 ? Wc : 64;
  p.Ht = (64 / p.Wt) < Hc ? (64 / p.Wt) : Hc;
  p.Nt = 64 / (p.Wt * p.Ht);
  p.tiles_w = Wc / p.Wt; p.tiles_h = Hc / p.Ht; p.tiles_nb = (N + p.Nt - 1) / p.Nt;
  p.pix_blocks = p.tiles_w * p.tiles_h * p.tiles_nb;
  p.Cp = Cp; p.Cg = Cg;
  p.Ntile = (Cp % 256 == 0) ? 256 : ((Cp % 128 == 0) ? 128 : 64);
  p.tiles_n = Cp / p.Ntile;
  const int T = R * S;
  p.m_items = (Cg == 64) ? (T + 1) / 2 : T * (Cg / 128);
  p.R = R; p.S = S; p.stride = stride; p.pad = pad;
  p.s_p = s_p; p.s_g = s_g; p.w_tap = w_tap;
  const int base_items = p.m_items * p.tiles_n;
  int ksplit = (2 * 148 + base_items - 1) / base_items;
  if (ksplit > p.pix_blocks) ksplit = p.pix_blocks;
  if (ksplit < 1) ksplit = 1;
  p.ksplit = ksplit;
  const int stage_bytes = kABytes + p.Ntile * 128;
  p.stages = (200 * 1024) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;

  CUtensorMap tmG, tmP;
  {
    const long long dims[4] = {Cg, Wf, Hf, N};
    const long long str[4] = {1, Cg, (long long)Wf * Cg, (long long)Hf * Wf * Cg};
    const int box[4] = {64, p.Wt * stride, p.Ht * stride, p.Nt};
    const int es[4] = {1, stride, stride, 1};
    if (!make_tmap_bf16(&tmG, gathered, 4, dims, str, box, es)) return AWR_ERR_DRIVER;
  }
  {
    const long long dims[4] = {Cp, Wc, Hc, N};
    const long long str[4] = {1, Cp, (long long)Wc * Cp, (long long)Hc * Wc * Cp};
    const int box[4] = {64, p.Wt, p.Ht, p.Nt};
    if (!make_tmap_bf16(&tmP, pointwise, 4, dims, str, box, nullptr)) return AWR_ERR_DRIVER;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  wgrad_tc_kernel<<<base_items * ksplit, kThreads, smem, (cudaStream_t)stream>>>(tmG, tmP, dW, p);
  AWR_LAUNCH_CHECK();
  return AWR_OK;
}

}  // extern "C"
