// Shared helpers for the sm_100a kernels of libawr_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

int awr_sm_budget();          // api.cu: grid cap of the persistent tensor-core kernels (148 unless the trainer reserves SMs for NCCL)

#define AWR_OK 0
#define AWR_ERR_BAD_ARG (-1)
#define AWR_ERR_UNSUPPORTED (-2)
#define AWR_ERR_DRIVER (-3)

#define AWR_HOST_CHECK(cond) do { if (!(cond)) return AWR_ERR_BAD_ARG; } while (0)
#define AWR_LAUNCH_CHECK() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; } while (0)

typedef __nv_bfloat16 bf16;

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------
// Every kernel of the step is launched with cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may be scheduled while
// the previous kernel of the stream is still draining, run their prologue (smem carve-up, mbarrier init, TMEM alloc, tensor-map
// prefetch) and then block in pdl_wait() until the previous grid has completed and flushed.  No global memory is touched before
// pdl_wait().  pdl_trigger() at kernel entry lets the NEXT kernel do the same with respect to this one.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_trigger(); pdl_wait(); }

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// same, as a thread-block cluster of `cluster` CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = (unsigned)cluster; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- order-independent cross-CTA accumulators --------------------------------------------------------------------------
// Per-channel sums that many CTAs contribute to (BatchNorm batch statistics and their backward sums) are accumulated as
// two-limb 128-bit fixed point, value = hi * 2^-24 + lo * 2^-72, with 64-bit integer atomics: every fp32 partial is split exactly
// (|p| < 2^39; bits below 2^-72 are rounded per addend), integer addition is associative, so the total does not depend on the
// order in which CTAs arrive -- together with fixed per-CTA work assignment this makes the statistics bit-reproducible run to run,
// which fp32 atomicAdd is not (the reference is bit-deterministic on CPU).  Layout: AwrAcc[n] = n x 16 bytes, zero-filled by the caller.
//
// Two builds share this 16-byte slot layout (and so one ABI): libawr_b200_det.so (-DAWR_DETERMINISTIC, `make DET=1`) uses the two-limb
// integers; the default libawr_b200.so keeps ONE fp32 atomic per contribution in the slot's first four bytes -- the integer pairs cost
// ~60 us per 2 ms training step (two 64-bit atomics per value, 592-way contended, in every BatchNorm reduction).  Weight gradients follow
// the same switch through GradT / grad_add: fp32 atomics into the flat gradient buffer, or AwrAcc slots that awr_grad_acc_finalize folds
// into it.
struct __align__(16) AwrAcc { long long hi, lo; };
#ifdef AWR_DETERMINISTIC
__device__ __forceinline__ void acc_add(AwrAcc* dst, float p) {
  const long long a = __float2ll_rn(p * 0x1p24f);
  const float r = fmaf(-__ll2float_rn(a), 0x1p-24f, p);          // exact remainder, |r| <= 2^-25
  const long long b = __float2ll_rn(r * 0x1p72f);
  if (a) atomicAdd(reinterpret_cast<unsigned long long*>(&dst->hi), (unsigned long long)a);
  if (b) atomicAdd(reinterpret_cast<unsigned long long*>(&dst->lo), (unsigned long long)b);
}
__device__ __forceinline__ float acc_value(longlong2 v) { return fmaf(__ll2float_rn(v.y), 0x1p-72f, __ll2float_rn(v.x) * 0x1p-24f); }
typedef AwrAcc GradT;
#else
__device__ __forceinline__ void acc_add(AwrAcc* dst, float p) { atomicAdd(reinterpret_cast<float*>(dst), p); }
__device__ __forceinline__ float acc_value(longlong2 v) { return __int_as_float((int)(unsigned)(v.x & 0xffffffffll)); }
typedef float GradT;
#endif
// one contribution to a weight-gradient element shared by several CTAs (split-K)
__device__ __forceinline__ void grad_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void grad_add(AwrAcc* p, float v) { acc_add(p, v); }
// eight consecutive accumulators -> floats: the eight 16-byte loads are issued together, then converted (32 registers in flight)
__device__ __forceinline__ void acc_get8(const AwrAcc* __restrict__ src, float (&out)[8]) {
  longlong2 r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = *reinterpret_cast<const longlong2*>(src + k);
#pragma unroll
  for (int k = 0; k < 8; ++k) out[k] = acc_value(r[k]);
}
// totals written by an EARLIER kernel: plain loads (L1 is invalidated at kernel boundaries; an L2-only load here makes every thread
// of a BatchNorm pass hammer the same few L2 lines -- measured 3.5x slower passes).  acc_get_cg: totals produced inside the same kernel.
__device__ __forceinline__ float acc_get(const AwrAcc* src) { return acc_value(*reinterpret_cast<const longlong2*>(src)); }
__device__ __forceinline__ float acc_get_cg(const AwrAcc* src) { return acc_value(__ldcg(reinterpret_cast<const longlong2*>(src))); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of NV values per thread. All threads get the result in v[]. smem: NV*32 floats.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[i * 32 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float x = (lane < nwarp) ? smem[i * 32 + lane] : 0.f;
    v[i] = warp_sum(x);
  }
}

// Huber / SmoothL1 with delta = 0.01 (reference model/loss.py:12-25)
#define AWR_HUBER_DELTA 0.01f
__device__ __forceinline__ float huber_val(float z) {
  float a = fabsf(z);
  return (a < AWR_HUBER_DELTA) ? 0.5f * z * z : AWR_HUBER_DELTA * (a - 0.5f * AWR_HUBER_DELTA);
}
__device__ __forceinline__ float huber_grad(float z) {
  float a = fabsf(z);
  return (a < AWR_HUBER_DELTA) ? z : ((z > 0.f) ? AWR_HUBER_DELTA : ((z < 0.f) ? -AWR_HUBER_DELTA : 0.f));
}

// dtype <-> float helpers for templated storage type (float or bf16)
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// value as it reads back after being stored in T (bf16 rounding; identity for float)
template <typename T> __device__ __forceinline__ float from_store(float v) { return to_f<T>(from_f<T>(v)); }

// 8-wide channel vector load/store (NHWC, C % 8 == 0): fp32 -> 2x float4, bf16 -> 1x uint4
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  __device__ __forceinline__ static void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ __forceinline__ static void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<bf16> {
  __device__ __forceinline__ static void load(const bf16* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ __forceinline__ static void store(bf16* p, const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};

// Raw8<T>: the 8 channels of Vec8<T> as loaded (bf16: one uint4; fp32: two float4), so that a batch of loads can be issued
// back to back and converted only when consumed.  Bandwidth-bound passes keep U of these per tensor in flight per thread:
// one 16-byte load per thread at 2-3 resident CTAs/SM is ~12 KB in flight per SM, a third of what HBM3e latency needs.
template <typename T> struct Raw8;
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <> struct Raw8<bf16> {
  uint4 r;
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};
