// Host-side description of one tensor-core convolution call, shared by conv_tc.cu (one TMA box per tap) and
// conv_halo_tc.cu (all taps from one halo tile).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int kConvMaxTaps = 16;
constexpr int kConvMaxClasses = 4;

struct ConvTapClass {
  int ntaps, py, px, pad_;
  short oy[kConvMaxTaps], ox[kConvMaxTaps], widx[kConvMaxTaps];
};

struct ConvGeom {
  int N, Hi, Wi, Ck;            // gathered tensor (NHWC bf16) and contraction channels
  int Ho, Wo, Cn;               // produced tensor
  int Hc, Wc, out_s;            // coarse grid (GEMM rows) and output pixels per coarse pixel
  int a_stride;                 // element stride of the gather per coarse pixel
  int R, S, w_sk, w_sn, w_tap, b_mn;
  int out_mode, n_valid, accumulate;
  int nclasses;
  ConvTapClass cls[kConvMaxClasses];
};

bool conv_make_weight_map(CUtensorMap* m, const ConvGeom& g, const void* w, int Ntile);
bool conv_halo_supported(const ConvGeom& g);
int conv_halo_launch(const ConvGeom& g, const void* in, const void* w, const float* bias, void* out, void* stats /* AwrAcc[2*Cn] or NULL */, cudaStream_t stream);
