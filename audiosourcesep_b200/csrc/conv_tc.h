// Generic stride-1 "same" convolution (1x1 or 3x3, dilation 1/2/4) of the NCSN score networks as an
// implicit GEMM on the 5th-generation tensor cores (reference call sites: ncsn/score_network.py:13,38,66,
// 131-159,234,238 and ncsn/score_network_v2.py, Keras Conv2D padding='same').  See conv_tc.cu.
#pragma once
#include <vector>
#include "common.cuh"

namespace asep {

struct ConvWeightsTC {
  __nv_bfloat16* img = nullptr;   // [taps * kpanels][Cout rows x 64 k] bf16 SWIZZLE_128B K-major tile images
  float* bias = nullptr;          // [Cout] or NULL
  int Cin = 0, Cout = 0, ksize = 0, dil = 1;
  size_t bytes = 0;
  bool bias_owned = true;         // false: bias points into a parameter vector owned by the model (training)
};

// uninitialised image of the given geometry (filled on the device: ncsn_train_kernels.h launch_build_conv_image)
void conv_tc_alloc(ConvWeightsTC& w, int ksize, int Cin, int Cout, int dilation);

// kernel: Keras HWIO [k,k,Cin,Cout] on the host; bias may be NULL.
void conv_tc_prepare(ConvWeightsTC& w, const float* kernel_hwio, const float* bias, int ksize, int Cin, int Cout,
                     int dilation);
void conv_tc_release(ConvWeightsTC& w);
bool conv_tc_supported(int Cin, int Cout, int H, int W);

// out[n,h,w,:] = conv(xin)[n,h,w,:] + bias + add[n,h,w,:]   (add may be NULL or alias out)
// xin: bf16 NHWC [N,H,W,Cin] (already normalised / activated), out/add: fp32 NHWC [N,H,W,Cout].
// stats (may be NULL): [N, Cout, 2] doubles that receive the per-(image, channel) sum and sum of squares of `out`
// (zeroed here, accumulated by the epilogue) - the statistics of the instance norm that usually follows.
void conv_tc_forward(const ConvWeightsTC& w, const __nv_bfloat16* xin, const float* add, float* out, int N, int H,
                     int W, cudaStream_t s, double* stats = nullptr, __nv_bfloat16* out_bf16 = nullptr, const float* add2 = nullptr);
// out_bf16 (may be NULL): bf16 NHWC copy of `out`, written by the same epilogue.  add2 (may be NULL): a second fp32 tensor
// summed into the output like `add`.

// profiling of the conv launches (same contract as nn_tc_profile)
void conv_tc_profile(int on);
bool conv_tc_profile_enabled();
void conv_tc_profile_read(double* total_ms, long long* launches, double* flops);

}  // namespace asep
