// Launchers of the HBM-bound NCSN kernels (ncsn_kernels.cu).  fp32 NHWC unless stated.
#pragma once
#include "common.cuh"

namespace asep {

// sums[N,C,2] (double) = per-(n,c) sum and sum of squares over HW (zeroed by the launcher)
void launch_in_stats(const float* x, double* sums, int N, int HW, int C, cudaStream_t s);
// (Conditional)InstanceNorm2dPlus folded to out = a*x + b per (n,c)  (score_network.py:203-221,
// score_network_v2.py:188-199).  gamma / alpha / beta: v1 = the three thirds of the Embedding table [classes, 3C], row
// idx[n] (gab_stride_n = 3C floats between rows); v2 = three shared vectors (idx = NULL, stride 0).
void launch_in_coef(const double* sums, const float* gamma, const float* alpha, const float* beta, int gab_stride_n,
                    const int* idx, const float* in_gamma,
                    const float* in_beta, float2* coef, int N, int HW, int C, cudaStream_t s);
// y (bf16) = act(coef.a * x + coef.b); coef may be NULL (plain cast); do_elu applies ELU(alpha=1)
// y_lo (may be NULL) receives bf16(v - y): the low word of the split-bf16 operand (ASEP_PREC_BF16X3)
void launch_prep(const float* x, const float2* coef, __nv_bfloat16* y, __nv_bfloat16* y_lo, int N, int HW, int C, int do_elu,
                 cudaStream_t s);
// 5x5 stride-1 'same' pooling: average over in-bounds taps (Keras AveragePooling2D) or max (MaxPooling2D)
// (one rolling pass; `tmp` is unused and kept for the callers' signature)
void launch_pool5(const float* x, float* tmp, float* y, int N, int H, int W, int C, int is_max, cudaStream_t s);
// AveragePooling2D(2): [N,2Hout,2Wout,C] -> [N,Hout,Wout,C]
void launch_avgpool2(const float* x, float* y, int N, int Hout, int Wout, int C, cudaStream_t s);
// y[N,2h,2w,C] = add + tf.image.resize(x[N,h,w,C], bilinear, half-pixel centres); add may be NULL
void launch_resize2x_add(const float* x, const float* add, float* y, int N, int h, int w, int C, cudaStream_t s);
void launch_elu(const float* x, float* y, long long n, cudaStream_t s);
// y = elu(x) and sums[N,C,2] (double, zeroed here) = per-(n,c) sum / sum of squares of y in the same pass
void launch_elu_stats(const float* x, float* y, double* sums, int N, int HW, int C, cudaStream_t s);
void launch_add(const float* x, const float* z, float* y, long long n, cudaStream_t s);
// begin_conv 3x3 (1 -> Cout) with bias; rescale=1 applies x <- 2x-1 first (v1)
void launch_begin_conv(const float* x, const float* k, const float* bias, float* y, int N, int H, int W, int Cout,
                       int rescale, cudaStream_t s);
// end_conv 3x3 (C -> 1) on the bf16 normalised/activated tensor; sigmas != NULL divides by sigmas[idx[n]] (v2)
// (bias: device pointer to the single bias value)
void launch_end_conv(const __nv_bfloat16* x, const __nv_bfloat16* x_lo, const float* k, const float* bias, const float* sigmas, const int* idx, float* y,
                     int N, int H, int W, int C, cudaStream_t s);

}  // namespace asep
