// Launchers of the backward / training kernels of the NCSN score networks (ncsn_train_kernels.cu, conv_wgrad_tc.cu).
// Reference: train_ncsn.py:26-57 (denoising score matching step); layer semantics as ncsn_kernels.h.
// Every "+=" below ACCUMULATES into its destination (a tensor may feed several consumers).
#pragma once
#include "common.cuh"

namespace asep {

// x~ = x + sigma[idx[n]] * noise  (train_ncsn.py:36-41; noise is a standard-normal draw)
void launch_dsm_perturb(const float* x, const float* noise, const float* sigmas, const int* idx, float* xt, int N, int HW,
                        cudaStream_t s);
// loss += sum_n 1/2 * sum_p (score + noise/sigma_n)^2 * sigma_n^2 / global_batch   (train_ncsn.py:26-29,42-43: target =
// -noise*sigma/sigma^2, sample weight sigma^2); gscore = d loss / d score
void launch_dsm_loss(const float* score, const float* noise, const float* sigmas, const int* idx, float* gscore, double* loss,
                     int N, int HW, double inv_global_batch, cudaStream_t s);
void launch_double_to_float(const double* src, float* dst, cudaStream_t s);

// ---- end_conv (C -> 1): gs [N,H,W] is d loss / d score; inv_sigma: v2 divides the raw output by sigma[idx[n]]
// gin [N,H,W,C] (written) = d loss / d (bf16 operand of end_conv)
void launch_end_conv_bwd_data(const float* gs, const float* k, const float* sigmas, const int* idx, float* gin, int N, int H,
                              int W, int C, cudaStream_t s);
// dk [9,C] +=, dbias [1] +=
void launch_end_conv_bwd_w(const float* gs, const __nv_bfloat16* x, const __nv_bfloat16* x_lo, const float* sigmas,
                           const int* idx, float* dk, float* dbias, int N, int H, int W, int C, cudaStream_t s);
// ---- begin_conv (1 -> Cout): dk [9,Cout] +=, dbias [Cout] +=
void launch_begin_conv_bwd_w(const float* x, const float* gout, float* dk, float* dbias, int N, int H, int W, int Cout,
                             int rescale, cudaStream_t s);

// ---- y = act(a*xv + b) (k_prep) backward.  g' = gy * act'(a*xv+b).
// red [N,C,2] doubles (zeroed here) <- (sum g', sum g' * xv) per (n,c)
void launch_prep_bwd_reduce(const float* xv, const float* gy, const float2* coef, int do_elu, double* red, int N, int HW, int C,
                            cudaStream_t s);
// Instance-norm++ chain rule per (n,c) (score_network.py:203-221): from red, the forward statistics `sums` and the
// parameters -> qr [N,C] = (Q, R) with d loss / d x_s = Q * x_s + R (the path through mean / variance / mu~), and the
// parameter gradients (+=): d_gamma / d_alpha / d_beta at row idx[n] (stride_n floats between rows; 0 for v2's shared
// vectors), d_in_gamma, d_in_beta.
void launch_norm_bwd_coef(const double* red, const double* sums, const float* gamma, const float* alpha, const float* beta,
                          int stride_n, const int* idx, const float* in_gamma, const float* in_beta, float* d_gamma,
                          float* d_alpha, float* d_beta, float* d_in_gamma, float* d_in_beta, float2* qr, int N, int HW, int C,
                          cudaStream_t s);
// gx += [gy != NULL] a * gy * act'(a*x+b)  +  [qr != NULL] (Q * x + R)      (coef == NULL: a = 1, b = 0)
void launch_prep_bwd_apply(const float* x, const float* gy, const float2* coef, int do_elu, const float2* qr, float* gx, int N,
                           int HW, int C, cudaStream_t s);

// ---- element-wise / stencil backward passes
void launch_axpy(const float* g, float* dst, long long n, cudaStream_t s);                                 // dst += g
void launch_elu_bwd(const float* x, const float* gy, float* gx, long long n, cudaStream_t s);              // gx += gy * elu'(x)
void launch_avgpool2_bwd(const float* gout, float* gin, int N, int Hout, int Wout, int C, cudaStream_t s); // gin [N,2H,2W,C] +=
// 5x5 'same' average pooling (divisor = in-bounds taps): gin += box-sum of gout / count; tmp: scratch of the same size
void launch_pool5_avg_bwd(const float* gout, float* tmp, float* gin, int N, int H, int W, int C, cudaStream_t s);
// 5x5 'same' max pooling: gin[argmax of the window in x] += gout (first maximum in row-major order)
void launch_pool5_max_bwd(const float* x, const float* gout, float* gin, int N, int H, int W, int C, cudaStream_t s);
// transpose of the bilinear x2 up-sampling: glow [N,h,w,C] += R^T gout [N,2h,2w,C]
void launch_resize2x_bwd(const float* gout, float* glow, int N, int h, int w, int C, cudaStream_t s);
// dbias [C] += sum over pixels of g [P,C]
void launch_colsum_f32(const float* g, float* dbias, long long P, int C, cudaStream_t s);

// ---- tile images of a convolution kernel (Keras HWIO [taps,Cin,Cout] fp32 in the flat parameter vector) rebuilt on the
// device after an optimizer step.  transposed = 0: the forward operand (conv_tc.cu layout: per (tap, 64-input-channel
// panel) a [Cout x 64] SWIZZLE_128B image); transposed = 1: the data-gradient operand, i.e. the forward image of the
// kernel K'[tap'][co][ci] = K[taps-1-tap'][ci][co].  lo = 1 writes bf16(v - bf16(v)) (split-bf16 second term).
void launch_build_conv_image(const float* kernel, __nv_bfloat16* img, int taps, int Cin, int Cout, int transposed, int lo,
                             cudaStream_t s);

// ---- weight gradient of a 'same' stride-1 convolution on tcgen05 (conv_wgrad_tc.cu):
// dk[tap][ci][co] += sum_p x[p + offset(tap)][ci] * g[p][co]      x [N,H,W,Cin] bf16, g [N,H,W,Cout] bf16, dk fp32 HWIO
void conv_wgrad_tc(const __nv_bfloat16* x, const __nv_bfloat16* g, float* dk, int N, int H, int W, int Cin, int Cout, int ksize,
                   int dil, cudaStream_t s);
bool conv_wgrad_tc_supported(int Cin, int Cout, int H, int W);

}  // namespace asep
