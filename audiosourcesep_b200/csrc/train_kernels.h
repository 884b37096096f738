// Launchers of the Glow training-step kernels (train_kernels.cu).
#pragma once
#include "common.cuh"

namespace asep {

// Device pointers of one flow step's parameters / derived constants and the offsets of its trainables in the flat
// gradient vector (order of weights.py:glow_param_shapes filtered by is_trainable).
struct StepTrainPtrs {
  int C, F;
  const float *an_ls, *an_shift, *P, *L, *U, *logS, *signS;
  const float *k1, *c1, *bn1_gamma, *bn1_beta, *bn1_mean, *bn1_var;
  const float *k2, *c2, *bn2_gamma, *bn2_beta, *bn2_mean, *bn2_var;
  const float *k3, *c3;
  float *sc, *g1f, *b1f, *g2f, *b2f, *k2t;          // derived (written by k_derive_step)
  long long o_an_ls, o_an_shift, o_L, o_U, o_logS, o_k1, o_c1, o_bn1_gamma, o_bn1_beta, o_k2, o_c2, o_bn2_gamma,
      o_bn2_beta, o_k3, o_c3;
};

// One row per flow step of the device-side refresh table: everything k_derive_step / k_build_tc_images /
// k_build_tc_biases need, so that the constants of ALL steps are rebuilt by three launches after an optimizer step
// (grid dimension = step) instead of three launches per step.
struct StepRefresh {
  StepTrainPtrs sp;
  __nv_bfloat16 *fwd_img, *bwd_img;
  int k1p_f, n3p_f, k1p_b, n3p_b;
  float *bias1, *bias2, *const3, *c3;
  double HW;
  int f16;
};

void launch_axpy(const float* x, const float* n, float sigma, float* y, long long total, cudaStream_t s);
void launch_colsum(const float* X, float* out, long long M, int F, cudaStream_t s);                     // out += column sums
void launch_wgrad_tn(const float* A, const float* B, float* Q, long long M, int F, cudaStream_t s);     // Q += A^T B
void launch_wgrad_conv3(const float* a2, const float* gr, float* R3, float* S3, int N, int H, int W, int C, int F,
                        cudaStream_t s);
void launch_wgrad_conv1(const float* state, const float* gp1, float* dK1, float* dc1, int N, int H, int W, int C, int F,
                        float gs, cudaStream_t s);
void launch_step_stats(const float* gu, const float* gxb, const float* u, const float* sc, double* stats, long long M,
                       int C, cudaStream_t s);
void launch_finalize_step(const StepTrainPtrs& sp, const float* Q2, const float* dc2, const float* R3, const float* S3,
                          const double* stats, float* grads, double Mpix, float gs, cudaStream_t s);
// tensor-core path: R3t[k][tap*C+c] (row stride r3_ld) and D1t[f][tap*Ch+ci] (row stride d1_ld) come from wgrad_tc
void launch_finalize_step_tc(const StepTrainPtrs& sp, const float* Q2, const float* dc2, const float* R3t, int r3_ld,
                             const float* S3, const float* D1t, int d1_ld, const float* dc1, const double* stats,
                             float* grads, double Mpix, float gs, cudaStream_t s);
void launch_im2col_gr(const float* gr, __nv_bfloat16* G9, int N, int H, int W, int C, int ld, cudaStream_t s);
void launch_im2col_xb(const float* state, __nv_bfloat16* X9, int N, int H, int W, int C, int ld, cudaStream_t s);
void launch_s3(const float* gr, float* S3, int N, int H, int W, int C, cudaStream_t s);
void launch_colsum_bf16(const __nv_bfloat16* X, float* out, long long M, int F, cudaStream_t s);
// device twin of nn_tc_prepare: tile images of both directions + folded biases of one step
// (table: device array of n_steps rows)
void launch_build_tc_all(const StepRefresh* table, int n_steps, cudaStream_t s);
void launch_prior_grads(const float* z, const float* loc, const float* ls, float* gloc, float* gls, int N, int D, float gs,
                        cudaStream_t s);
void launch_loss(const double* acc_ld, const double* acc_prior, const double* cst, double extra_const, int N,
                 double inv_batch, float* loss, cudaStream_t s);
// ldc[step] = H*W*(sum log_scale + sum log_S) of every step
void launch_derive_all(const StepRefresh* table, int n_steps, double* ldc, int need_k2t, cudaStream_t s);
void launch_sum_doubles(const double* v, int n, double* out, cudaStream_t s);
void launch_adamax(float* theta, const float* g, float* m, float* u, long long n, float lr_t, float b1, float b2, float eps,
                   cudaStream_t s);
void launch_adam(float* theta, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2, float eps,
                 cudaStream_t s);

}  // namespace asep
