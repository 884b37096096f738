// Parameter-gradient, derived-constant and optimiser kernels of the Glow training step
// (reference: train_glow.py:29-44 loss / tape.gradient / apply_gradients, train_utils.py:23-41 Adamax,
// train_noisy_glow.py:30-33 noise perturbation).  The data-gradient sweep is the one grad_log_prob uses; these
// kernels add, per flow step, the weight gradients of the three convolutions (CUDA-core fp32 "exact" mode in this
// round: split-K outer products with fp32 atomics), the folded-BatchNorm chain rule, the ActNorm / LU-parameterised
// 1x1 gradients and the prior gradients, and refresh every per-step constant on the device after an update.
#include "train_kernels.h"

#include <cuda_fp16.h>

namespace asep {

namespace {

constexpr double kBnEpsD = 1e-3;

// ------------------------------------------------------------------ x <- x + sigma * noise (train_noisy_glow.py:31-32)
__global__ void k_axpy(const float* __restrict__ x, const float* __restrict__ n, float sigma, float* __restrict__ y, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) y[i] = x[i] + sigma * n[i];
}

// ------------------------------------------------------------------ column sums of X [M,F]
__global__ void __launch_bounds__(512) k_colsum(const float* __restrict__ X, float* __restrict__ out, long long M, int F,
                                                int rows_per_block) {
  const int c = threadIdx.x;
  if (c >= F) return;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += X[r * F + c];
  atomicAdd(out + c, s);
}

// ------------------------------------------------------------------ Q[i][o] += sum_p A[p][i] * B[p][o]   (split over p)
__global__ void __launch_bounds__(256) k_wgrad_tn(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Q,
                                                  long long M, int F, int rows_per_block) {
  constexpr int BT = 64, BK = 16;
  __shared__ float As[BK][BT], Bs[BK][BT];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int i0 = blockIdx.y * BT, o0 = blockIdx.x * BT;
  const long long p0 = (long long)blockIdx.z * rows_per_block, p1 = min(M, p0 + rows_per_block);
  float acc[4][4] = {};
  for (long long pk = p0; pk < p1; pk += BK) {
    for (int t = threadIdx.x; t < BK * BT; t += 256) {
      const int kk = t / BT, cc = t % BT;
      const long long p = pk + kk;
      As[kk][cc] = p < p1 ? A[p * F + i0 + cc] : 0.f;
      Bs[kk][cc] = p < p1 ? B[p * F + o0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(Q + (size_t)(i0 + ty * 4 + i) * F + o0 + tx * 4 + j, acc[i][j]);
}

// ------------------------------------------------------------------ conv3 weight gradient pieces
// R3[tap][k][c] += sum_p a2[p+off(tap)][k] * gr[p][c];  S3[tap][c] += sum_{p: p+off in bounds} gr[p][c]
template <int C>
__global__ void __launch_bounds__(512) k_wgrad_conv3(const float* __restrict__ a2, const float* __restrict__ gr,
                                                     float* __restrict__ R3, float* __restrict__ S3, int H, int W, int F,
                                                     long long M, int px_per_block) {
  extern __shared__ float sg[];                 // [px_per_block][C]
  const int k = threadIdx.x;
  const long long p0 = (long long)blockIdx.x * px_per_block;
  const int np = (int)min((long long)px_per_block, M - p0);
  for (int t = threadIdx.x; t < np * C; t += blockDim.x) sg[t] = gr[p0 * C + t];
  __syncthreads();
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    float acc[C];
    float ssum = 0.f;                            // S3 partial: thread c < C owns channel c
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int i = 0; i < np; ++i) {
      const long long p = p0 + i;
      const int w = (int)(p % W), h = (int)((p / W) % H);
      const int hh = h + dy, ww = w + dx;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      if (k < F) {
        const float v = a2[(p + (long long)dy * W + dx) * F + k];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = fmaf(v, sg[i * C + c], acc[c]);
      }
      if (k < C) ssum += sg[i * C + k];
    }
    if (k < F) {
#pragma unroll
      for (int c = 0; c < C; ++c) atomicAdd(R3 + ((size_t)tap * F + k) * C + c, acc[c]);
    }
    if (k < C) atomicAdd(S3 + tap * C + k, ssum);
  }
}

// ------------------------------------------------------------------ conv1 weight and bias gradient (straight into grads)
// dK1[tap][ci][f] += gs * sum_p xb[p+off(tap)][ci] * gp1[p][f];   dc1[f] += gs * sum_p gp1[p][f]
template <int Ch>
__global__ void __launch_bounds__(512) k_wgrad_conv1(const float* __restrict__ state, const float* __restrict__ gp1,
                                                     float* __restrict__ dK1, float* __restrict__ dc1, int H, int W, int F,
                                                     long long M, int px_per_block, float gs) {
  const int f = threadIdx.x;
  if (f >= F) return;
  constexpr int C = 2 * Ch;
  const long long p0 = (long long)blockIdx.x * px_per_block, p1 = min(M, p0 + px_per_block);
  float acc[9][Ch];
  float accb = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < Ch; ++c) acc[t][c] = 0.f;
  for (long long p = p0; p < p1; ++p) {
    const float g = gp1[p * F + f];
    accb += g;
    const int w = (int)(p % W), h = (int)((p / W) % H);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3 - 1, dx = t % 3 - 1;
      const int hh = h + dy, ww = w + dx;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const float* xb = state + (p + (long long)dy * W + dx) * C + Ch;
#pragma unroll
      for (int c = 0; c < Ch; ++c) acc[t][c] = fmaf(__ldg(xb + c), g, acc[t][c]);
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < Ch; ++c) atomicAdd(dK1 + ((size_t)t * Ch + c) * F + f, gs * acc[t][c]);
  atomicAdd(dc1 + f, gs * accb);
}

// ------------------------------------------------------------------ ActNorm / 1x1 statistics of one step
// a = u . W^-1 (= x*scale + shift), gu_full = [gu_a, gu_b + gxb], ga = gu_full . W^T
// stats: [0,C) sum ga   [C,2C) sum ga*(a - shift)   [2C, 2C+C*C) sum a[i]*gu_full[o]
template <int C>
__global__ void __launch_bounds__(256) k_step_stats(const float* __restrict__ gu, const float* __restrict__ gxb,
                                                    const float* __restrict__ u, const float* __restrict__ sc,
                                                    double* __restrict__ stats, long long M) {
  constexpr int PB = C >= 16 ? 128 : 256;           // pixels per block (= blockDim.x); keeps static smem < 48 KB
  __shared__ float cst[2 * C + 2 * C * C];
  __shared__ float sa[PB][C + 1], sgu[PB][C + 1], sga[PB][C + 1];
  for (int i = threadIdx.x; i < 2 * C + 2 * C * C; i += blockDim.x) cst[i] = sc[i];
  __syncthreads();
  const long long p = (long long)blockIdx.x * PB + threadIdx.x;
  {
    float uv[C], g[C];
    if (p < M) {
#pragma unroll
      for (int c = 0; c < C; ++c) { uv[c] = u[p * C + c]; g[c] = gu[p * C + c]; }
#pragma unroll
      for (int c = 0; c < C / 2; ++c) g[C / 2 + c] += gxb[p * (C / 2) + c];
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) { uv[c] = 0.f; g[c] = 0.f; }
    }
    const float* Wm = cst + 2 * C;
    const float* Wi = cst + 2 * C + C * C;
#pragma unroll
    for (int o = 0; o < C; ++o) {
      float a = 0.f, ga = 0.f;
#pragma unroll
      for (int i = 0; i < C; ++i) { a = fmaf(uv[i], Wi[i * C + o], a); ga = fmaf(g[i], Wm[o * C + i], ga); }
      sa[threadIdx.x][o] = p < M ? a : 0.f;
      sga[threadIdx.x][o] = ga;
      sgu[threadIdx.x][o] = g[o];
    }
  }
  __syncthreads();
  // outer-product sums: thread t < C*C owns (i,o); threads C*C .. C*C+2C own the vector sums
  for (int t = threadIdx.x; t < C * C + 2 * C; t += blockDim.x) {
    double s = 0.0;
    if (t < C * C) {
      const int i = t / C, o = t % C;
      float f = 0.f;
      for (int q = 0; q < PB; ++q) f = fmaf(sa[q][i], sgu[q][o], f);
      s = f;
      atomicAdd(stats + 2 * C + t, s);
    } else if (t < C * C + C) {
      const int c = t - C * C;
      float f = 0.f;
      for (int q = 0; q < PB; ++q) f += sga[q][c];
      atomicAdd(stats + c, (double)f);
    } else {
      const int c = t - C * C - C;
      float f = 0.f;
      const long long base = (long long)blockIdx.x * PB;
      for (int q = 0; q < PB; ++q)
        if (base + q < M) f = fmaf(sga[q][c], sa[q][c] - cst[C + c], f);
      atomicAdd(stats + C + c, (double)f);
    }
  }
}

// ------------------------------------------------------------------ per-step chain rule -> parameter gradients
// conv2 + BN1: one block per input channel i, threads over output channels o (coalesced rows of Q2 / K2)
__global__ void __launch_bounds__(512) k_fin_conv2(const StepTrainPtrs sp, const float* __restrict__ Q2,
                                                   const float* __restrict__ dc2, float* __restrict__ grads, float gs) {
  const int F = sp.F, i = blockIdx.x, t = threadIdx.x;
  __shared__ float red[2][16];
  const float g1 = sp.g1f[i], b1 = sp.b1f[i];
  float dg = 0.f, db = 0.f;
  for (int o = t; o < F; o += blockDim.x) {
    const float q = Q2[(size_t)i * F + o], k = sp.k2[(size_t)i * F + o], d2 = dc2[o];
    grads[sp.o_k2 + (size_t)i * F + o] = gs * (g1 * q + b1 * d2);
    dg = fmaf(q, k, dg);
    db = fmaf(d2, k, db);
  }
  dg = warp_sum(dg); db = warp_sum(db);
  if ((t & 31) == 0) { red[0][t >> 5] = dg; red[1][t >> 5] = db; }
  __syncthreads();
  if (t == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) { a += red[0][w]; b += red[1][w]; }
    const float s = sqrtf(sp.bn1_var[i] + (float)kBnEpsD);
    grads[sp.o_bn1_gamma + i] = gs * (a / s - b * sp.bn1_mean[i] / s);
    grads[sp.o_bn1_beta + i] = gs * b;
    grads[sp.o_c2 + i] = gs * dc2[i];
  }
}

// conv3 + BN2 (+ conv1 copy on the tensor-core path): one block per hidden channel k, threads over (tap, c)
__global__ void __launch_bounds__(256) k_fin_conv3(const StepTrainPtrs sp, const float* __restrict__ R3, int r3_tap_stride,
                                                   int r3_k_stride, const float* __restrict__ S3, const float* __restrict__ D1t,
                                                   int d1_ld, const float* __restrict__ dc1, float* __restrict__ grads, float gs) {
  const int F = sp.F, C = sp.C, k = blockIdx.x, t = threadIdx.x;
  __shared__ float red[2][8];
  const float g2 = sp.g2f[k], b2 = sp.b2f[k];
  float dg = 0.f, db = 0.f;
  for (int n = t; n < 9 * C; n += blockDim.x) {
    const int tap = n / C, c = n % C;
    const size_t idx = ((size_t)tap * F + k) * C + c;
    const float r = R3[(size_t)tap * r3_tap_stride + (size_t)k * r3_k_stride + c], s3 = S3[n], kk = sp.k3[idx];
    grads[sp.o_k3 + idx] = gs * (g2 * r + b2 * s3);
    dg = fmaf(r, kk, dg);
    db = fmaf(s3, kk, db);
  }
  dg = warp_sum(dg); db = warp_sum(db);
  if ((t & 31) == 0) { red[0][t >> 5] = dg; red[1][t >> 5] = db; }
  __syncthreads();
  if (t == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) { a += red[0][w]; b += red[1][w]; }
    const float s = sqrtf(sp.bn2_var[k] + (float)kBnEpsD);
    grads[sp.o_bn2_gamma + k] = gs * (a / s - b * sp.bn2_mean[k] / s);
    grads[sp.o_bn2_beta + k] = gs * b;
  }
  if (D1t != nullptr) {                                        // f = k: dK1[tap][ci][f] = D1t[f][tap*Ch+ci]
    const int Ch = C / 2;
    for (int n = t; n < 9 * Ch; n += blockDim.x) grads[sp.o_k1 + (size_t)n * F + k] = gs * D1t[(size_t)k * d1_ld + n];
    if (t == 0) grads[sp.o_c1 + k] = gs * dc1[k];
  }
}

// small tensors: conv3 bias, ActNorm, LU-parameterised 1x1
__global__ void __launch_bounds__(256) k_fin_small(const StepTrainPtrs sp, const float* __restrict__ S3,
                                                   const double* __restrict__ stats, float* __restrict__ grads, double Mpix,
                                                   float gs) {
  const int C = sp.C, t = threadIdx.x;
  __shared__ double sP[256], sL[256], sU[256], sT[256], sD[256];
  if (t < C) grads[sp.o_c3 + t] = gs * S3[4 * C + t];            // centre tap is always in bounds: sum_p gr[p][c]
  if (t < C) {
    grads[sp.o_an_shift + t] = gs * (float)stats[t];
    grads[sp.o_an_ls + t] = gs * (float)(stats[C + t] + Mpix);   // + d(H*W*sum log_scale)/d log_scale per sample
  }
  // 1x1: W = P L' U',  dL' = P^T dW U'^T,  dU' = (P L')^T dW
  if (t < C * C) {
    const int i = t / C, j = t % C;
    sP[t] = sp.P[t];
    sL[t] = i == j ? 1.0 : (j < i ? (double)sp.L[t] : 0.0);
    sU[t] = i == j ? (double)sp.signS[i] * exp((double)sp.logS[i]) : (j > i ? (double)sp.U[t] : 0.0);
    sD[t] = stats[2 * C + t];
  }
  __syncthreads();
  if (t < C * C) {                                                // T = P^T dW
    const int i = t / C, j = t % C;
    double a = 0.0;
    for (int k = 0; k < C; ++k) a += sP[k * C + i] * sD[k * C + j];
    sT[t] = a;
  }
  __syncthreads();
  if (t < C * C) {
    const int i = t / C, j = t % C;
    double dL = 0.0, dU = 0.0;
    for (int k = 0; k < C; ++k) dL += sT[i * C + k] * sU[j * C + k];          // (T U'^T)[i][j]
    for (int k = 0; k < C; ++k) dU += sL[k * C + i] * sT[k * C + j];          // (L'^T T)[i][j]
    grads[sp.o_L + t] = j < i ? gs * (float)dL : 0.f;
    grads[sp.o_U + t] = j > i ? gs * (float)dU : 0.f;
    if (i == j) grads[sp.o_logS + i] = gs * (float)(dU * sU[t] + Mpix);        // d diag / d log_s = diag; + log-det term
  }
}

// ------------------------------------------------------------------ prior gradients (flow_builder.py:132-139)
__global__ void k_prior_grads(const float* __restrict__ z, const float* __restrict__ loc, const float* __restrict__ ls,
                              float* __restrict__ gloc, float* __restrict__ gls, int N, int D, float gs) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const float inv = expf(-ls[d]);
  float a = 0.f, b = 0.f;
  for (int n = 0; n < N; ++n) {
    const float q = (z[(size_t)n * D + d] - loc[d]) * inv;
    a += q * inv;
    b += q * q - 1.f;
  }
  gloc[d] = gs * a;
  gls[d] = gs * b;
}

// loss = -(sum_n (acc_ld[n] + acc_prior[n]) + N * const) / global_batch
__global__ void k_loss(const double* __restrict__ acc_ld, const double* __restrict__ acc_prior, const double* __restrict__ cst,
                       double extra_const, int N, double inv_batch, float* __restrict__ loss) {
  __shared__ double red[256];
  double s = 0.0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) s += acc_ld[n] + acc_prior[n];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(-(red[0] + (double)N * (cst[0] + extra_const)) * inv_batch);
}

// ------------------------------------------------------------------ derived constants of one step, on the device
__global__ void __launch_bounds__(512) k_derive_step(const StepRefresh* __restrict__ table, double* __restrict__ ldc_all, int need_k2t) {
  const StepTrainPtrs sp = table[blockIdx.x].sp;                   // one block per flow step
  const double HW = table[blockIdx.x].HW;
  double* ldc = ldc_all + blockIdx.x;
  const int C = sp.C, F = sp.F, t = threadIdx.x;
  __shared__ double sP[256], sL[256], sU[256], sA[256], sLi[256], sUi[256];
  if (t < C * C) {
    const int i = t / C, j = t % C;
    sP[t] = sp.P[t];
    sL[t] = i == j ? 1.0 : (j < i ? (double)sp.L[t] : 0.0);
    sU[t] = i == j ? (double)sp.signS[i] * exp((double)sp.logS[i]) : (j > i ? (double)sp.U[t] : 0.0);
  }
  __syncthreads();
  if (t < C * C) {                                  // A = L' U'
    const int i = t / C, j = t % C;
    double a = 0.0;
    for (int k = 0; k < C; ++k) a += sL[i * C + k] * sU[k * C + j];
    sA[t] = a;
  }
  // triangular inverses, one column per thread
  if (t < C) {
    const int j = t;
    for (int i = 0; i < C; ++i) {                    // L'^-1 (unit lower): forward substitution
      double v = i == j ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) v -= sL[i * C + k] * sLi[k * C + j];
      sLi[i * C + j] = v;
    }
    for (int i = C - 1; i >= 0; --i) {               // U'^-1: back substitution
      double v = i == j ? 1.0 : 0.0;
      for (int k = i + 1; k < C; ++k) v -= sU[i * C + k] * sUi[k * C + j];
      sUi[i * C + j] = v / sU[i * C + i];
    }
  }
  __syncthreads();
  if (t < C * C) {
    const int i = t / C, j = t % C;
    double w = 0.0, b = 0.0;
    for (int k = 0; k < C; ++k) w += sP[i * C + k] * sA[k * C + j];                 // W = P (L' U')
    for (int k = 0; k < C; ++k) b += sUi[i * C + k] * sLi[k * C + j];               // B = U'^-1 L'^-1
    sp.sc[2 * C + t] = (float)w;
    sL[t] = b;                                       // L' is no longer needed
  }
  __syncthreads();
  if (t < C * C) {                                   // W^-1 = B P^T  (P is a permutation matrix)
    const int i = t / C, j = t % C;
    double v = 0.0;
    for (int k = 0; k < C; ++k) v += sL[i * C + k] * sP[j * C + k];
    sp.sc[2 * C + C * C + t] = (float)v;
  }
  if (t < C) {
    sp.sc[t] = expf(sp.an_ls[t]);
    sp.sc[C + t] = sp.an_shift[t];
  }
  if (t == 0) {
    double s = 0.0;
    for (int c = 0; c < C; ++c) s += (double)sp.an_ls[c] + (double)sp.logS[c];
    ldc[0] = HW * s;                                 // H*W*(sum log_scale + sum log_S)
  }
  // inference-mode BatchNorm folded to y = g'*x + b'
  for (int i = t; i < F; i += blockDim.x) {
    const double g1 = (double)sp.bn1_gamma[i] / sqrt((double)sp.bn1_var[i] + kBnEpsD);
    sp.g1f[i] = (float)g1;
    sp.b1f[i] = (float)((double)sp.bn1_beta[i] - g1 * (double)sp.bn1_mean[i]);
    const double g2 = (double)sp.bn2_gamma[i] / sqrt((double)sp.bn2_var[i] + kBnEpsD);
    sp.g2f[i] = (float)g2;
    sp.b2f[i] = (float)((double)sp.bn2_beta[i] - g2 * (double)sp.bn2_mean[i]);
  }
  if (need_k2t)
    for (int idx = t; idx < F * F; idx += blockDim.x) {
      const int i = idx / F, j = idx % F;
      sp.k2t[(size_t)j * F + i] = sp.k2[idx];
    }
}

__global__ void k_sum_doubles(const double* __restrict__ v, int n, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += v[i];
    out[0] = s;
  }
}

// ------------------------------------------------------------------ Keras Adamax (train_utils.py:29-30)
__global__ void k_adamax(float* __restrict__ theta, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ u,
                         long long n, float lr_t, float b1, float b2, float eps) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float ui = fmaxf(b2 * u[i], fabsf(gi));
  m[i] = mi;
  u[i] = ui;
  theta[i] -= lr_t * mi / (ui + eps);
}


// Keras Adam (train_utils.py:27-28; OptimizerV2 non-amsgrad form): lr_t = lr sqrt(1 - b2^t) / (1 - b1^t) on the host
__global__ void k_adam(float* __restrict__ theta, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       long long n, float lr_t, float b1, float b2, float eps) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  theta[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

// ------------------------------------------------------------------ tensor-core training path helpers
// G9[q][tap*C+c] = gr[q - off(tap)][c] (0 outside the image), bf16, row stride ld (multiple of 64, zero padded)
__global__ void __launch_bounds__(256) k_im2col_gr(const float* __restrict__ gr, __nv_bfloat16* __restrict__ G9, int H, int W,
                                                   int C, int ld, long long M) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int groups = ld / C + 1;                       // taps 0..8 plus padding groups
  const long long q = idx / groups;
  const int g = (int)(idx % groups);
  if (q >= M) return;
  const int c0 = g * C;
  if (c0 >= ld) return;
  const int w = (int)(q % W), h = (int)((q / W) % H);
  bool ok = g < 9;
  long long src = 0;
  if (ok) {
    const int dy = g / 3 - 1, dx = g % 3 - 1;
    const int hh = h - dy, ww = w - dx;
    ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
    src = q - (long long)dy * W - dx;
  }
  for (int c = 0; c < C && c0 + c < ld; ++c)
    G9[q * ld + c0 + c] = __float2bfloat16_rn(ok ? gr[src * C + c] : 0.f);
}

// X9[p][tap*Ch+ci] = state[p + off(tap)][Ch+ci] (0 outside), bf16, row stride ld
__global__ void __launch_bounds__(256) k_im2col_xb(const float* __restrict__ state, __nv_bfloat16* __restrict__ X9, int H, int W,
                                                   int C, int ld, long long M) {
  const int Ch = C / 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int groups = ld / Ch + 1;
  const long long p = idx / groups;
  const int g = (int)(idx % groups);
  if (p >= M) return;
  const int c0 = g * Ch;
  if (c0 >= ld) return;
  const int w = (int)(p % W), h = (int)((p / W) % H);
  bool ok = g < 9;
  long long src = 0;
  if (ok) {
    const int dy = g / 3 - 1, dx = g % 3 - 1;
    const int hh = h + dy, ww = w + dx;
    ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
    src = p + (long long)dy * W + dx;
  }
  for (int c = 0; c < Ch && c0 + c < ld; ++c)
    X9[p * ld + c0 + c] = __float2bfloat16_rn(ok ? state[src * C + Ch + c] : 0.f);
}

// S3[tap][c] += sum_{p: p+off(tap) in bounds} gr[p][c].  One block per image row: the row's channel sums over all
// columns, the first column and the last column give every tap's contribution (dx = -1 drops w = 0, dx = +1 drops
// w = W-1; dy = -1 drops the row h = 0, dy = +1 the row h = H-1).
__global__ void __launch_bounds__(128) k_s3(const float* __restrict__ gr, float* __restrict__ S3, int H, int W, int C,
                                            int num_rows) {
  __shared__ float part[128];
  // threads: c = t % C, w-lane = t / C; a block walks image rows blockIdx.x, +gridDim.x, ... and keeps the nine
  // per-tap sums of its rows in registers (threads < C), so each block issues 9*C atomics in total
  const int c = threadIdx.x % C, wl = threadIdx.x / C, wstep = blockDim.x / C;
  float acc[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) acc[tap] = 0.f;
  for (int r = blockIdx.x; r < num_rows; r += gridDim.x) {
    const int h = r % H;
    const float* row = gr + (size_t)r * W * C;
    float s = 0.f;
    if (wl < wstep)
      for (int w = wl; w < W; w += wstep) s += row[(size_t)w * C + c];
    __syncthreads();
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < C) {
      float all = 0.f;
      for (int g = 0; g < wstep; ++g) all += part[g * C + threadIdx.x];
      const float first = row[threadIdx.x], last = row[(size_t)(W - 1) * C + threadIdx.x];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        if (h + dy < 0 || h + dy >= H) continue;
        acc[tap] += all - (dx == -1 ? first : 0.f) - (dx == 1 ? last : 0.f);
      }
    }
  }
  if (threadIdx.x < C) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) atomicAdd(S3 + tap * C + threadIdx.x, acc[tap]);
  }
}

__global__ void __launch_bounds__(512) k_colsum_bf16(const __nv_bfloat16* __restrict__ X, float* __restrict__ out, long long M,
                                                     int F, int rows_per_block) {
  const int c = threadIdx.x;
  if (c >= F) return;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += __bfloat162float(X[r * F + c]);
  atomicAdd(out + c, s);
}

// bf16 SWIZZLE_128B tile images + folded biases of one step, rebuilt on the device after a parameter update
// (device twin of nn_tc_prepare in nn_tc.cu; same image order and element mapping)
__device__ __forceinline__ size_t img_off(int r, int k) { return (size_t)r * 64 + (size_t)((((k >> 3) ^ (r & 7)) << 3) + (k & 7)); }

__global__ void __launch_bounds__(256) k_build_tc_images(const StepRefresh* __restrict__ table) {
  const StepRefresh& row = table[blockIdx.z];                      // grid.z = flow step
  const StepTrainPtrs sp = row.sp;
  __nv_bfloat16* fwd = row.fwd_img;
  __nv_bfloat16* bwd = row.bwd_img;
  const int k1p_f = row.k1p_f, n3p_f = row.n3p_f, k1p_b = row.k1p_b, n3p_b = row.n3p_b, f16 = row.f16;
  const int F = sp.F, C = sp.C, Ch = C / 2;
  const int dir = blockIdx.y;                                   // 0 forward set, 1 backward set
  const int k1p = dir == 0 ? k1p_f : k1p_b, n3p = dir == 0 ? n3p_f : n3p_b;
  const int K1h = dir == 0 ? 9 * Ch : 9 * C, K1 = 2 * K1h, N3 = dir == 0 ? 9 * C : 9 * Ch;
  __nv_bfloat16* dst = dir == 0 ? fwd : bwd;
  // work item = (image, 8-wide k chunk, row); rows vary fastest so that loads over n coalesce
  const long long n1 = (long long)2 * k1p * 8 * 256, n2 = (long long)16 * 8 * 256, n3 = (long long)8 * 8 * n3p;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n1 + n2 + n3; e += (long long)gridDim.x * blockDim.x) {
    float v[8];
    size_t off;
    if (e < n1) {
      const int img = (int)(e / (8 * 256)), ch = (int)((e / 256) % 8), r = (int)(e % 256);
      const int half = img / k1p, kp = img % k1p, n = half * 256 + r;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = kp * 64 + ch * 8 + q;
        float x = 0.f;
        if (k < K1) {
          const int kh = k % K1h;
          if (dir == 0) x = sp.k1[(size_t)kh * F + n];
          else { const int tap = kh / C, c = kh % C; x = sp.g2f[n] * sp.k3[((size_t)tap * F + n) * C + c]; }
        }
        v[q] = x;
      }
      off = (size_t)img * 256 * 64 + (size_t)r * 64 + (size_t)((ch ^ (r & 7)) << 3);
    } else if (e < n1 + n2) {
      const long long e2 = e - n1;
      const int img = (int)(e2 / (8 * 256)), ch = (int)((e2 / 256) % 8), r = (int)(e2 % 256);
      const int half = img / 8, kp = img % 8, n = half * 256 + r;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = kp * 64 + ch * 8 + q;
        v[q] = dir == 0 ? sp.g1f[k] * sp.k2[(size_t)k * F + n] : sp.g1f[n] * sp.k2[(size_t)n * F + k];
      }
      off = (size_t)n1 * 8 + (size_t)img * 256 * 64 + (size_t)r * 64 + (size_t)((ch ^ (r & 7)) << 3);
    } else {
      const long long e3 = e - n1 - n2;
      const int kp = (int)(e3 / (8LL * n3p)), ch = (int)((e3 / n3p) % 8), r = (int)(e3 % n3p);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = kp * 64 + ch * 8 + q;
        float x = 0.f;
        if (r < N3) {
          if (dir == 0) { const int tap = r / C, c = r % C; x = sp.g2f[k] * sp.k3[((size_t)tap * F + k) * C + c]; }
          else x = sp.k1[(size_t)r * F + k];
        }
        v[q] = x;
      }
      off = (size_t)(n1 + n2) * 8 + (size_t)kp * n3p * 64 + (size_t)r * 64 + (size_t)((ch ^ (r & 7)) << 3);
    }
    uint4 pk;
    if (f16 && dir == 0 && e >= n1) {      // ASEP_PREC_FP16: forward stage-2 / stage-3 weights are fp16 (nn_tc_prepare)
      __half2 h2;
      h2 = __floats2half2_rn(v[0], v[1]); pk.x = *reinterpret_cast<uint32_t*>(&h2);
      h2 = __floats2half2_rn(v[2], v[3]); pk.y = *reinterpret_cast<uint32_t*>(&h2);
      h2 = __floats2half2_rn(v[4], v[5]); pk.z = *reinterpret_cast<uint32_t*>(&h2);
      h2 = __floats2half2_rn(v[6], v[7]); pk.w = *reinterpret_cast<uint32_t*>(&h2);
    } else {
      __nv_bfloat162 t2;
      t2 = __floats2bfloat162_rn(v[0], v[1]); pk.x = *reinterpret_cast<uint32_t*>(&t2);
      t2 = __floats2bfloat162_rn(v[2], v[3]); pk.y = *reinterpret_cast<uint32_t*>(&t2);
      t2 = __floats2bfloat162_rn(v[4], v[5]); pk.z = *reinterpret_cast<uint32_t*>(&t2);
      t2 = __floats2bfloat162_rn(v[6], v[7]); pk.w = *reinterpret_cast<uint32_t*>(&t2);
    }
    *reinterpret_cast<uint4*>(dst + off) = pk;
  }
}

// bias1 = c1, bias2 = c2 + b1' K2, const3[tap][c] = sum_k b2'[k] K3[tap][k][c], c3.
// grid 16 blocks x 512 threads = (16 k-groups x 32 outputs): every thread reduces a 32-long k slice, the 16 slices are
// combined in a fixed order through shared memory (fp32; the host twin accumulates in double, difference ~1e-7 relative).
__global__ void __launch_bounds__(512) k_build_tc_biases(const StepRefresh* __restrict__ table) {
  const StepRefresh& row = table[blockIdx.y];                      // grid.y = flow step
  const StepTrainPtrs sp = row.sp;
  float* bias1 = row.bias1;
  float* bias2 = row.bias2;
  float* const3 = row.const3;
  float* c3 = row.c3;
  __shared__ float part[16][33];
  const int F = sp.F, C = sp.C, t = threadIdx.x;
  const int ol = t & 31, kg = t >> 5;
  {
    const int n = blockIdx.x * 32 + ol;                 // 16 blocks x 32 = 512 = F columns of K2
    float a = 0.f;
    if (n < F)
      for (int k = kg; k < F; k += 16) a = fmaf(sp.b1f[k], sp.k2[(size_t)k * F + n], a);
    part[kg][ol] = a;
    __syncthreads();
    if (kg == 0 && n < F) {
      float sum = 0.f;
#pragma unroll
      for (int g = 0; g < 16; ++g) sum += part[g][ol];
      bias2[n] = sp.c2[n] + sum;
      bias1[n] = sp.c1[n];
    }
    __syncthreads();
  }
  {
    const int i = blockIdx.x * 32 + ol;                 // 9*C <= 144 outputs: blocks 0..4
    float a = 0.f;
    if (i < 9 * C) {
      const int tap = i / C, c = i % C;
      for (int k = kg; k < F; k += 16) a = fmaf(sp.b2f[k], sp.k3[((size_t)tap * F + k) * C + c], a);
    }
    part[kg][ol] = a;
    __syncthreads();
    if (kg == 0 && i < 9 * C) {
      float sum = 0.f;
#pragma unroll
      for (int g = 0; g < 16; ++g) sum += part[g][ol];
      const3[i] = sum;
    }
  }
  if (blockIdx.x == 0 && t < C) c3[t] = sp.c3[t];
}

}  // namespace

void launch_axpy(const float* x, const float* n, float sigma, float* y, long long total, cudaStream_t s) {
  k_axpy<<<cdiv(total, 256), 256, 0, s>>>(x, n, sigma, y, total);
  ASEP_LAUNCH_CHECK();
}

void launch_colsum(const float* X, float* out, long long M, int F, cudaStream_t s) {
  ASEP_CHECK(F <= 512, ASEP_ERR_UNSUPPORTED, "colsum: F > 512");
  const int rows = 512;
  k_colsum<<<cdiv(M, rows), 512, 0, s>>>(X, out, M, F, rows);
  ASEP_LAUNCH_CHECK();
}

void launch_wgrad_tn(const float* A, const float* B, float* Q, long long M, int F, cudaStream_t s) {
  ASEP_CHECK(F % 64 == 0, ASEP_ERR_UNSUPPORTED, "wgrad: F %% 64 != 0");
  const int tiles = (F / 64) * (F / 64);
  int nsplit = (int)std::max<long long>(1, std::min<long long>((M + 255) / 256, (592 + tiles - 1) / tiles));
  const int rows = (int)(((M + nsplit - 1) / nsplit + 15) / 16 * 16);
  nsplit = (int)((M + rows - 1) / rows);
  dim3 grid(F / 64, F / 64, nsplit);
  k_wgrad_tn<<<grid, 256, 0, s>>>(A, B, Q, M, F, rows);
  ASEP_LAUNCH_CHECK();
}

void launch_wgrad_conv3(const float* a2, const float* gr, float* R3, float* S3, int N, int H, int W, int C, int F,
                        cudaStream_t s) {
  const long long M = (long long)N * H * W;
  ASEP_CHECK(F <= 512, ASEP_ERR_UNSUPPORTED, "wgrad conv3: F > 512");
  const int px = 256;
  const int threads = std::max(F, 32);
  const size_t smem = (size_t)px * C * sizeof(float);
  switch (C) {
    case 2: k_wgrad_conv3<2><<<cdiv(M, px), threads, smem, s>>>(a2, gr, R3, S3, H, W, F, M, px); break;
    case 4: k_wgrad_conv3<4><<<cdiv(M, px), threads, smem, s>>>(a2, gr, R3, S3, H, W, F, M, px); break;
    case 8: k_wgrad_conv3<8><<<cdiv(M, px), threads, smem, s>>>(a2, gr, R3, S3, H, W, F, M, px); break;
    case 16: k_wgrad_conv3<16><<<cdiv(M, px), threads, smem, s>>>(a2, gr, R3, S3, H, W, F, M, px); break;
    case 32: k_wgrad_conv3<32><<<cdiv(M, px), threads, smem, s>>>(a2, gr, R3, S3, H, W, F, M, px); break;
    default: throw Error(ASEP_ERR_UNSUPPORTED, strfmt("wgrad conv3: channel count %d not built", C));
  }
  ASEP_LAUNCH_CHECK();
}

void launch_wgrad_conv1(const float* state, const float* gp1, float* dK1, float* dc1, int N, int H, int W, int C, int F,
                        float gs, cudaStream_t s) {
  const long long M = (long long)N * H * W;
  ASEP_CHECK(F <= 512, ASEP_ERR_UNSUPPORTED, "wgrad conv1: F > 512");
  const int px = 256;
  const int threads = std::max(F, 32);
  switch (C / 2) {
    case 1: k_wgrad_conv1<1><<<cdiv(M, px), threads, 0, s>>>(state, gp1, dK1, dc1, H, W, F, M, px, gs); break;
    case 2: k_wgrad_conv1<2><<<cdiv(M, px), threads, 0, s>>>(state, gp1, dK1, dc1, H, W, F, M, px, gs); break;
    case 4: k_wgrad_conv1<4><<<cdiv(M, px), threads, 0, s>>>(state, gp1, dK1, dc1, H, W, F, M, px, gs); break;
    case 8: k_wgrad_conv1<8><<<cdiv(M, px), threads, 0, s>>>(state, gp1, dK1, dc1, H, W, F, M, px, gs); break;
    case 16: k_wgrad_conv1<16><<<cdiv(M, px), threads, 0, s>>>(state, gp1, dK1, dc1, H, W, F, M, px, gs); break;
    default: throw Error(ASEP_ERR_UNSUPPORTED, strfmt("wgrad conv1: channel count %d not built", C));
  }
  ASEP_LAUNCH_CHECK();
}

void launch_step_stats(const float* gu, const float* gxb, const float* u, const float* sc, double* stats, long long M,
                       int C, cudaStream_t s) {
  switch (C) {
    case 2: k_step_stats<2><<<cdiv(M, 256), 256, 0, s>>>(gu, gxb, u, sc, stats, M); break;
    case 4: k_step_stats<4><<<cdiv(M, 256), 256, 0, s>>>(gu, gxb, u, sc, stats, M); break;
    case 8: k_step_stats<8><<<cdiv(M, 256), 256, 0, s>>>(gu, gxb, u, sc, stats, M); break;
    case 16: k_step_stats<16><<<cdiv(M, 128), 128, 0, s>>>(gu, gxb, u, sc, stats, M); break;
    default: throw Error(ASEP_ERR_UNSUPPORTED, strfmt("step stats: channel count %d not built", C));
  }
  ASEP_LAUNCH_CHECK();
}

void launch_finalize_step(const StepTrainPtrs& sp, const float* Q2, const float* dc2, const float* R3, const float* S3,
                          const double* stats, float* grads, double Mpix, float gs, cudaStream_t s) {
  ASEP_CHECK(sp.C * sp.C <= 256, ASEP_ERR_UNSUPPORTED, "finalize: C > 16");
  k_fin_conv2<<<sp.F, 512, 0, s>>>(sp, Q2, dc2, grads, gs);
  ASEP_LAUNCH_CHECK();
  k_fin_conv3<<<sp.F, 256, 0, s>>>(sp, R3, sp.F * sp.C, sp.C, S3, nullptr, 0, nullptr, grads, gs);
  ASEP_LAUNCH_CHECK();
  k_fin_small<<<1, 256, 0, s>>>(sp, S3, stats, grads, Mpix, gs);
  ASEP_LAUNCH_CHECK();
}

void launch_finalize_step_tc(const StepTrainPtrs& sp, const float* Q2, const float* dc2, const float* R3t, int r3_ld,
                             const float* S3, const float* D1t, int d1_ld, const float* dc1, const double* stats,
                             float* grads, double Mpix, float gs, cudaStream_t s) {
  ASEP_CHECK(sp.C * sp.C <= 256, ASEP_ERR_UNSUPPORTED, "finalize: C > 16");
  // R3t[k][tap*C + c]
  k_fin_conv2<<<sp.F, 512, 0, s>>>(sp, Q2, dc2, grads, gs);
  ASEP_LAUNCH_CHECK();
  k_fin_conv3<<<sp.F, 256, 0, s>>>(sp, R3t, sp.C, r3_ld, S3, D1t, d1_ld, dc1, grads, gs);
  ASEP_LAUNCH_CHECK();
  k_fin_small<<<1, 256, 0, s>>>(sp, S3, stats, grads, Mpix, gs);
  ASEP_LAUNCH_CHECK();
}

void launch_prior_grads(const float* z, const float* loc, const float* ls, float* gloc, float* gls, int N, int D, float gs,
                        cudaStream_t s) {
  k_prior_grads<<<cdiv(D, 256), 256, 0, s>>>(z, loc, ls, gloc, gls, N, D, gs);
  ASEP_LAUNCH_CHECK();
}

void launch_loss(const double* acc_ld, const double* acc_prior, const double* cst, double extra_const, int N,
                 double inv_batch, float* loss, cudaStream_t s) {
  k_loss<<<1, 256, 0, s>>>(acc_ld, acc_prior, cst, extra_const, N, inv_batch, loss);
  ASEP_LAUNCH_CHECK();
}

void launch_derive_all(const StepRefresh* table, int n_steps, double* ldc, int need_k2t, cudaStream_t s) {
  k_derive_step<<<n_steps, 512, 0, s>>>(table, ldc, need_k2t);
  ASEP_LAUNCH_CHECK();
}

void launch_sum_doubles(const double* v, int n, double* out, cudaStream_t s) {
  k_sum_doubles<<<1, 32, 0, s>>>(v, n, out);
  ASEP_LAUNCH_CHECK();
}

void launch_adamax(float* theta, const float* g, float* m, float* u, long long n, float lr_t, float b1, float b2, float eps,
                   cudaStream_t s) {
  k_adamax<<<cdiv(n, 256), 256, 0, s>>>(theta, g, m, u, n, lr_t, b1, b2, eps);
  ASEP_LAUNCH_CHECK();
}

void launch_adam(float* theta, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2, float eps,
                 cudaStream_t s) {
  k_adam<<<cdiv(n, 256), 256, 0, s>>>(theta, g, m, v, n, lr_t, b1, b2, eps);
  ASEP_LAUNCH_CHECK();
}

void launch_im2col_gr(const float* gr, __nv_bfloat16* G9, int N, int H, int W, int C, int ld, cudaStream_t s) {
  const long long M = (long long)N * H * W;
  const long long total = M * (ld / C + 1);
  k_im2col_gr<<<cdiv(total, 256), 256, 0, s>>>(gr, G9, H, W, C, ld, M);
  ASEP_LAUNCH_CHECK();
}

void launch_im2col_xb(const float* state, __nv_bfloat16* X9, int N, int H, int W, int C, int ld, cudaStream_t s) {
  const long long M = (long long)N * H * W;
  const long long total = M * (ld / (C / 2) + 1);
  k_im2col_xb<<<cdiv(total, 256), 256, 0, s>>>(state, X9, H, W, C, ld, M);
  ASEP_LAUNCH_CHECK();
}

void launch_s3(const float* gr, float* S3, int N, int H, int W, int C, cudaStream_t s) {
  ASEP_CHECK(C <= 128 && 128 % C == 0, ASEP_ERR_UNSUPPORTED, "s3: channel count %d", C);
  k_s3<<<std::min(N * H, 296), 128, 0, s>>>(gr, S3, H, W, C, N * H);
  ASEP_LAUNCH_CHECK();
}

void launch_colsum_bf16(const __nv_bfloat16* X, float* out, long long M, int F, cudaStream_t s) {
  ASEP_CHECK(F <= 512, ASEP_ERR_UNSUPPORTED, "colsum: F > 512");
  const int rows = M >= 148 * 2 * 64 ? 128 : 32;
  k_colsum_bf16<<<cdiv(M, rows), 512, 0, s>>>(X, out, M, F, rows);
  ASEP_LAUNCH_CHECK();
}

void launch_build_tc_all(const StepRefresh* table, int n_steps, cudaStream_t s) {
  dim3 grid(48, 2, n_steps);                                       // ~37-45 k work items per (step, direction), grid-stride
  k_build_tc_images<<<grid, 256, 0, s>>>(table);
  ASEP_LAUNCH_CHECK();
  k_build_tc_biases<<<dim3(16, n_steps), 512, 0, s>>>(table);       // F = 512 (checked by nn_tc_prepare)
  ASEP_LAUNCH_CHECK();
}

}  // namespace asep
