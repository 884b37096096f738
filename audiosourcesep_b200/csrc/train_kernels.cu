// Parameter-gradient, derived-constant and optimiser kernels of the Glow training step
// (reference: train_glow.py:29-44 loss / tape.gradient / apply_gradients, train_utils.py:23-41 Adamax,
// train_noisy_glow.py:30-33 noise perturbation).  The data-gradient sweep is the one grad_log_prob uses; these
// kernels add, per flow step, the weight gradients of the three convolutions (CUDA-core fp32 "exact" mode in this
// round: split-K outer products with fp32 atomics), the folded-BatchNorm chain rule, the ActNorm / LU-parameterised
// 1x1 gradients and the prior gradients, and refresh every per-step constant on the device after an update.
#include "train_kernels.h"

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
__global__ void __launch_bounds__(512) k_finalize_step(const StepTrainPtrs sp, const float* __restrict__ Q2,
                                                       const float* __restrict__ dc2, const float* __restrict__ R3,
                                                       const float* __restrict__ S3, const double* __restrict__ stats,
                                                       float* __restrict__ grads, double Mpix, float gs) {
  const int F = sp.F, C = sp.C, t = threadIdx.x;
  __shared__ double sP[256], sL[256], sU[256], sT[256], sD[256];
  // ---- conv2 + BN1
  for (int i = t; i < F; i += blockDim.x) {
    const float g1 = sp.g1f[i], b1 = sp.b1f[i];
    float dg = 0.f, db = 0.f;
    for (int o = 0; o < F; ++o) {
      const float q = Q2[(size_t)i * F + o], k = sp.k2[(size_t)i * F + o], d2 = dc2[o];
      grads[sp.o_k2 + (size_t)i * F + o] = gs * (g1 * q + b1 * d2);
      dg = fmaf(q, k, dg);
      db = fmaf(d2, k, db);
    }
    const float s = sqrtf(sp.bn1_var[i] + (float)kBnEpsD);
    grads[sp.o_bn1_gamma + i] = gs * (dg / s - db * sp.bn1_mean[i] / s);
    grads[sp.o_bn1_beta + i] = gs * db;
    grads[sp.o_c2 + i] = gs * dc2[i];
  }
  // ---- conv3 + BN2
  for (int k = t; k < F; k += blockDim.x) {
    const float g2 = sp.g2f[k], b2 = sp.b2f[k];
    float dg = 0.f, db = 0.f;
    for (int tap = 0; tap < 9; ++tap)
      for (int c = 0; c < C; ++c) {
        const size_t idx = ((size_t)tap * F + k) * C + c;
        const float r = R3[idx], s3 = S3[tap * C + c], kk = sp.k3[idx];
        grads[sp.o_k3 + idx] = gs * (g2 * r + b2 * s3);
        dg = fmaf(r, kk, dg);
        db = fmaf(s3, kk, db);
      }
    const float s = sqrtf(sp.bn2_var[k] + (float)kBnEpsD);
    grads[sp.o_bn2_gamma + k] = gs * (dg / s - db * sp.bn2_mean[k] / s);
    grads[sp.o_bn2_beta + k] = gs * db;
  }
  if (t < C) grads[sp.o_c3 + t] = gs * S3[4 * C + t];            // centre tap is always in bounds: sum_p gr[p][c]
  // ---- ActNorm
  if (t < C) {
    grads[sp.o_an_shift + t] = gs * (float)stats[t];
    grads[sp.o_an_ls + t] = gs * (float)(stats[C + t] + Mpix);   // + d(H*W*sum log_scale)/d log_scale per sample
  }
  // ---- 1x1: W = P L' U',  dL' = P^T dW U'^T,  dU' = (P L')^T dW
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
  double dL = 0.0, dU = 0.0;
  if (t < C * C) {
    const int i = t / C, j = t % C;
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
__global__ void __launch_bounds__(512) k_derive_step(const StepTrainPtrs sp, double HW, double* __restrict__ ldc) {
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
  k_finalize_step<<<1, 512, 0, s>>>(sp, Q2, dc2, R3, S3, stats, grads, Mpix, gs);
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

void launch_derive_step(const StepTrainPtrs& sp, double HW, double* ldc, cudaStream_t s) {
  ASEP_CHECK(sp.C * sp.C <= 256, ASEP_ERR_UNSUPPORTED, "derive: C > 16");
  k_derive_step<<<1, 512, 0, s>>>(sp, HW, ldc);
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

}  // namespace asep
