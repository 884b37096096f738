// BSS Eval v4 on the device (reference: bsseval_v4.py).  The reference computes every correlation with zero-padded
// FFTs of the whole signals (:480-507, :535-545) and only keeps the lags |l| < filters_len; here those lags are
// summed directly in float64 (same numbers up to round-off, no FFT), the block-Toeplitz normal equations (G + eps I) C = D
// (:551-556) are solved by Gaussian elimination with partial pivoting (what np.linalg.solve / LAPACK dgesv does), and
// the four components of :427-435 are never materialised: a sample's projections are formed in registers and only the
// seven energy sums of _bss_crit (:570-595) leave the kernel.  Mono images only (nchan = 1).
#include "bsseval.h"

#include <cmath>
#include <vector>

namespace asep {

namespace {

__device__ __forceinline__ double warp_sum_dd(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double block_sum_dd(double v, double* sh) {
  v = warp_sum_dd(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x + 31) / 32 ? sh[threadIdx.x] : 0.0;
    t = warp_sum_dd(t);
  }
  return t;            // valid in warp 0
}

// R[a][b][l + L - 1] = sum_n x_a[n + l] * x_b[n] over n, n + l in [f0, f1), for |l| < L.  grid (2L-1, pairs), block 256.
__global__ void __launch_bounds__(256) k_xcorr(const double* __restrict__ xa, const double* __restrict__ xb, long long stride_a,
                                               long long stride_b, int na, int nb, long long f0, long long f1, int L,
                                               double* __restrict__ R) {
  __shared__ double sh[32];
  const int l = (int)blockIdx.x - (L - 1);
  const int a = blockIdx.y / nb, b = blockIdx.y % nb;
  const double* pa = xa + (size_t)a * stride_a;
  const double* pb = xb + (size_t)b * stride_b;
  const long long lo = max(f0, f0 - l), hi = min(f1, f1 - l);     // n in [lo, hi): both n and n + l inside [f0, f1)
  double acc = 0.0;
  for (long long n = lo + threadIdx.x; n < hi; n += blockDim.x) acc = fma(pa[n + l], pb[n], acc);
  acc = block_sum_dd(acc, sh);
  if (threadIdx.x == 0) R[((size_t)a * nb + b) * (2 * L - 1) + blockIdx.x] = acc;
}

// augmented system of the joint projection: rows (j, p), columns (i, q) then one right-hand side per estimate:
// G[(j,p)][(i,q)] = R_ss[j][i][q - p] (+ eps on the diagonal), D[(j,p)][e] = R_se[j][e][-p]   (bsseval_v4.py:497-507,538-549)
__global__ void k_build_system(const double* __restrict__ Rss, const double* __restrict__ Rse, int nsrc_rows, int src0, int nsrc_all,
                               int nest, int L, double eps, double* __restrict__ A) {
  const int n = nsrc_rows * L, ld = n + nest;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * ld) return;
  const int r = (int)(i / ld), c = (int)(i % ld);
  const int j = src0 + r / L, p = r % L;
  double v;
  if (c < n) {
    const int ii = src0 + c / L, q = c % L;
    v = Rss[((size_t)j * nsrc_all + ii) * (2 * L - 1) + (q - p + L - 1)];
    if (r == c) v += eps;
  } else {
    const int e = c - n;
    v = Rse[((size_t)j * nest + e) * (2 * L - 1) + (-p + L - 1)];
  }
  A[i] = v;
}

// ---- Gaussian elimination with partial pivoting on the augmented matrix A [n x ld] (row-major)
__global__ void __launch_bounds__(1024) k_lu_pivot(double* __restrict__ A, int n, int ld, int k) {
  __shared__ double sv[32];
  __shared__ int si[32];
  __shared__ int piv;
  double best = -1.0;
  int arg = k;
  for (int i = k + threadIdx.x; i < n; i += blockDim.x) {
    const double v = fabs(A[(size_t)i * ld + k]);
    if (v > best) { best = v; arg = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = arg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x + 31) / 32; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < arg)) { best = sv[w]; arg = si[w]; }
    piv = arg;
  }
  __syncthreads();
  const int p = piv;
  if (p != k)
    for (int j = threadIdx.x; j < ld; j += blockDim.x) {
      const double t = A[(size_t)k * ld + j];
      A[(size_t)k * ld + j] = A[(size_t)p * ld + j];
      A[(size_t)p * ld + j] = t;
    }
  __syncthreads();
  const double d = A[(size_t)k * ld + k];
  for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x) A[(size_t)i * ld + k] /= d;      // multipliers, in place
}

__global__ void __launch_bounds__(256) k_lu_update(double* __restrict__ A, int n, int ld, int k) {
  const int j = k + 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int i0 = k + 1 + blockIdx.y * 8;
  if (j >= ld) return;
  const double u = A[(size_t)k * ld + j];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int i = i0 + t;
    if (i < n) A[(size_t)i * ld + j] = fma(-A[(size_t)i * ld + k], u, A[(size_t)i * ld + j]);
  }
}

// back substitution of every right-hand side: grid nrhs, block 256; X [nrhs, n]
__global__ void __launch_bounds__(256) k_lu_backsolve(const double* __restrict__ A, int n, int ld, double* __restrict__ X) {
  extern __shared__ double xs[];          // [n]
  __shared__ double sh[32];
  const int e = blockIdx.x;
  for (int i = n - 1; i >= 0; --i) {
    double acc = 0.0;
    for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) acc = fma(A[(size_t)i * ld + j], xs[j], acc);
    acc = block_sum_dd(acc, sh);
    if (threadIdx.x == 0) xs[i] = (A[(size_t)i * ld + n + e] - acc) / A[(size_t)i * ld + i];
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) X[(size_t)e * n + i] = xs[i];
}

// energies of the decomposition (bsseval_v4.py:419-435, :570-595) of estimate `jest` against true source `jtrue` on the
// window [w0, w1): sample m in [0, len + L - 1) of the zero-padded window.
//   Pj   = sum_k Cj[jtrue][jest][k] ref_jtrue[m - k]         (= s_true + e_spat)
//   Pall = sum_i sum_k C[jest][i][k] ref_i[m - k]            (= s_true + e_spat + e_interf)
// acc[7] += (s_true^2, (est - s_true)^2, e_spat^2, Pj^2, e_interf^2, Pall^2, e_artif^2)
__global__ void __launch_bounds__(256) k_decomp(const double* __restrict__ refs, const double* __restrict__ ests, long long nsampl,
                                                int nsrc, int L, const double* __restrict__ C, const double* __restrict__ Cj,
                                                long long w0, long long w1, int jtrue, int jest, double* __restrict__ acc) {
  extern __shared__ double filt[];       // [nsrc * L] joint filters of this estimate, then [L] the single-source filter
  __shared__ double sh[32];
  double* fj = filt + (size_t)nsrc * L;
  for (int i = threadIdx.x; i < nsrc * L; i += blockDim.x) filt[i] = C[(size_t)jest * nsrc * L + i];
  for (int i = threadIdx.x; i < L; i += blockDim.x) fj[i] = Cj[((size_t)jtrue * nsrc + jest) * L + i];
  __syncthreads();
  const long long len = w1 - w0;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double e[7] = {0, 0, 0, 0, 0, 0, 0};
  if (m < len + L - 1) {
    double pj = 0.0, pall = 0.0;
    const int k0 = (int)max(0LL, m - (len - 1)), k1 = (int)min((long long)L - 1, m);
    for (int i = 0; i < nsrc; ++i) {
      const double* r = refs + (size_t)i * nsampl + w0;
      double a = 0.0;
      for (int k = k0; k <= k1; ++k) a = fma(filt[i * L + k], r[m - k], a);
      pall += a;
    }
    {
      const double* r = refs + (size_t)jtrue * nsampl + w0;
      for (int k = k0; k <= k1; ++k) pj = fma(fj[k], r[m - k], pj);
    }
    const double st = m < len ? refs[(size_t)jtrue * nsampl + w0 + m] : 0.0;
    const double es = m < len ? ests[(size_t)jest * nsampl + w0 + m] : 0.0;
    const double e_spat = pj - st, e_interf = pall - pj, e_artif = es - pall;
    e[0] = st * st; e[1] = (es - st) * (es - st); e[2] = e_spat * e_spat; e[3] = pj * pj;
    e[4] = e_interf * e_interf; e[5] = pall * pall; e[6] = e_artif * e_artif;
  }
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    const double t = block_sum_dd(e[q], sh);
    if (threadIdx.x == 0) atomicAdd(acc + q, t);
  }
}

__device__ __forceinline__ double safe_db(double num, double den) { return den == 0.0 ? INFINITY : 10.0 * log10(num / den); }

// out[(m, jtrue, jest, t)] from the energy sums (bsseval_v4.py:570-595); NaN when a source is silent in the window (:258-279)
__global__ void k_criteria(const double* __restrict__ acc, const int* __restrict__ silent, int nsrc, int nwin, int sources_version,
                           double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsrc * nsrc * nwin) return;
  const int t = i % nwin;
  const double* e = acc + (size_t)i * 7;
  double sdr, isr, sir, sar;
  if (silent[t]) {
    sdr = isr = sir = sar = NAN;
  } else if (sources_version) {
    // s_filt = s_true + e_spat; sdr vs (e_interf + e_artif) = est - s_filt
    sdr = NAN; isr = NAN; sir = safe_db(e[3], e[4]); sar = safe_db(e[5], e[6]);
  } else {
    sdr = safe_db(e[0], e[1]); isr = safe_db(e[0], e[2]); sir = safe_db(e[3], e[4]); sar = safe_db(e[5], e[6]);
  }
  const size_t n = (size_t)nsrc * nsrc * nwin;
  out[i] = sdr; out[n + i] = isr; out[2 * n + i] = sir; out[3 * n + i] = sar;
}

// (est - s_filt)^2 for the bss_eval_sources SDR: needs its own sum (s_filt = Pj)
__global__ void __launch_bounds__(256) k_decomp_srcver(const double* __restrict__ refs, const double* __restrict__ ests,
                                                       long long nsampl, int nsrc, int L, const double* __restrict__ Cj, long long w0,
                                                       long long w1, int jtrue, int jest, double* __restrict__ acc) {
  extern __shared__ double filt[];
  __shared__ double sh[32];
  for (int i = threadIdx.x; i < L; i += blockDim.x) filt[i] = Cj[((size_t)jtrue * nsrc + jest) * L + i];
  __syncthreads();
  const long long len = w1 - w0;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (m < len + L - 1) {
    double pj = 0.0;
    const int k0 = (int)max(0LL, m - (len - 1)), k1 = (int)min((long long)L - 1, m);
    const double* r = refs + (size_t)jtrue * nsampl + w0;
    for (int k = k0; k <= k1; ++k) pj = fma(filt[k], r[m - k], pj);
    const double es = m < len ? ests[(size_t)jest * nsampl + w0 + m] : 0.0;
    v = (es - pj) * (es - pj);
  }
  const double t = block_sum_dd(v, sh);
  if (threadIdx.x == 0) atomicAdd(acc, t);
}

__global__ void k_srcver_sdr(const double* __restrict__ acc, const double* __restrict__ extra, const int* __restrict__ silent, int nsrc,
                             int nwin, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsrc * nsrc * nwin) return;
  if (silent[i % nwin]) return;
  out[i] = safe_db(acc[(size_t)i * 7 + 3], extra[i]);
}

// silent[t] = any source (reference or estimate) identically zero on window t  (bsseval_v4.py:73-76, :258-261)
__global__ void __launch_bounds__(256) k_silent(const double* __restrict__ refs, const double* __restrict__ ests, long long nsampl,
                                                int nsrc, long long w0, long long w1, int* __restrict__ nonzero) {
  const int sig = blockIdx.y;             // 0..2*nsrc-1
  const double* p = (sig < nsrc ? refs + (size_t)sig * nsampl : ests + (size_t)(sig - nsrc) * nsampl);
  // the reference tests sum(sources, axis=channels) == 0 for ALL samples; with one channel: every sample zero
  bool nz = false;
  for (long long n = w0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; n < w1; n += (long long)gridDim.x * blockDim.x) nz = nz || p[n] != 0.0;
  if (__syncthreads_or(nz) && threadIdx.x == 0) atomicOr(nonzero + sig, 1);
}
__global__ void k_silent_fin(const int* __restrict__ nonzero, int nsig, int* __restrict__ silent) {
  int s = 0;
  for (int i = 0; i < nsig; ++i) s |= nonzero[i] == 0;
  silent[0] = s;
}

// ---- ideal masks on mel spectrograms (oracle_systems.py:264-350)
__global__ void __launch_bounds__(256) k_ideal_mask(const float* __restrict__ mixture, const float* __restrict__ sources,
                                                    float* __restrict__ est, int nsrc, long long P, int binary, float theta) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const double eps = 2.220446049250313e-16;        // np.finfo(np.float).eps
  const double mix = (double)mixture[i];
  if (binary) {                                    // IBM_melspec :300-311: mask = source / (eps + mixture) >= theta
    for (int j = 0; j < nsrc; ++j) {
      const double m = (double)sources[(size_t)j * P + i] / (eps + mix);
      est[(size_t)j * P + i] = m >= (double)theta ? (float)mix : 0.f;
    }
  } else {                                         // IRM_melspec :338-348: mask = source / (sum sources + eps)
    double model = eps;
    for (int j = 0; j < nsrc; ++j) model += (double)sources[(size_t)j * P + i];
    for (int j = 0; j < nsrc; ++j) est[(size_t)j * P + i] = (float)(mix * ((double)sources[(size_t)j * P + i] / model));
  }
}

void lu_solve(double* A, int n, int nrhs, double* X, cudaStream_t s) {
  const int ld = n + nrhs;
  for (int k = 0; k < n; ++k) {
    k_lu_pivot<<<1, 1024, 0, s>>>(A, n, ld, k);
    ASEP_LAUNCH_CHECK();
    const int rows = n - k - 1, cols = ld - k - 1;
    if (rows > 0 && cols > 0) {
      dim3 grid(cdiv(cols, 256), cdiv(rows, 8));
      k_lu_update<<<grid, 256, 0, s>>>(A, n, ld, k);
      ASEP_LAUNCH_CHECK();
    }
  }
  k_lu_backsolve<<<nrhs, 256, (size_t)n * sizeof(double), s>>>(A, n, ld, X);
  ASEP_LAUNCH_CHECK();
}

}  // namespace

void bss_eval_core(const double* refs, const double* ests, int nsrc, long long nsampl, int L, long long f0, long long f1,
                   const long long* win0, const long long* win1, int nwin, int sources_version, double* out, cudaStream_t s) {
  ASEP_CHECK(nsrc >= 1 && nsrc <= 8 && L >= 1 && L <= 1024 && nsrc * L <= 4096, ASEP_ERR_UNSUPPORTED,
             "bss_eval: nsrc = %d, filters_len = %d", nsrc, L);
  ASEP_CHECK(0 <= f0 && f0 < f1 && f1 <= nsampl, ASEP_ERR_BAD_ARG, "bss_eval: filter range [%lld, %lld) of %lld", f0, f1, nsampl);
  const int n = nsrc * L, nl = 2 * L - 1;
  const double eps = 2.220446049250313e-16;
  for (int t = 0; t < nwin; ++t)
    ASEP_CHECK(0 <= win0[t] && win0[t] < win1[t] && win1[t] <= nsampl, ASEP_ERR_BAD_ARG, "bss_eval: window [%lld, %lld) of %lld",
               win0[t], win1[t], nsampl);
  // stream-ordered scratch, released on every exit path
  struct Scratch {
    cudaStream_t s;
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFreeAsync(p, s); }
    void* get(size_t bytes) {
      void* p = nullptr;
      CUDA_CHECK(cudaMallocAsync(&p, bytes, s));
      ptrs.push_back(p);
      return p;
    }
  } scratch{s, {}};
  double* Rss = static_cast<double*>(scratch.get((size_t)nsrc * nsrc * nl * sizeof(double)));
  double* Rse = static_cast<double*>(scratch.get((size_t)nsrc * nsrc * nl * sizeof(double)));
  double* A = static_cast<double*>(scratch.get((size_t)n * (n + nsrc) * sizeof(double)));
  double* C = static_cast<double*>(scratch.get((size_t)nsrc * n * sizeof(double)));
  double* Aj = static_cast<double*>(scratch.get((size_t)L * (L + nsrc) * sizeof(double)));
  double* Cj = static_cast<double*>(scratch.get((size_t)nsrc * nsrc * L * sizeof(double)));
  double* acc = static_cast<double*>(scratch.get((size_t)nsrc * nsrc * nwin * 7 * sizeof(double)));
  double* extra = static_cast<double*>(scratch.get((size_t)nsrc * nsrc * nwin * sizeof(double)));
  int* nonzero = static_cast<int*>(scratch.get((size_t)2 * nsrc * sizeof(int)));
  int* silent = static_cast<int*>(scratch.get((size_t)nwin * sizeof(int)));
  CUDA_CHECK(cudaMemsetAsync(acc, 0, (size_t)nsrc * nsrc * nwin * 7 * sizeof(double), s));
  CUDA_CHECK(cudaMemsetAsync(extra, 0, (size_t)nsrc * nsrc * nwin * sizeof(double), s));
  // correlations at the lags |l| < L: references x references, references x estimates
  k_xcorr<<<dim3(nl, nsrc * nsrc), 256, 0, s>>>(refs, refs, nsampl, nsampl, nsrc, nsrc, f0, f1, L, Rss);
  ASEP_LAUNCH_CHECK();
  k_xcorr<<<dim3(nl, nsrc * nsrc), 256, 0, s>>>(refs, ests, nsampl, nsampl, nsrc, nsrc, f0, f1, L, Rse);
  ASEP_LAUNCH_CHECK();
  // joint projection filters C[jest][i][k]  (compute_GsfC)
  k_build_system<<<cdiv((long long)n * (n + nsrc), 256), 256, 0, s>>>(Rss, Rse, nsrc, 0, nsrc, nsrc, L, eps, A);
  ASEP_LAUNCH_CHECK();
  lu_solve(A, n, nsrc, C, s);
  // single-source filters Cj[jtrue][jest][k]  (compute_Cj)
  for (int j = 0; j < nsrc; ++j) {
    k_build_system<<<cdiv((long long)L * (L + nsrc), 256), 256, 0, s>>>(Rss, Rse, 1, j, nsrc, nsrc, L, eps, Aj);
    ASEP_LAUNCH_CHECK();
    // lu_solve returns X [nrhs = jest][L]; store at Cj[j][jest][:]
    lu_solve(Aj, L, nsrc, Cj + (size_t)j * nsrc * L, s);
  }
  for (int t = 0; t < nwin; ++t) {
    const long long w0 = win0[t], w1 = win1[t];
    CUDA_CHECK(cudaMemsetAsync(nonzero, 0, (size_t)2 * nsrc * sizeof(int), s));
    k_silent<<<dim3(64, 2 * nsrc), 256, 0, s>>>(refs, ests, nsampl, nsrc, w0, w1, nonzero);
    ASEP_LAUNCH_CHECK();
    k_silent_fin<<<1, 1, 0, s>>>(nonzero, 2 * nsrc, silent + t);
    ASEP_LAUNCH_CHECK();
    const long long len = w1 - w0 + L - 1;
    for (int jt = 0; jt < nsrc; ++jt)
      for (int je = 0; je < nsrc; ++je) {
        double* a = acc + (((size_t)jt * nsrc + je) * nwin + t) * 7;
        k_decomp<<<cdiv(len, 256), 256, (size_t)(nsrc + 1) * L * sizeof(double), s>>>(refs, ests, nsampl, nsrc, L, C, Cj, w0, w1, jt, je, a);
        ASEP_LAUNCH_CHECK();
        if (sources_version) {
          k_decomp_srcver<<<cdiv(len, 256), 256, (size_t)L * sizeof(double), s>>>(refs, ests, nsampl, nsrc, L, Cj, w0, w1, jt, je,
                                                                                  extra + ((size_t)jt * nsrc + je) * nwin + t);
          ASEP_LAUNCH_CHECK();
        }
      }
  }
  k_criteria<<<cdiv(nsrc * nsrc * nwin, 128), 128, 0, s>>>(acc, silent, nsrc, nwin, sources_version, out);
  ASEP_LAUNCH_CHECK();
  if (sources_version) {
    k_srcver_sdr<<<cdiv(nsrc * nsrc * nwin, 128), 128, 0, s>>>(acc, extra, silent, nsrc, nwin, out);
    ASEP_LAUNCH_CHECK();
  }
}

void launch_ideal_mask(const float* mixture, const float* sources, float* estimates, int nsrc, long long P, int binary,
                       float theta, cudaStream_t s) {
  if (P == 0) return;
  k_ideal_mask<<<cdiv(P, 256), 256, 0, s>>>(mixture, sources, estimates, nsrc, P, binary, theta);
  ASEP_LAUNCH_CHECK();
}

}  // namespace asep
