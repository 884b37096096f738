// Backward / training kernels of the NCSN score networks (HBM-bound; the three GEMMs of every convolution run on
// tcgen05: conv_tc.cu forward and data gradient, conv_wgrad_tc.cu weight gradient).  Reference: train_ncsn.py:26-57
// (denoising score matching), ncsn/score_network.py:7-221, ncsn/score_network_v2.py:6-199 (the layers whose chain rule
// is written out here; TensorFlow's GradientTape derives it automatically).  All tensors fp32 NHWC unless stated.
#include "ncsn_train_kernels.h"

namespace asep {

namespace {

constexpr float kInEps = 1e-3f;     // tfa.InstanceNormalization epsilon
constexpr double kPlusEps = 1e-5;   // score_network.py:205

__device__ __forceinline__ float elu_grad(float pre) { return pre > 0.f ? 1.f : expf(pre); }
__device__ __forceinline__ float warp_sum_f(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float bf16_pair(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long i) {
  float v = __bfloat162float(hi[i]);
  if (lo) v += __bfloat162float(lo[i]);
  return v;
}

// ------------------------------------------------------------------ denoising score matching
__global__ void __launch_bounds__(256) k_dsm_perturb(const float* __restrict__ x, const float* __restrict__ noise,
                                                     const float* __restrict__ sigmas, const int* __restrict__ idx,
                                                     float* __restrict__ xt, int HW, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  xt[i] = x[i] + sigmas[idx[i / HW]] * noise[i];
}

__global__ void __launch_bounds__(256) k_dsm_loss(const float* __restrict__ score, const float* __restrict__ noise,
                                                  const float* __restrict__ sigmas, const int* __restrict__ idx,
                                                  float* __restrict__ gscore, double* __restrict__ loss, int HW,
                                                  long long total, double inv_b) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double l = 0.0;
  if (i < total) {
    const float sg = sigmas[idx[i / HW]];
    const float d = score[i] + noise[i] / sg;                 // score - target, target = -(sigma z) / sigma^2
    gscore[i] = (float)((double)d * sg * sg * inv_b);
    l = 0.5 * (double)d * d * sg * sg * inv_b;
  }
  l = warp_sum_d(l);
  if ((threadIdx.x & 31) == 0 && l != 0.0) atomicAdd(loss, l);
}

__global__ void k_double_to_float(const double* __restrict__ src, float* __restrict__ dst) { dst[0] = (float)src[0]; }

// ------------------------------------------------------------------ end_conv (C -> 1) backward
__global__ void __launch_bounds__(256) k_end_conv_bwd_data(const float* __restrict__ gs, const float* __restrict__ k,
                                                           const float* __restrict__ sigmas, const int* __restrict__ idx,
                                                           float* __restrict__ gin, int H, int W, int C4, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C4);
  const long long p = i / C4;
  const int w = (int)(p % W), h = (int)((p / W) % H);
  const float inv = sigmas ? 1.f / sigmas[idx[p / ((long long)H * W)]] : 1.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    // forward: y[q] += k[tap] . x[q + off(tap)]  =>  x[p] feeds y[p - off(tap)]
    const int hh = h - (tap / 3 - 1), ww = w - (tap % 3 - 1);
    if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
    const float g = __ldg(gs + p - (long long)(tap / 3 - 1) * W - (tap % 3 - 1)) * inv;
    const float4 kk = __ldg(reinterpret_cast<const float4*>(k) + tap * C4 + c);
    acc.x = fmaf(g, kk.x, acc.x); acc.y = fmaf(g, kk.y, acc.y); acc.z = fmaf(g, kk.z, acc.z); acc.w = fmaf(g, kk.w, acc.w);
  }
  reinterpret_cast<float4*>(gin)[i] = acc;
}

// dk[tap][c] += sum_q gs'[q] * x[q + off(tap)][c];  block = slab of pixels, thread = (4 channels, pixel lane)
__global__ void __launch_bounds__(256) k_end_conv_bwd_w(const float* __restrict__ gs, const __nv_bfloat16* __restrict__ x,
                                                        const __nv_bfloat16* __restrict__ x_lo, const float* __restrict__ sigmas,
                                                        const int* __restrict__ idx, float* __restrict__ dk,
                                                        float* __restrict__ dbias, int H, int W, int C, long long pixels,
                                                        int pix_per_block) {
  extern __shared__ float sh[];                       // [9 * C] + 1
  const int C4 = C / 4;
  for (int i = threadIdx.x; i < 9 * C + 1; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int rstep = blockDim.x / C4;
  const int c = threadIdx.x % C4, r0 = threadIdx.x / C4;
  float acc[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
  float bsum = 0.f;
  if (r0 < rstep) {
    const long long q0 = (long long)blockIdx.x * pix_per_block, q1 = min(pixels, q0 + pix_per_block);
    for (long long q = q0 + r0; q < q1; q += rstep) {
      const int w = (int)(q % W), h = (int)((q / W) % H);
      float g = __ldg(gs + q);
      if (sigmas) g /= sigmas[idx[q / ((long long)H * W)]];
      if (c == 0) bsum += g;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
        const long long off = (q + (long long)(tap / 3 - 1) * W + (tap % 3 - 1)) * C + 4 * c;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[tap][j] = fmaf(g, bf16_pair(x, x_lo, off + j), acc[tap][j]);
      }
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&sh[tap * C + 4 * c + j], acc[tap][j]);
    if (c == 0) atomicAdd(&sh[9 * C], bsum);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) atomicAdd(dk + i, sh[i]);
  if (threadIdx.x == 0) atomicAdd(dbias, sh[9 * C]);
}

// ------------------------------------------------------------------ begin_conv (1 -> Cout) backward (weights only)
__global__ void __launch_bounds__(256) k_begin_conv_bwd_w(const float* __restrict__ x, const float* __restrict__ gout,
                                                          float* __restrict__ dk, float* __restrict__ dbias, int H, int W,
                                                          int C, int rescale, long long pixels, int pix_per_block) {
  extern __shared__ float sh[];                       // [10 * C]
  const int C4 = C / 4;
  for (int i = threadIdx.x; i < 10 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int rstep = blockDim.x / C4;
  const int c = threadIdx.x % C4, r0 = threadIdx.x / C4;
  float acc[10][4];
#pragma unroll
  for (int t = 0; t < 10; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
  if (r0 < rstep) {
    const long long p0 = (long long)blockIdx.x * pix_per_block, p1 = min(pixels, p0 + pix_per_block);
    for (long long p = p0 + r0; p < p1; p += rstep) {
      const int w = (int)(p % W), h = (int)((p / W) % H);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gout) + p * C4 + c);
      acc[9][0] += g.x; acc[9][1] += g.y; acc[9][2] += g.z; acc[9][3] += g.w;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
        float v = __ldg(x + p + (long long)(tap / 3 - 1) * W + (tap % 3 - 1));
        if (rescale) v = 2.f * v - 1.f;
        acc[tap][0] = fmaf(v, g.x, acc[tap][0]); acc[tap][1] = fmaf(v, g.y, acc[tap][1]);
        acc[tap][2] = fmaf(v, g.z, acc[tap][2]); acc[tap][3] = fmaf(v, g.w, acc[tap][3]);
      }
    }
#pragma unroll
    for (int t = 0; t < 10; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&sh[t * C + 4 * c + j], acc[t][j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) atomicAdd(dk + i, sh[i]);
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dbias + i, sh[9 * C + i]);
}

// ------------------------------------------------------------------ normalise (+ELU) + cast backward
// red[n][c] = (sum g', sum g' * xv), g' = gy * act'(a*xv + b).  grid (row slabs, N); thread = 4 channels, strided rows.
__global__ void __launch_bounds__(256) k_prep_bwd_reduce(const float* __restrict__ xv, const float* __restrict__ gy,
                                                         const float2* __restrict__ coef, int do_elu, double* __restrict__ red,
                                                         int HW, int C4, int rows_per_block) {
  __shared__ float4 sm[2][256];
  const int n = blockIdx.y;
  const int rstep = blockDim.x / C4;
  const int c = threadIdx.x % C4, r0 = threadIdx.x / C4;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  if (r0 < rstep) {
    float a[4] = {1.f, 1.f, 1.f, 1.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if (coef) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 t = __ldg(coef + ((size_t)n * C4 + c) * 4 + j); a[j] = t.x; b[j] = t.y; }
    }
    const int rbeg = blockIdx.x * rows_per_block, rend = min(HW, rbeg + rows_per_block);
    const float4* xb = reinterpret_cast<const float4*>(xv) + ((size_t)n * HW) * C4 + c;
    const float4* gb = reinterpret_cast<const float4*>(gy) + ((size_t)n * HW) * C4 + c;
    for (int r = rbeg + r0; r < rend; r += rstep) {
      const float4 x = __ldg(xb + (size_t)r * C4);
      float4 g = __ldg(gb + (size_t)r * C4);
      if (do_elu) {
        g.x *= elu_grad(fmaf(a[0], x.x, b[0])); g.y *= elu_grad(fmaf(a[1], x.y, b[1]));
        g.z *= elu_grad(fmaf(a[2], x.z, b[2])); g.w *= elu_grad(fmaf(a[3], x.w, b[3]));
      }
      s1.x += g.x; s1.y += g.y; s1.z += g.z; s1.w += g.w;
      s2.x = fmaf(g.x, x.x, s2.x); s2.y = fmaf(g.y, x.y, s2.y); s2.z = fmaf(g.z, x.z, s2.z); s2.w = fmaf(g.w, x.w, s2.w);
    }
  }
  sm[0][threadIdx.x] = s1;
  sm[1][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.x < C4) {
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f), v = u;
    for (int g = 0; g < rstep; ++g) {
      const float4 t = sm[0][g * C4 + threadIdx.x], w = sm[1][g * C4 + threadIdx.x];
      u.x += t.x; u.y += t.y; u.z += t.z; u.w += t.w;
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    double* st = red + ((size_t)n * C4 + threadIdx.x) * 8;
    atomicAdd(st + 0, (double)u.x); atomicAdd(st + 1, (double)v.x);
    atomicAdd(st + 2, (double)u.y); atomicAdd(st + 3, (double)v.y);
    atomicAdd(st + 4, (double)u.z); atomicAdd(st + 5, (double)v.z);
    atomicAdd(st + 6, (double)u.w); atomicAdd(st + 7, (double)v.w);
  }
}

// block-wide sums of two doubles (blockDim <= 512)
__device__ __forceinline__ void block_sum2(double& a, double& b, double (*sh)[16], double* out) {
  a = warp_sum_d(a);
  b = warp_sum_d(b);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int i = 0; i < (int)(blockDim.x + 31) / 32; ++i) { ta += sh[0][i]; tb += sh[1][i]; }
    out[0] = ta;
    out[1] = tb;
  }
  __syncthreads();
  a = out[0];
  b = out[1];
}

// out = A x + B per (n,c) with A = gamma gin r, B = gamma (bin - gin mu r) + alpha mu~ + beta, r = (var + eps)^-1/2,
// mu~ = (mu - m) / sqrt(v + eps') over the channels of one sample.  S1 = d loss / dB, S2 = d loss / dA.
__global__ void __launch_bounds__(512) k_norm_bwd_coef(const double* __restrict__ red, const double* __restrict__ sums,
                                                       const float* __restrict__ gamma, const float* __restrict__ alpha,
                                                       const float* __restrict__ beta, int stride_n, const int* __restrict__ idx,
                                                       const float* __restrict__ in_gamma, const float* __restrict__ in_beta,
                                                       float* __restrict__ d_gamma, float* __restrict__ d_alpha,
                                                       float* __restrict__ d_beta, float* __restrict__ d_in_gamma,
                                                       float* __restrict__ d_in_beta, float2* __restrict__ qr, int HW, int C) {
  __shared__ double sh[2][16];
  __shared__ double res[2];
  const int n = blockIdx.x, c = threadIdx.x;
  const bool on = c < C;
  double mu = 0.0, var = 0.0, S1 = 0.0, S2 = 0.0;
  if (on) {
    mu = sums[((size_t)n * C + c) * 2] / HW;
    var = sums[((size_t)n * C + c) * 2 + 1] / HW - mu * mu;
    if (var < 0.0) var = 0.0;
    S1 = red[((size_t)n * C + c) * 2];
    S2 = red[((size_t)n * C + c) * 2 + 1];
  }
  double a = on ? mu : 0.0, b = on ? mu * mu : 0.0;
  block_sum2(a, b, sh, res);
  const double m = a / C;
  double vv = b / C - m * m;
  if (vv < 0.0) vv = 0.0;
  const double rs = 1.0 / sqrt(vv + kPlusEps);
  const double mt = (mu - m) * rs;
  const size_t row = (size_t)(idx ? idx[n] : 0) * stride_n;
  double g = 0.0, al = 0.0, gi = 0.0, bi = 0.0;
  if (on) { g = gamma[row + c]; al = alpha[row + c]; gi = in_gamma[c]; bi = in_beta[c]; }
  const double r = 1.0 / sqrt(var + (double)kInEps);
  const double q = on ? al * S1 : 0.0;               // d loss / d mu~_c
  double qa = q, qb = q * mt;
  block_sum2(qa, qb, sh, res);
  if (!on) return;
  const double dmu_t = rs * (q - qa / C - mt * (qb / C));
  const double cen = S2 - mu * S1;                   // sum g' (x - mu)
  atomicAdd(d_gamma + row + c, (float)(gi * r * cen + S1 * bi));
  atomicAdd(d_alpha + row + c, (float)(S1 * mt));
  atomicAdd(d_beta + row + c, (float)S1);
  atomicAdd(d_in_gamma + c, (float)(g * r * cen));
  atomicAdd(d_in_beta + c, (float)(g * S1));
  const double dv = -0.5 * r * r * r * g * gi * cen;
  const double dmu = -S1 * g * gi * r + dmu_t;
  const double Q = 2.0 * dv / HW;
  qr[(size_t)n * C + c] = make_float2((float)Q, (float)(dmu / HW - Q * mu));
}

__global__ void __launch_bounds__(256) k_prep_bwd_apply(const float* __restrict__ x, const float* __restrict__ gy,
                                                        const float2* __restrict__ coef, int do_elu,
                                                        const float2* __restrict__ qr, float* __restrict__ gx, int HW, int C4,
                                                        int rows_per_block) {
  const int n = blockIdx.y;
  const int rstep = blockDim.x / C4;
  const int c = threadIdx.x % C4, r0 = threadIdx.x / C4;
  if (r0 >= rstep) return;
  float a[4] = {1.f, 1.f, 1.f, 1.f}, b[4] = {0.f, 0.f, 0.f, 0.f}, Q[4] = {0.f, 0.f, 0.f, 0.f}, R[4] = {0.f, 0.f, 0.f, 0.f};
  if (coef) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = __ldg(coef + ((size_t)n * C4 + c) * 4 + j); a[j] = t.x; b[j] = t.y; }
  }
  if (qr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = __ldg(qr + ((size_t)n * C4 + c) * 4 + j); Q[j] = t.x; R[j] = t.y; }
  }
  const int rbeg = blockIdx.x * rows_per_block, rend = min(HW, rbeg + rows_per_block);
  const size_t base = ((size_t)n * HW) * C4 + c;
  for (int r = rbeg + r0; r < rend; r += rstep) {
    const size_t i = base + (size_t)r * C4;
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
    float4 o = reinterpret_cast<float4*>(gx)[i];
    if (gy) {
      float4 g = __ldg(reinterpret_cast<const float4*>(gy) + i);
      if (do_elu) {
        g.x *= elu_grad(fmaf(a[0], xv.x, b[0])); g.y *= elu_grad(fmaf(a[1], xv.y, b[1]));
        g.z *= elu_grad(fmaf(a[2], xv.z, b[2])); g.w *= elu_grad(fmaf(a[3], xv.w, b[3]));
      }
      o.x = fmaf(a[0], g.x, o.x); o.y = fmaf(a[1], g.y, o.y); o.z = fmaf(a[2], g.z, o.z); o.w = fmaf(a[3], g.w, o.w);
    }
    if (qr) {
      o.x += fmaf(Q[0], xv.x, R[0]); o.y += fmaf(Q[1], xv.y, R[1]); o.z += fmaf(Q[2], xv.z, R[2]); o.w += fmaf(Q[3], xv.w, R[3]);
    }
    reinterpret_cast<float4*>(gx)[i] = o;
  }
}

// ------------------------------------------------------------------ element-wise / stencil backward
__global__ void __launch_bounds__(256) k_axpy(const float* __restrict__ g, float* __restrict__ dst, long long nvec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(g) + i);
  float4 d = reinterpret_cast<float4*>(dst)[i];
  d.x += a.x; d.y += a.y; d.z += a.z; d.w += a.w;
  reinterpret_cast<float4*>(dst)[i] = d;
}

__global__ void __launch_bounds__(256) k_elu_bwd(const float* __restrict__ x, const float* __restrict__ gy,
                                                 float* __restrict__ gx, long long nvec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i), g = __ldg(reinterpret_cast<const float4*>(gy) + i);
  float4 d = reinterpret_cast<float4*>(gx)[i];
  d.x = fmaf(g.x, elu_grad(v.x), d.x); d.y = fmaf(g.y, elu_grad(v.y), d.y);
  d.z = fmaf(g.z, elu_grad(v.z), d.z); d.w = fmaf(g.w, elu_grad(v.w), d.w);
  reinterpret_cast<float4*>(gx)[i] = d;
}

// gin [N,2H,2W,C] += 0.25 * gout [N,H,W,C]; thread = one float4 of gin
__global__ void __launch_bounds__(256) k_avgpool2_bwd(const float* __restrict__ gout, float* __restrict__ gin, int H, int W,
                                                      int C4, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C4);
  long long p = i / C4;
  const int w = (int)(p % (2 * W)), h = (int)((p / (2 * W)) % (2 * H));
  const long long img = p / ((long long)4 * W * H);
  const float4 g = __ldg(reinterpret_cast<const float4*>(gout) + ((img * H + h / 2) * W + w / 2) * C4 + c);
  float4 d = reinterpret_cast<float4*>(gin)[i];
  d.x = fmaf(0.25f, g.x, d.x); d.y = fmaf(0.25f, g.y, d.y); d.z = fmaf(0.25f, g.z, d.z); d.w = fmaf(0.25f, g.w, d.w);
  reinterpret_cast<float4*>(gin)[i] = d;
}

// 5x5 'same' average pooling backward, separable.  kAxis = 0: tmp[p] = sum_{dw} gout[p+dw] / count(p+dw);
// kAxis = 1: gin[p] += sum_{dh} tmp[p + dh W]   (zero outside the image)
template <int kAxis>
__global__ void __launch_bounds__(256) k_pool5_avg_bwd(const float* __restrict__ src, float* __restrict__ dst, int H, int W,
                                                       int C4, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long p = i / C4;
  const int w = (int)(p % W), h = (int)((p / W) % H);
  const float4* base = reinterpret_cast<const float4*>(src) + i;
  const int pos = kAxis == 0 ? w : h, lim = kAxis == 0 ? W : H;
  const long long step = kAxis == 0 ? (long long)C4 : (long long)W * C4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int d = -2; d <= 2; ++d) {
    if (pos + d < 0 || pos + d >= lim) continue;
    float4 v = __ldg(base + d * step);
    if (kAxis == 0) {
      const int ww = w + d;
      const int ch = min(h + 2, H - 1) - max(h - 2, 0) + 1, cw = min(ww + 2, W - 1) - max(ww - 2, 0) + 1;
      const float r = 1.f / (float)(ch * cw);
      v.x *= r; v.y *= r; v.z *= r; v.w *= r;
    }
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (kAxis == 1) {
    const float4 o = reinterpret_cast<float4*>(dst)[i];
    acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
  }
  reinterpret_cast<float4*>(dst)[i] = acc;
}

// 5x5 'same' max pooling backward: thread = (output pixel, 4 channels); first maximum in row-major window order
__global__ void __launch_bounds__(256) k_pool5_max_bwd(const float* __restrict__ x, const float* __restrict__ gout,
                                                       float* __restrict__ gin, int H, int W, int C4, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C4);
  const long long p = i / C4;
  const int w = (int)(p % W), h = (int)((p / W) % H);
  const float4* base = reinterpret_cast<const float4*>(x) + i;
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  long long arg[4] = {0, 0, 0, 0};
  for (int dh = -2; dh <= 2; ++dh) {
    if (h + dh < 0 || h + dh >= H) continue;
    for (int dw = -2; dw <= 2; ++dw) {
      if (w + dw < 0 || w + dw >= W) continue;
      const long long off = ((long long)dh * W + dw) * C4;
      const float4 v = __ldg(base + off);
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (vv[j] > best[j]) { best[j] = vv[j]; arg[j] = off; }
    }
  }
  const float4 g = __ldg(reinterpret_cast<const float4*>(gout) + i);
  const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) atomicAdd(gin + (i + arg[j]) * 4 + j, gg[j]);
  (void)c;
}

// transpose of k_resize2x_add's interpolation: thread = (output pixel, 4 channels), scatter to its four sources
__global__ void __launch_bounds__(256) k_resize2x_bwd(const float* __restrict__ gout, float* __restrict__ glow, int h, int w,
                                                      int C4, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int H = 2 * h, W = 2 * w;
  const int c = (int)(i % C4);
  long long p = i / C4;
  const int ox = (int)(p % W), oy = (int)((p / W) % H);
  const long long img = p / ((long long)W * H);
  const float sy = fmaxf(0.f, (oy + 0.5f) * 0.5f - 0.5f), sx = fmaxf(0.f, (ox + 0.5f) * 0.5f - 0.5f);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float fy = sy - y0, fx = sx - x0;
  const float4 g = __ldg(reinterpret_cast<const float4*>(gout) + i);
  float* base = glow + (img * h * w * C4 + c) * 4;
  const float wt[4] = {(1.f - fy) * (1.f - fx), (1.f - fy) * fx, fy * (1.f - fx), fy * fx};
  const long long off[4] = {((long long)y0 * w + x0) * C4 * 4, ((long long)y0 * w + x1) * C4 * 4,
                            ((long long)y1 * w + x0) * C4 * 4, ((long long)y1 * w + x1) * C4 * 4};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (wt[t] == 0.f) continue;
    atomicAdd(base + off[t] + 0, wt[t] * g.x); atomicAdd(base + off[t] + 1, wt[t] * g.y);
    atomicAdd(base + off[t] + 2, wt[t] * g.z); atomicAdd(base + off[t] + 3, wt[t] * g.w);
  }
}

// dbias[c] += sum_p g[p][c]
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ g, float* __restrict__ dbias, long long P, int C4,
                                                int rows_per_block) {
  __shared__ float4 sm[256];
  const int rstep = blockDim.x / C4;
  const int c = threadIdx.x % C4, r0 = threadIdx.x / C4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r0 < rstep) {
    const long long rbeg = (long long)blockIdx.x * rows_per_block, rend = min(P, rbeg + rows_per_block);
    for (long long r = rbeg + r0; r < rend; r += rstep) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g) + r * C4 + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < C4) {
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < rstep; ++k) { const float4 t = sm[k * C4 + threadIdx.x]; u.x += t.x; u.y += t.y; u.z += t.z; u.w += t.w; }
    atomicAdd(dbias + 4 * threadIdx.x + 0, u.x); atomicAdd(dbias + 4 * threadIdx.x + 1, u.y);
    atomicAdd(dbias + 4 * threadIdx.x + 2, u.z); atomicAdd(dbias + 4 * threadIdx.x + 3, u.w);
  }
}

// ------------------------------------------------------------------ tile images from the fp32 master kernel
__global__ void __launch_bounds__(256) k_build_conv_image(const float* __restrict__ kernel, __nv_bfloat16* __restrict__ img,
                                                          int taps, int Cin, int Cout, int transposed, int lo,
                                                          long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ci_img = transposed ? Cout : Cin, co_img = transposed ? Cin : Cout;   // contraction / output-row extents of the image
  const int kpanels = ci_img / 64;
  const int k = (int)(i % 64);
  long long t = i / 64;
  const int r = (int)(t % co_img);
  t /= co_img;
  const int kp = (int)(t % kpanels), tap = (int)(t / kpanels);
  float v;
  if (!transposed) v = kernel[((size_t)tap * Cin + kp * 64 + k) * Cout + r];
  else v = kernel[((size_t)(taps - 1 - tap) * Cin + r) * Cout + kp * 64 + k];
  const __nv_bfloat16 hi = __float2bfloat16(v);
  const __nv_bfloat16 o = lo ? __float2bfloat16(v - __bfloat162float(hi)) : hi;
  const size_t base = ((size_t)tap * kpanels + kp) * co_img * 64;
  img[base + (size_t)r * 64 + (size_t)((((k >> 3) ^ (r & 7)) << 3) + (k & 7))] = o;
}

int slab_rows(int HW, int N) {
  // enough (slab, n) blocks to fill the GPU a few times over, at least 32 rows per block
  int slabs = std::max(1, std::min(HW / 32, (148 * 8 + N - 1) / N));
  return (HW + slabs - 1) / slabs;
}

}  // namespace

void launch_dsm_perturb(const float* x, const float* noise, const float* sigmas, const int* idx, float* xt, int N, int HW,
                        cudaStream_t s) {
  const long long total = (long long)N * HW;
  k_dsm_perturb<<<cdiv(total, 256), 256, 0, s>>>(x, noise, sigmas, idx, xt, HW, total);
  ASEP_LAUNCH_CHECK();
}

void launch_dsm_loss(const float* score, const float* noise, const float* sigmas, const int* idx, float* gscore, double* loss,
                     int N, int HW, double inv_global_batch, cudaStream_t s) {
  const long long total = (long long)N * HW;
  k_dsm_loss<<<cdiv(total, 256), 256, 0, s>>>(score, noise, sigmas, idx, gscore, loss, HW, total, inv_global_batch);
  ASEP_LAUNCH_CHECK();
}

void launch_double_to_float(const double* src, float* dst, cudaStream_t s) {
  k_double_to_float<<<1, 1, 0, s>>>(src, dst);
  ASEP_LAUNCH_CHECK();
}

void launch_end_conv_bwd_data(const float* gs, const float* k, const float* sigmas, const int* idx, float* gin, int N, int H,
                              int W, int C, cudaStream_t s) {
  const long long total = (long long)N * H * W * (C / 4);
  k_end_conv_bwd_data<<<cdiv(total, 256), 256, 0, s>>>(gs, k, sigmas, idx, gin, H, W, C / 4, total);
  ASEP_LAUNCH_CHECK();
}

void launch_end_conv_bwd_w(const float* gs, const __nv_bfloat16* x, const __nv_bfloat16* x_lo, const float* sigmas,
                           const int* idx, float* dk, float* dbias, int N, int H, int W, int C, cudaStream_t s) {
  ASEP_CHECK(C % 4 == 0 && C / 4 <= 256, ASEP_ERR_UNSUPPORTED, "end_conv backward: C = %d", C);
  const long long pixels = (long long)N * H * W;
  const int ppb = (int)std::max<long long>(64, (pixels + 591) / 592);
  k_end_conv_bwd_w<<<cdiv(pixels, ppb), 256, (9 * C + 1) * sizeof(float), s>>>(gs, x, x_lo, sigmas, idx, dk, dbias, H, W, C,
                                                                               pixels, ppb);
  ASEP_LAUNCH_CHECK();
}

void launch_begin_conv_bwd_w(const float* x, const float* gout, float* dk, float* dbias, int N, int H, int W, int Cout,
                             int rescale, cudaStream_t s) {
  ASEP_CHECK(Cout % 4 == 0 && Cout / 4 <= 256, ASEP_ERR_UNSUPPORTED, "begin_conv backward: Cout = %d", Cout);
  const long long pixels = (long long)N * H * W;
  const int ppb = (int)std::max<long long>(64, (pixels + 591) / 592);
  k_begin_conv_bwd_w<<<cdiv(pixels, ppb), 256, 10 * Cout * sizeof(float), s>>>(x, gout, dk, dbias, H, W, Cout, rescale, pixels,
                                                                               ppb);
  ASEP_LAUNCH_CHECK();
}

void launch_prep_bwd_reduce(const float* xv, const float* gy, const float2* coef, int do_elu, double* red, int N, int HW, int C,
                            cudaStream_t s) {
  ASEP_CHECK(C % 4 == 0 && C / 4 <= 256, ASEP_ERR_UNSUPPORTED, "prep backward: C = %d", C);
  CUDA_CHECK(cudaMemsetAsync(red, 0, (size_t)N * C * 2 * sizeof(double), s));
  const int rpb = slab_rows(HW, N);
  dim3 grid(cdiv(HW, rpb), N);
  k_prep_bwd_reduce<<<grid, 256, 0, s>>>(xv, gy, coef, do_elu, red, HW, C / 4, rpb);
  ASEP_LAUNCH_CHECK();
}

void launch_norm_bwd_coef(const double* red, const double* sums, const float* gamma, const float* alpha, const float* beta,
                          int stride_n, const int* idx, const float* in_gamma, const float* in_beta, float* d_gamma,
                          float* d_alpha, float* d_beta, float* d_in_gamma, float* d_in_beta, float2* qr, int N, int HW, int C,
                          cudaStream_t s) {
  ASEP_CHECK(C <= 512, ASEP_ERR_UNSUPPORTED, "norm backward: C = %d > 512", C);
  const int threads = ((C + 31) / 32) * 32;
  k_norm_bwd_coef<<<N, threads, 0, s>>>(red, sums, gamma, alpha, beta, stride_n, idx, in_gamma, in_beta, d_gamma, d_alpha,
                                         d_beta, d_in_gamma, d_in_beta, qr, HW, C);
  ASEP_LAUNCH_CHECK();
}

void launch_prep_bwd_apply(const float* x, const float* gy, const float2* coef, int do_elu, const float2* qr, float* gx, int N,
                           int HW, int C, cudaStream_t s) {
  const int rpb = slab_rows(HW, N);
  dim3 grid(cdiv(HW, rpb), N);
  k_prep_bwd_apply<<<grid, 256, 0, s>>>(x, gy, coef, do_elu, qr, gx, HW, C / 4, rpb);
  ASEP_LAUNCH_CHECK();
}

void launch_axpy(const float* g, float* dst, long long n, cudaStream_t s) {
  k_axpy<<<cdiv(n / 4, 256), 256, 0, s>>>(g, dst, n / 4);
  ASEP_LAUNCH_CHECK();
}

void launch_elu_bwd(const float* x, const float* gy, float* gx, long long n, cudaStream_t s) {
  k_elu_bwd<<<cdiv(n / 4, 256), 256, 0, s>>>(x, gy, gx, n / 4);
  ASEP_LAUNCH_CHECK();
}

void launch_avgpool2_bwd(const float* gout, float* gin, int N, int Hout, int Wout, int C, cudaStream_t s) {
  const long long total = (long long)N * 4 * Hout * Wout * (C / 4);
  k_avgpool2_bwd<<<cdiv(total, 256), 256, 0, s>>>(gout, gin, Hout, Wout, C / 4, total);
  ASEP_LAUNCH_CHECK();
}

void launch_pool5_avg_bwd(const float* gout, float* tmp, float* gin, int N, int H, int W, int C, cudaStream_t s) {
  const long long total = (long long)N * H * W * (C / 4);
  k_pool5_avg_bwd<0><<<cdiv(total, 256), 256, 0, s>>>(gout, tmp, H, W, C / 4, total);
  ASEP_LAUNCH_CHECK();
  k_pool5_avg_bwd<1><<<cdiv(total, 256), 256, 0, s>>>(tmp, gin, H, W, C / 4, total);
  ASEP_LAUNCH_CHECK();
}

void launch_pool5_max_bwd(const float* x, const float* gout, float* gin, int N, int H, int W, int C, cudaStream_t s) {
  const long long total = (long long)N * H * W * (C / 4);
  k_pool5_max_bwd<<<cdiv(total, 256), 256, 0, s>>>(x, gout, gin, H, W, C / 4, total);
  ASEP_LAUNCH_CHECK();
}

void launch_resize2x_bwd(const float* gout, float* glow, int N, int h, int w, int C, cudaStream_t s) {
  const long long total = (long long)N * 4 * h * w * (C / 4);
  k_resize2x_bwd<<<cdiv(total, 256), 256, 0, s>>>(gout, glow, h, w, C / 4, total);
  ASEP_LAUNCH_CHECK();
}

void launch_colsum_f32(const float* g, float* dbias, long long P, int C, cudaStream_t s) {
  ASEP_CHECK(C % 4 == 0 && C / 4 <= 256, ASEP_ERR_UNSUPPORTED, "colsum: C = %d", C);
  const int rpb = (int)std::max<long long>(64, (P + 591) / 592);
  k_colsum<<<cdiv(P, rpb), 256, 0, s>>>(g, dbias, P, C / 4, rpb);
  ASEP_LAUNCH_CHECK();
}

void launch_build_conv_image(const float* kernel, __nv_bfloat16* img, int taps, int Cin, int Cout, int transposed, int lo,
                             cudaStream_t s) {
  const long long total = (long long)taps * Cin * Cout;
  k_build_conv_image<<<cdiv(total, 256), 256, 0, s>>>(kernel, img, taps, Cin, Cout, transposed, lo, total);
  ASEP_LAUNCH_CHECK();
}

}  // namespace asep
