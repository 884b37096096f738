// HBM-bound kernels of the NCSN score networks: conditional instance-norm++ statistics and coefficients,
// normalise + ELU + bf16 cast (the producer of every tcgen05 convolution operand), 5x5 'same' average / max
// pooling, 2x2 average pooling, bilinear x2 up-sampling, the 1-channel begin convolution and the 1-channel end
// convolution.  Reference: ncsn/score_network.py:7-221, ncsn/score_network_v2.py:6-199 (layer semantics restated
// in oracle/ncsn_oracle.py).  All tensors NHWC; channels are the fastest index so every warp access is coalesced.
#include "ncsn_kernels.h"

namespace asep {

namespace {

constexpr float kInEps = 1e-3f;     // tfa.InstanceNormalization epsilon
constexpr double kPlusEps = 1e-5;   // score_network.py:205

__device__ __forceinline__ float elu(float v) { return v > 0.f ? v : expm1f(v); }

// ---- per-(n,c) sum and sum of squares over H*W.  grid (row slabs, N); a thread owns 4 channels (float4 loads) and
// every (blockDim / (C/4))-th pixel of the slab, four loads in flight.
__global__ void __launch_bounds__(256) k_in_stats(const float* __restrict__ x, double* __restrict__ sums, int HW, int C4,
                                                  int rows_per_block) {
  __shared__ float4 red[2][256];
  const int n = blockIdx.y;
  const int rstep = blockDim.x / C4;
  const int c = threadIdx.x % C4, r0 = threadIdx.x / C4;
  const int rbeg = blockIdx.x * rows_per_block, rend = min(HW, rbeg + rows_per_block);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = s;
  if (r0 < rstep) {
    const float4* base = reinterpret_cast<const float4*>(x) + ((size_t)n * HW) * C4 + c;
    int r = rbeg + r0;
    for (; r + 3 * rstep < rend; r += 4 * rstep) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(base + (size_t)(r + u * rstep) * C4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w;
        ss.x = fmaf(v[u].x, v[u].x, ss.x); ss.y = fmaf(v[u].y, v[u].y, ss.y);
        ss.z = fmaf(v[u].z, v[u].z, ss.z); ss.w = fmaf(v[u].w, v[u].w, ss.w);
      }
    }
    for (; r < rend; r += rstep) {
      const float4 v = __ldg(base + (size_t)r * C4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      ss.x = fmaf(v.x, v.x, ss.x); ss.y = fmaf(v.y, v.y, ss.y); ss.z = fmaf(v.z, v.z, ss.z); ss.w = fmaf(v.w, v.w, ss.w);
    }
  }
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.x < C4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    for (int g = 0; g < rstep; ++g) {
      const float4 t = red[0][g * C4 + threadIdx.x], u = red[1][g * C4 + threadIdx.x];
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      b.x += u.x; b.y += u.y; b.z += u.z; b.w += u.w;
    }
    double* st = sums + ((size_t)n * C4 + threadIdx.x) * 8;
    atomicAdd(st + 0, (double)a.x); atomicAdd(st + 1, (double)b.x);
    atomicAdd(st + 2, (double)a.y); atomicAdd(st + 3, (double)b.y);
    atomicAdd(st + 4, (double)a.z); atomicAdd(st + 5, (double)b.z);
    atomicAdd(st + 6, (double)a.w); atomicAdd(st + 7, (double)b.w);
  }
}

// ---- y = elu(x) together with the per-(n,c) sum / sum of squares of y (the CRP block normalises its ELU output next):
// same slab structure as k_in_stats, one read of x instead of two passes
__global__ void __launch_bounds__(256) k_elu_stats(const float* __restrict__ x, float* __restrict__ y, double* __restrict__ sums,
                                                   int HW, int C4, int rows_per_block) {
  __shared__ float4 red[2][256];
  const int n = blockIdx.y;
  const int rstep = blockDim.x / C4;
  const int c = threadIdx.x % C4, r0 = threadIdx.x / C4;
  const int rbeg = blockIdx.x * rows_per_block, rend = min(HW, rbeg + rows_per_block);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = s;
  if (r0 < rstep) {
    const size_t base = ((size_t)n * HW) * C4 + c;
    for (int r = rbeg + r0; r < rend; r += rstep) {
      float4 v = __ldg(reinterpret_cast<const float4*>(x) + base + (size_t)r * C4);
      v.x = elu(v.x); v.y = elu(v.y); v.z = elu(v.z); v.w = elu(v.w);
      reinterpret_cast<float4*>(y)[base + (size_t)r * C4] = v;
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      ss.x = fmaf(v.x, v.x, ss.x); ss.y = fmaf(v.y, v.y, ss.y); ss.z = fmaf(v.z, v.z, ss.z); ss.w = fmaf(v.w, v.w, ss.w);
    }
  }
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.x < C4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    for (int g = 0; g < rstep; ++g) {
      const float4 t = red[0][g * C4 + threadIdx.x], u = red[1][g * C4 + threadIdx.x];
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      b.x += u.x; b.y += u.y; b.z += u.z; b.w += u.w;
    }
    double* st = sums + ((size_t)n * C4 + threadIdx.x) * 8;
    atomicAdd(st + 0, (double)a.x); atomicAdd(st + 1, (double)b.x);
    atomicAdd(st + 2, (double)a.y); atomicAdd(st + 3, (double)b.y);
    atomicAdd(st + 4, (double)a.z); atomicAdd(st + 5, (double)b.z);
    atomicAdd(st + 6, (double)a.w); atomicAdd(st + 7, (double)b.w);
  }
}

// ---- out = gamma*(gin*(x-mu)*rsqrt(var+eps)+bin) + alpha*mu_tilde + beta  ==  a*x + b per (n,c).  grid N.
__global__ void __launch_bounds__(512) k_in_coef(const double* __restrict__ sums, const float* __restrict__ g_gamma,
                                                 const float* __restrict__ g_alpha, const float* __restrict__ g_beta,
                                                 int gab_stride_n, const int* __restrict__ idx,
                                                 const float* __restrict__ in_gamma, const float* __restrict__ in_beta,
                                                 float2* __restrict__ coef, int HW, int C) {
  __shared__ double sh[2][16];
  __shared__ double stat[2];
  const int n = blockIdx.x, c = threadIdx.x;
  double mu = 0.0, var = 0.0;
  if (c < C) {
    mu = sums[((size_t)n * C + c) * 2] / HW;
    var = sums[((size_t)n * C + c) * 2 + 1] / HW - mu * mu;
    if (var < 0.0) var = 0.0;
  }
  // cross-channel mean / population variance of the per-channel means
  double a = c < C ? mu : 0.0, b = c < C ? mu * mu : 0.0;
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) { ta += sh[0][i]; tb += sh[1][i]; }
    const double m = ta / C;
    double v = tb / C - m * m;
    if (v < 0.0) v = 0.0;
    stat[0] = m;
    stat[1] = v;
  }
  __syncthreads();
  if (c >= C) return;
  const double mt = (mu - stat[0]) / sqrt(stat[1] + kPlusEps);
  const size_t row = (size_t)(idx ? idx[n] : 0) * gab_stride_n;            // Embedding row of this sample (v1)
  const float gamma = g_gamma[row + c], alpha = g_alpha[row + c], beta = g_beta[row + c];
  const float rs = rsqrtf((float)var + kInEps);
  const float gi = in_gamma[c], bi = in_beta[c];
  const float aa = gamma * gi * rs;
  const float bb = gamma * (bi - gi * (float)mu * rs) + alpha * (float)mt + beta;
  coef[(size_t)n * C + c] = make_float2(aa, bb);
}

// ---- y_bf16 = act(a*x+b).  grid (row slabs, N); a thread owns 8 channels - its 16 coefficients stay in registers -
// and every (blockDim / (C/8))-th pixel of the slab, two pixels in flight.
// returns the bf16 "hi" words; lo (if requested) receives bf16(v - hi), the second term of the split-bf16 operand
__device__ __forceinline__ uint4 prep8(const float4 v0, const float4 v1, const float (&ca)[8], const float (&cb)[8], int do_elu,
                                       uint4* lo) {
  float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = fmaf(ca[k], v[k], cb[k]);
  if (do_elu) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = elu(v[k]);
  }
  uint32_t w[4], wl[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
    w[k] = *reinterpret_cast<const uint32_t*>(&t);
    const float2 f = __bfloat1622float2(t);
    const __nv_bfloat162 u = __floats2bfloat162_rn(v[2 * k] - f.x, v[2 * k + 1] - f.y);
    wl[k] = *reinterpret_cast<const uint32_t*>(&u);
  }
  if (lo) *lo = make_uint4(wl[0], wl[1], wl[2], wl[3]);
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(256) k_prep(const float* __restrict__ x, const float2* __restrict__ coef,
                                              __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ y_lo, int HW, int C8,
                                              int rows_per_block, int do_elu) {
  const int n = blockIdx.y;
  const int rstep = blockDim.x / C8;
  const int c = threadIdx.x % C8, r0 = threadIdx.x / C8;
  if (r0 >= rstep) return;
  float ca[8], cb[8];
  if (coef) {
    const float4* cf = reinterpret_cast<const float4*>(coef + ((size_t)n * C8 + c) * 8);      // 8 x (a, b)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 t = __ldg(cf + k);
      ca[2 * k] = t.x; cb[2 * k] = t.y; ca[2 * k + 1] = t.z; cb[2 * k + 1] = t.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) { ca[k] = 1.f; cb[k] = 0.f; }
  }
  const int rbeg = blockIdx.x * rows_per_block, rend = min(HW, rbeg + rows_per_block);
  const float4* xb = reinterpret_cast<const float4*>(x) + ((size_t)n * HW) * (2 * C8) + 2 * c;
  uint4* yb = reinterpret_cast<uint4*>(y) + ((size_t)n * HW) * C8 + c;
  uint4* yl = y_lo ? reinterpret_cast<uint4*>(y_lo) + ((size_t)n * HW) * C8 + c : nullptr;
  int r = rbeg + r0;
  for (; r + rstep < rend; r += 2 * rstep) {
    const float4 a0 = __ldg(xb + (size_t)r * (2 * C8)), a1 = __ldg(xb + (size_t)r * (2 * C8) + 1);
    const float4 b0 = __ldg(xb + (size_t)(r + rstep) * (2 * C8)), b1 = __ldg(xb + (size_t)(r + rstep) * (2 * C8) + 1);
    yb[(size_t)r * C8] = prep8(a0, a1, ca, cb, do_elu, yl ? yl + (size_t)r * C8 : nullptr);
    yb[(size_t)(r + rstep) * C8] = prep8(b0, b1, ca, cb, do_elu, yl ? yl + (size_t)(r + rstep) * C8 : nullptr);
  }
  for (; r < rend; r += rstep) {
    const float4 a0 = __ldg(xb + (size_t)r * (2 * C8)), a1 = __ldg(xb + (size_t)r * (2 * C8) + 1);
    yb[(size_t)r * C8] = prep8(a0, a1, ca, cb, do_elu, yl ? yl + (size_t)r * C8 : nullptr);
  }
}

// ---- 5x5 stride-1 'same' pooling: average over the in-bounds taps (Keras AveragePooling2D) / max ignoring out-of-bounds
// taps (MaxPooling2D), in ONE pass: a thread owns one (column, 4 channels) strip of `seg_rows` image rows and walks down
// it with a ring of five horizontal 5-tap reductions, so the tensor is read once (+ 4 warm-up rows per strip) and written
// once (a separable two-kernel form moved it four times: 69 -> 45 us per pooling at 30 segments).
template <bool kMax>
__global__ void __launch_bounds__(256) k_pool5_roll(const float* __restrict__ x, float* __restrict__ y, int H, int W, int C4,
                                                    int seg_rows, int nseg, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C4);
  const int w = (int)((i / C4) % W);
  const int seg = (int)((i / ((long long)C4 * W)) % nseg);
  const long long n = i / ((long long)C4 * W * nseg);
  const int h0 = seg * seg_rows, h1 = min(H, h0 + seg_rows);
  const long long rs = (long long)W * C4;
  const float4* xb = reinterpret_cast<const float4*>(x) + (n * H * W + w) * C4 + c;
  float4* yb = reinterpret_cast<float4*>(y) + (n * H * W + w) * C4 + c;
  const float ident = kMax ? -INFINITY : 0.f;
  auto op = [](const float4 a, const float4 b) {
    return kMax ? make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w))
                : make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  };
  auto hrow = [&](int hh) {
    float4 acc = make_float4(ident, ident, ident, ident);
    if (hh < 0 || hh >= H) return acc;
#pragma unroll
    for (int d = -2; d <= 2; ++d) {
      if (w + d < 0 || w + d >= W) continue;
      acc = op(acc, __ldg(xb + hh * rs + d * C4));
    }
    return acc;
  };
  const int cw = min(w + 2, W - 1) - max(w - 2, 0) + 1;
  float4 r0 = hrow(h0 - 2), r1 = hrow(h0 - 1), r2 = hrow(h0), r3 = hrow(h0 + 1);
  for (int r = h0; r < h1; ++r) {
    const float4 r4 = hrow(r + 2);
    float4 o = op(op(op(r0, r1), op(r2, r3)), r4);
    if (!kMax) {
      const int ch = min(r + 2, H - 1) - max(r - 2, 0) + 1;
      const float q = 1.f / (float)(ch * cw);
      o.x *= q; o.y *= q; o.z *= q; o.w *= q;
    }
    yb[r * rs] = o;
    r0 = r1; r1 = r2; r2 = r3; r3 = r4;
  }
}

// ---- 2x2 stride-2 average pooling (H, W = output size)
__global__ void __launch_bounds__(256) k_avgpool2(const float* __restrict__ x, float* __restrict__ y, int H, int W, int C4,
                                                  long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C4);
  long long p = i / C4;
  const int w = (int)(p % W), h = (int)((p / W) % H);
  const long long img = p / ((long long)W * H);
  const float4* base = reinterpret_cast<const float4*>(x) + (img * 2 * H * 2 * W) * C4 + c;
  const long long r0 = ((long long)(2 * h) * 2 * W + 2 * w) * C4, r1 = r0 + (long long)2 * W * C4;
  const float4 a = __ldg(base + r0), b = __ldg(base + r0 + C4), d = __ldg(base + r1), e = __ldg(base + r1 + C4);
  reinterpret_cast<float4*>(y)[i] = make_float4(0.25f * (a.x + b.x + d.x + e.x), 0.25f * (a.y + b.y + d.y + e.y),
                                                0.25f * (a.z + b.z + d.z + e.z), 0.25f * (a.w + b.w + d.w + e.w));
}

// ---- y[N,2h,2w,C] = add + bilinear_x2(x[N,h,w,C]) (half-pixel centres, edge clamp; tf.image.resize bilinear)
__global__ void __launch_bounds__(256) k_resize2x_add(const float* __restrict__ x, const float* __restrict__ add,
                                                      float* __restrict__ y, int h, int w, int C4, long long total) {
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
  const float4* base = reinterpret_cast<const float4*>(x) + img * h * w * C4 + c;
  const float4 a = __ldg(base + ((long long)y0 * w + x0) * C4), b = __ldg(base + ((long long)y0 * w + x1) * C4);
  const float4 d = __ldg(base + ((long long)y1 * w + x0) * C4), e = __ldg(base + ((long long)y1 * w + x1) * C4);
  float4 o;
  o.x = (1.f - fy) * ((1.f - fx) * a.x + fx * b.x) + fy * ((1.f - fx) * d.x + fx * e.x);
  o.y = (1.f - fy) * ((1.f - fx) * a.y + fx * b.y) + fy * ((1.f - fx) * d.y + fx * e.y);
  o.z = (1.f - fy) * ((1.f - fx) * a.z + fx * b.z) + fy * ((1.f - fx) * d.z + fx * e.z);
  o.w = (1.f - fy) * ((1.f - fx) * a.w + fx * b.w) + fy * ((1.f - fx) * d.w + fx * e.w);
  if (add) { const float4 t = reinterpret_cast<const float4*>(add)[i]; o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
  reinterpret_cast<float4*>(y)[i] = o;
}

// ---- mode 0: y = elu(x); mode 1: y = x + z
__global__ void __launch_bounds__(256) k_ew(const float* __restrict__ x, const float* __restrict__ z, float* __restrict__ y,
                                            long long nvec, int mode) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  float4 v = reinterpret_cast<const float4*>(x)[i];
  if (mode == 0) { v.x = elu(v.x); v.y = elu(v.y); v.z = elu(v.z); v.w = elu(v.w); }
  else { const float4 t = reinterpret_cast<const float4*>(z)[i]; v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
  reinterpret_cast<float4*>(y)[i] = v;
}

// ---- begin_conv: 3x3 'same', 1 -> Cout channels, bias; v1 rescales the input 2x-1 first (score_network.py:277-280)
__global__ void __launch_bounds__(256) k_begin_conv(const float* __restrict__ x, const float* __restrict__ k,
                                                    const float* __restrict__ bias, float* __restrict__ y, int H, int W,
                                                    int C4, int rescale, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C4);
  long long p = i / C4;
  const int w = (int)(p % W), h = (int)((p / W) % H);
  float4 acc = __ldg(reinterpret_cast<const float4*>(bias) + c);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
    if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
    float v = __ldg(x + p + (long long)(tap / 3 - 1) * W + (tap % 3 - 1));
    if (rescale) v = 2.f * v - 1.f;
    const float4 kk = __ldg(reinterpret_cast<const float4*>(k) + tap * C4 + c);
    acc.x = fmaf(v, kk.x, acc.x); acc.y = fmaf(v, kk.y, acc.y); acc.z = fmaf(v, kk.z, acc.z); acc.w = fmaf(v, kk.w, acc.w);
  }
  reinterpret_cast<float4*>(y)[i] = acc;
}

// ---- end_conv: 3x3 'same', C -> 1 channel, bias, optional division by sigma[idx[n]].
// One warp per pixel; a lane owns 8 channels.  The nine 16-byte tap loads of a lane are issued back to back (out-of-image
// taps load nothing and contribute zero), the 9 x C kernel sits in shared memory.
__device__ __forceinline__ float dot8_bf16(const uint4 u, const float4 ka, const float4 kb, float acc) {
  acc = fmaf(__uint_as_float(u.x << 16), ka.x, acc); acc = fmaf(__uint_as_float(u.x & 0xffff0000u), ka.y, acc);
  acc = fmaf(__uint_as_float(u.y << 16), ka.z, acc); acc = fmaf(__uint_as_float(u.y & 0xffff0000u), ka.w, acc);
  acc = fmaf(__uint_as_float(u.z << 16), kb.x, acc); acc = fmaf(__uint_as_float(u.z & 0xffff0000u), kb.y, acc);
  acc = fmaf(__uint_as_float(u.w << 16), kb.z, acc); acc = fmaf(__uint_as_float(u.w & 0xffff0000u), kb.w, acc);
  return acc;
}

__global__ void __launch_bounds__(256) k_end_conv(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ x_lo,
                                                  const float* __restrict__ k,
                                                  const float* __restrict__ bias, const float* __restrict__ sigmas, const int* __restrict__ idx,
                                                  float* __restrict__ y, int H, int W, int C, long long pixels) {
  extern __shared__ __align__(16) float skw[];      // [9 * C]
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) skw[i] = __ldg(k + i);
  __syncthreads();
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= pixels) return;
  const int c0 = lane * 8;
  const bool active = c0 < C;                       // C <= 256 (checked by the launcher), C % 8 == 0
  const int w = (int)(p % W), h = (int)((p / W) % H);
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  uint4 u[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
    const bool ok = active && hh >= 0 && hh < H && ww >= 0 && ww < W;
    const long long off = (p + (long long)(tap / 3 - 1) * W + (tap % 3 - 1)) * C + c0;
    u[tap] = ok ? __ldg(reinterpret_cast<const uint4*>(x + off)) : zero;
  }
  float acc = 0.f;
  const int cw = active ? c0 : 0;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const float4 ka = *reinterpret_cast<const float4*>(skw + tap * C + cw), kb = *reinterpret_cast<const float4*>(skw + tap * C + cw + 4);
    acc = dot8_bf16(u[tap], ka, kb, acc);
  }
  if (x_lo) {                                       // split-bf16 operand: second term
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
      const bool ok = active && hh >= 0 && hh < H && ww >= 0 && ww < W;
      const long long off = (p + (long long)(tap / 3 - 1) * W + (tap % 3 - 1)) * C + c0;
      u[tap] = ok ? __ldg(reinterpret_cast<const uint4*>(x_lo + off)) : zero;
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float4 ka = *reinterpret_cast<const float4*>(skw + tap * C + cw), kb = *reinterpret_cast<const float4*>(skw + tap * C + cw + 4);
      acc = dot8_bf16(u[tap], ka, kb, acc);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float o = acc + __ldg(bias);
    if (sigmas) o /= sigmas[idx[p / ((long long)H * W)]];
    y[p] = o;
  }
}

}  // namespace

void launch_in_stats(const float* x, double* sums, int N, int HW, int C, cudaStream_t s) {
  CUDA_CHECK(cudaMemsetAsync(sums, 0, (size_t)N * C * 2 * sizeof(double), s));
  ASEP_CHECK(C % 4 == 0 && C / 4 <= 256, ASEP_ERR_UNSUPPORTED, "instance-norm statistics: C = %d (multiple of 4, <= 1024)", C);
  const int C4 = C / 4;
  const int threads = C4 * (256 / C4);
  const int rows = 32;
  dim3 grid((HW + rows - 1) / rows, N);
  k_in_stats<<<grid, threads, 0, s>>>(x, sums, HW, C4, rows);
  ASEP_LAUNCH_CHECK();
}

void launch_in_coef(const double* sums, const float* gamma, const float* alpha, const float* beta, int gab_stride_n,
                    const int* idx, const float* in_gamma,
                    const float* in_beta, float2* coef, int N, int HW, int C, cudaStream_t s) {
  const int threads = (C + 31) / 32 * 32;
  ASEP_CHECK(threads <= 512, ASEP_ERR_UNSUPPORTED, "instance-norm coefficients: C = %d > 512", C);
  k_in_coef<<<N, threads, 0, s>>>(sums, gamma, alpha, beta, gab_stride_n, idx, in_gamma, in_beta, coef, HW, C);
  ASEP_LAUNCH_CHECK();
}

void launch_prep(const float* x, const float2* coef, __nv_bfloat16* y, __nv_bfloat16* y_lo, int N, int HW, int C, int do_elu,
                 cudaStream_t s) {
  ASEP_CHECK(C % 8 == 0 && C / 8 <= 256, ASEP_ERR_UNSUPPORTED, "prep: C = %d (multiple of 8, <= 2048)", C);
  // reads the fp32 tensor once, writes the bf16 operand (and its low word in the split mode)
  HbmScope prof(kHbmPrep, (4.0 + (y_lo ? 4.0 : 2.0)) * (double)N * HW * C, s);
  const int C8 = C / 8;
  const int threads = C8 * (256 / C8);
  const int rows = 32;
  dim3 grid((HW + rows - 1) / rows, N);
  k_prep<<<grid, threads, 0, s>>>(x, coef, y, y_lo, HW, C8, rows, do_elu);
  ASEP_LAUNCH_CHECK();
}

void launch_pool5(const float* x, float* tmp, float* y, int N, int H, int W, int C, int is_max, cudaStream_t s) {
  (void)tmp;
  HbmScope prof(kHbmPool, 8.0 * (double)N * H * W * C, s);
  const int seg_rows = 24, nseg = (H + seg_rows - 1) / seg_rows;
  const long long total = (long long)N * nseg * W * (C / 4);
  if (is_max) k_pool5_roll<true><<<cdiv(total, 256), 256, 0, s>>>(x, y, H, W, C / 4, seg_rows, nseg, total);
  else k_pool5_roll<false><<<cdiv(total, 256), 256, 0, s>>>(x, y, H, W, C / 4, seg_rows, nseg, total);
  ASEP_LAUNCH_CHECK();
}

void launch_avgpool2(const float* x, float* y, int N, int Hout, int Wout, int C, cudaStream_t s) {
  const long long total = (long long)N * Hout * Wout * (C / 4);
  HbmScope prof(kHbmPool, 20.0 * (double)N * Hout * Wout * C, s);   // reads 4 inputs per output, writes 1
  k_avgpool2<<<cdiv(total, 256), 256, 0, s>>>(x, y, Hout, Wout, C / 4, total);
  ASEP_LAUNCH_CHECK();
}

void launch_resize2x_add(const float* x, const float* add, float* y, int N, int h, int w, int C, cudaStream_t s) {
  const long long total = (long long)N * 4 * h * w * (C / 4);
  HbmScope prof(kHbmPool, (4.0 + (add ? 16.0 : 0.0) + 16.0) * (double)N * h * w * C, s);   // x once, add + y at 4x the pixels
  k_resize2x_add<<<cdiv(total, 256), 256, 0, s>>>(x, add, y, h, w, C / 4, total);
  ASEP_LAUNCH_CHECK();
}

void launch_elu_stats(const float* x, float* y, double* sums, int N, int HW, int C, cudaStream_t s) {
  CUDA_CHECK(cudaMemsetAsync(sums, 0, (size_t)N * C * 2 * sizeof(double), s));
  ASEP_CHECK(C % 4 == 0 && C / 4 <= 256, ASEP_ERR_UNSUPPORTED, "elu + statistics: C = %d (multiple of 4, <= 1024)", C);
  const int C4 = C / 4;
  const int threads = C4 * (256 / C4);
  const int rows = 64;
  dim3 grid((HW + rows - 1) / rows, N);
  k_elu_stats<<<grid, threads, 0, s>>>(x, y, sums, HW, C4, rows);
  ASEP_LAUNCH_CHECK();
}

void launch_elu(const float* x, float* y, long long n, cudaStream_t s) {
  k_ew<<<cdiv(n / 4, 256), 256, 0, s>>>(x, nullptr, y, n / 4, 0);
  ASEP_LAUNCH_CHECK();
}

void launch_add(const float* x, const float* z, float* y, long long n, cudaStream_t s) {
  k_ew<<<cdiv(n / 4, 256), 256, 0, s>>>(x, z, y, n / 4, 1);
  ASEP_LAUNCH_CHECK();
}

void launch_begin_conv(const float* x, const float* k, const float* bias, float* y, int N, int H, int W, int Cout,
                       int rescale, cudaStream_t s) {
  const long long total = (long long)N * H * W * (Cout / 4);
  k_begin_conv<<<cdiv(total, 256), 256, 0, s>>>(x, k, bias, y, H, W, Cout / 4, rescale, total);
  ASEP_LAUNCH_CHECK();
}

void launch_end_conv(const __nv_bfloat16* x, const __nv_bfloat16* x_lo, const float* k, const float* bias, const float* sigmas, const int* idx, float* y,
                     int N, int H, int W, int C, cudaStream_t s) {
  const long long pixels = (long long)N * H * W;
  ASEP_CHECK(C % 8 == 0 && C <= 256, ASEP_ERR_UNSUPPORTED, "end_conv: C = %d (multiple of 8, <= 256)", C);
  k_end_conv<<<cdiv(pixels * 32, 256), 256, 9 * C * sizeof(float), s>>>(x, x_lo, k, bias, sigmas, idx, y, H, W, C, pixels);
  ASEP_LAUNCH_CHECK();
}

}  // namespace asep
