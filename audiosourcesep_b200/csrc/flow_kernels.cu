// Element-wise kernels of the Glow bijector chain: ActNorm, invertible 1x1 convolution, affine
// coupling with fused per-sample log-det reduction, squeeze / factor-out plumbing, prior.
// HBM-bound (state rows of C floats per pixel, one pixel per thread, 16-byte vector access);
// log-dets are warp-shuffle reduced and accumulated in double so the result does not depend
// on the order atomics land in.
//
// Reference math: flow_models/flow_tfp_bijectors.py:124-153 (coupling), :156-199 (squeeze),
// :202-253 (ActNorm), :256-322 (1x1), :364-396 (SpecPreprocessing); flow_models/flow_glow.py:176-196
// (factor-out reshapes); flow_models/flow_builder.py:132-139 (prior).
#include "kernels.h"

namespace asep {

namespace {

constexpr int kThreads = 128;

template <int C>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&v)[C]) {
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int i = 0; i < C / 4; ++i) {
      float4 t = reinterpret_cast<const float4*>(p)[i];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else if constexpr (C % 2 == 0) {
#pragma unroll
    for (int i = 0; i < C / 2; ++i) {
      float2 t = reinterpret_cast<const float2*>(p)[i];
      v[2 * i] = t.x; v[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < C; ++i) v[i] = p[i];
  }
}

template <int C>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&v)[C]) {
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int i = 0; i < C / 4; ++i)
      reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if constexpr (C % 2 == 0) {
#pragma unroll
    for (int i = 0; i < C / 2; ++i) reinterpret_cast<float2*>(p)[i] = make_float2(v[2 * i], v[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < C; ++i) p[i] = v[i];
  }
}

template <int C>
__device__ __forceinline__ void stage_consts(const float* __restrict__ sc, float* smem) {
  constexpr int n = 2 * C + 2 * C * C;
  for (int i = threadIdx.x; i < n; i += blockDim.x) smem[i] = sc[i];
  __syncthreads();
}

// a = x*scale + shift ; u = a . W
template <int C>
__device__ __forceinline__ void pre_apply(const float (&x)[C], float (&u)[C], const float* smem) {
  float a[C];
#pragma unroll
  for (int i = 0; i < C; ++i) a[i] = x[i] * smem[i] + smem[C + i];
  const float* Wm = smem + 2 * C;
#pragma unroll
  for (int o = 0; o < C; ++o) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) acc = fmaf(a[i], Wm[i * C + o], acc);
    u[o] = acc;
  }
}

// Adds a per-pixel value to the per-sample double accumulator with one atomic per warp
// when the whole warp sits inside one sample.
__device__ __forceinline__ void accumulate_sample(double* acc, float v, long long pix, long long M, int HW) {
  const bool valid = pix < M;
  const int n = valid ? (int)(pix / HW) : -1;
  const int n0 = __shfl_sync(0xffffffffu, n, 0);
  const bool uniform = __all_sync(0xffffffffu, n == n0 || !valid);
  if (uniform) {
    float s = warp_sum(valid ? v : 0.f);
    if ((threadIdx.x & 31) == 0 && n0 >= 0) atomicAdd(acc + n0, (double)s);
  } else if (valid) {
    atomicAdd(acc + n, (double)v);
  }
}

// Device view of GatherSrc (kernel argument).
struct GArg {
  const float* G; const float* const3; const float* c3;
  int n3p, nparts, H, W;
  long long part_stride;
};

// r[c] = c3[c] + sum_{tap in bounds} (G[p + off(tap)][tap*C + c] + const3[tap*C + c])     (same order and association
// as k_gather_vec / k_gather_fwd in nn_tc_shared.cuh, so fused and unfused results are bit-identical)
template <int C>
__device__ __forceinline__ void gather_fwd_row(const GArg& g, long long p, float (&rv)[C]) {
  const int w = (int)(p % g.W), h = (int)((p / g.W) % g.H);
#pragma unroll
  for (int c = 0; c < C; ++c) rv[c] = __ldg(g.c3 + c);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int hh = h + dy, ww = w + dx;
    if (hh < 0 || hh >= g.H || ww < 0 || ww >= g.W) continue;
    const long long pp = p + (long long)dy * g.W + dx;
    const float* base = g.G + (pp >> 7) * (128ll * g.n3p) + (pp & 127) * 4;
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
      const int col = tap * C + 4 * q;
      const float* src = base + (long long)(col >> 2) * 512;
      float4 t = __ldg(reinterpret_cast<const float4*>(src));
      for (int part = 1; part < g.nparts; ++part) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + part * g.part_stride));
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      const float4 k = __ldg(reinterpret_cast<const float4*>(g.const3 + col));
      t.x += k.x; t.y += k.y; t.z += k.z; t.w += k.w;
      rv[4 * q] += t.x; rv[4 * q + 1] += t.y; rv[4 * q + 2] += t.z; rv[4 * q + 3] += t.w;
    }
  }
}

// gxb[ci] = sum_{tap: p - off in bounds} G'[p - off(tap)][tap*Ch + ci]      (Ch = 2: 8-byte loads, else 16-byte)
template <int Ch>
__device__ __forceinline__ void gather_bwd_row(const GArg& g, long long p, float (&gb)[Ch]) {
  const int w = (int)(p % g.W), h = (int)((p / g.W) % g.H);
#pragma unroll
  for (int c = 0; c < Ch; ++c) gb[c] = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = -(tap / 3 - 1), dx = -(tap % 3 - 1);
    const int hh = h + dy, ww = w + dx;
    if (hh < 0 || hh >= g.H || ww < 0 || ww >= g.W) continue;
    const long long pp = p + (long long)dy * g.W + dx;
    const float* base = g.G + (pp >> 7) * (128ll * g.n3p) + (pp & 127) * 4;
    if constexpr (Ch % 4 == 0) {
#pragma unroll
      for (int q = 0; q < Ch / 4; ++q) {
        const int col = tap * Ch + 4 * q;
        const float* src = base + (long long)(col >> 2) * 512;
        float4 t = __ldg(reinterpret_cast<const float4*>(src));
        for (int part = 1; part < g.nparts; ++part) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src + part * g.part_stride));
          t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        gb[4 * q] += t.x; gb[4 * q + 1] += t.y; gb[4 * q + 2] += t.z; gb[4 * q + 3] += t.w;
      }
    } else {
      static_assert(Ch == 2, "gather_bwd_row: channel count");
      const int col = tap * 2;
      const float* src = base + (long long)(col >> 2) * 512 + (col & 3);
      float2 t = __ldg(reinterpret_cast<const float2*>(src));
      for (int part = 1; part < g.nparts; ++part) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(src + part * g.part_stride));
        t.x += v.x; t.y += v.y;
      }
      gb[0] += t.x; gb[1] += t.y;
    }
  }
}

template <int C>
__global__ void __launch_bounds__(kThreads) k_pre(const float* __restrict__ x, float* __restrict__ u,
                                                  const float* __restrict__ sc, long long M) {
  __shared__ float smem[2 * C + 2 * C * C];
  stage_consts<C>(sc, smem);
  long long p = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (p >= M) return;
  float xv[C], uv[C];
  load_row<C>(x + p * C, xv);
  pre_apply<C>(xv, uv, smem);
  store_row<C>(u + p * C, uv);
}

template <int C, bool kNext, bool kGather = false>
__global__ void __launch_bounds__(kThreads) k_post_pre(const float* __restrict__ u, const float* __restrict__ r,
                                                       float* __restrict__ out, const float* __restrict__ sc_next,
                                                       double* __restrict__ acc, long long M, int HW, const GArg g = GArg{},
                                                       float* __restrict__ r_out = nullptr) {
  __shared__ float smem[kNext ? 2 * C + 2 * C * C : 1];
  if constexpr (kNext) stage_consts<C>(sc_next, smem);
  long long p = (long long)blockIdx.x * kThreads + threadIdx.x;
  float sum_sl = 0.f;
  if (p < M) {
    float uv[C], rv[C], y[C];
    load_row<C>(u + p * C, uv);
    if constexpr (kGather) {
      gather_fwd_row<C>(g, p, rv);
      if (r_out != nullptr) store_row<C>(r_out + p * C, rv);
    } else {
      load_row<C>(r + p * C, rv);
    }
#pragma unroll
    for (int c = 0; c < C / 2; ++c) {
      float sl = tanhf(rv[c]);                       // flow_tfk_layers.py:83
      sum_sl += sl;
      y[c] = expf(sl) * uv[c] + rv[C / 2 + c];       // flow_tfp_bijectors.py:137-138
      y[C / 2 + c] = uv[C / 2 + c];
    }
    if constexpr (kNext) {
      float o[C];
      pre_apply<C>(y, o, smem);
      store_row<C>(out + p * C, o);
    } else {
      store_row<C>(out + p * C, y);
    }
  }
  if (acc != nullptr) accumulate_sample(acc, sum_sl, p, M, HW);
}

template <int C, bool kGather = false>
__global__ void __launch_bounds__(kThreads) k_inv_step(const float* __restrict__ y, const float* __restrict__ r,
                                                       float* __restrict__ x, const float* __restrict__ sc,
                                                       double* __restrict__ acc, long long M, int HW, const GArg g = GArg{}) {
  __shared__ float smem[2 * C + 2 * C * C];
  stage_consts<C>(sc, smem);
  long long p = (long long)blockIdx.x * kThreads + threadIdx.x;
  float sum_sl = 0.f;
  if (p < M) {
    float yv[C], rv[C], uv[C], xv[C];
    load_row<C>(y + p * C, yv);
    if constexpr (kGather) gather_fwd_row<C>(g, p, rv);
    else load_row<C>(r + p * C, rv);
#pragma unroll
    for (int c = 0; c < C / 2; ++c) {
      float sl = tanhf(rv[c]);
      sum_sl += sl;
      uv[c] = (yv[c] - rv[C / 2 + c]) / expf(sl);    // flow_tfp_bijectors.py:146
      uv[C / 2 + c] = yv[C / 2 + c];
    }
    const float* Wi = smem + 2 * C + C * C;
#pragma unroll
    for (int o = 0; o < C; ++o) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < C; ++i) a = fmaf(uv[i], Wi[i * C + o], a);
      xv[o] = (a - smem[C + o]) / smem[o];           // flow_tfp_bijectors.py:246-247
    }
    store_row<C>(x + p * C, xv);
  }
  if (acc != nullptr) accumulate_sample(acc, -sum_sl, p, M, HW);
}

template <int C>
__global__ void __launch_bounds__(kThreads) k_bwd_coupling(const float* __restrict__ gy, const float* __restrict__ u,
                                                           const float* __restrict__ r, float* __restrict__ gr,
                                                           float* __restrict__ gu, long long M) {
  long long p = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (p >= M) return;
  float g[C], uv[C], rv[C], grv[C], guv[C];
  load_row<C>(gy + p * C, g);
  load_row<C>(u + p * C, uv);
  load_row<C>(r + p * C, rv);
#pragma unroll
  for (int c = 0; c < C / 2; ++c) {
    float sl = tanhf(rv[c]);
    float e = expf(sl);
    guv[c] = g[c] * e;
    guv[C / 2 + c] = g[C / 2 + c];
    float gsl = g[c] * uv[c] * e + 1.0f;             // + d(sum sl)/d sl
    grv[c] = gsl * (1.0f - sl * sl);
    grv[C / 2 + c] = g[c];
  }
  store_row<C>(gr + p * C, grv);
  store_row<C>(gu + p * C, guv);
}

template <int C, bool kGather = false>
__global__ void __launch_bounds__(kThreads) k_bwd_pre(const float* __restrict__ gu, const float* __restrict__ gxb,
                                                      float* __restrict__ gx, const float* __restrict__ sc,
                                                      long long M, const GArg ga = GArg{}) {
  __shared__ float smem[2 * C + 2 * C * C];
  stage_consts<C>(sc, smem);
  long long p = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (p >= M) return;
  float g[C], gb[C / 2], o[C];
  load_row<C>(gu + p * C, g);
  if constexpr (kGather) gather_bwd_row<C / 2>(ga, p, gb);
  else load_row<C / 2>(gxb + p * (C / 2), gb);
#pragma unroll
  for (int c = 0; c < C / 2; ++c) g[C / 2 + c] += gb[c];
  const float* Wm = smem + 2 * C;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) a = fmaf(g[k], Wm[i * C + k], a);
    o[i] = a * smem[i];
  }
  store_row<C>(gx + p * C, o);
}

__device__ __forceinline__ float squeeze_affine(float v, int mode, float p0, float p1) {
  switch (mode) {
    case 1: return (v - p0) / (p1 - p0) - 0.5f;      // SpecPreprocessing._forward (:372-379)
    case 2: return (v + 0.5f) * (p1 - p0) + p0;      // SpecPreprocessing._inverse (:381-388)
    case 3: return v * p0;
    default: return v;
  }
}

__global__ void k_squeeze(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int mode,
                          float p0, float p1, int inverse) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * H * W * C;   // H, W, C describe the UNSQUEEZED tensor
  if (idx >= total) return;
  int c = idx % C;
  long long t = idx / C;
  int w = t % W; t /= W;
  int h = t % H;
  int n = t / H;
  long long sq = (((long long)n * (H / 2) + h / 2) * (W / 2) + w / 2) * (4 * C) + c * 4 + (h & 1) * 2 + (w & 1);
  if (!inverse) y[sq] = squeeze_affine(x[idx], mode, p0, p1);
  else y[idx] = squeeze_affine(x[sq], mode, p0, p1);
}

// o [N,H,W,C]: first Cz channels -> latent (row-major reshape to [Hl*Wl, nb] at channel offset coff of CL),
// remaining C-Cz channels -> squeezed next state [N,H/2,W/2,4*(C-Cz)].
__global__ void k_split_merge(float* __restrict__ o, float* __restrict__ z, float* __restrict__ next, int N, int H,
                              int W, int C, int Cz, int nb, int CL, int coff, long long Dl, int merge) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * H * W * C;
  if (idx >= total) return;
  int c = idx % C;
  long long t = idx / C;
  int w = t % W; t /= W;
  int h = t % H;
  int n = t / H;
  long long other;
  float* buf;
  if (c < Cz) {
    long long f = ((long long)h * W + w) * Cz + c;
    other = (long long)n * Dl + (f / nb) * CL + coff + (f % nb);
    buf = z;
  } else {
    int Cn = C - Cz;
    other = (((long long)n * (H / 2) + h / 2) * (W / 2) + w / 2) * (4 * Cn) + (c - Cz) * 4 + (h & 1) * 2 + (w & 1);
    buf = next;
  }
  if (!merge) buf[other] = o[idx];
  else o[idx] = buf[other];
}

__global__ void k_prior(const float* __restrict__ z, const float* __restrict__ loc, const float* __restrict__ ls,
                        double* __restrict__ acc, float* __restrict__ gz, int D) {
  const int n = blockIdx.x;
  float s = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float l = loc ? loc[d] : 0.f, v = ls ? ls[d] : 0.f;
    float inv = expf(-v);
    float q = (z[(long long)n * D + d] - l) * inv;
    s += -0.5f * q * q - v - 0.91893853320467274178f;
    if (gz) gz[(long long)n * D + d] = -q * inv;
  }
  __shared__ float red[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0 && acc) atomicAdd(acc + n, (double)v);
  }
}

__global__ void k_prior_sample(const float* __restrict__ eps, const float* __restrict__ loc,
                               const float* __restrict__ ls, float* __restrict__ z, long long total, int D) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int d = i % D;
  z[i] = (loc ? loc[d] : 0.f) + expf(ls ? ls[d] : 0.f) * eps[i];
}

__global__ void k_finish(const double* __restrict__ acc, float* __restrict__ out, double add, const double* __restrict__ dev_add,
                         double scale, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = (float)((acc[i] + add + (dev_add ? *dev_add : 0.0)) * scale);
}

// mean and population std per channel over M pixels (ActNorm init, flow_tfp_bijectors.py:222-226)
__global__ void k_channel_stats(const float* __restrict__ x, double* __restrict__ out, long long M, int C) {
  const int c = blockIdx.x;
  double s = 0.0, s2 = 0.0;
  for (long long p = threadIdx.x; p < M; p += blockDim.x) {
    double v = x[p * C + c];
    s += v; s2 += v * v;
  }
  __shared__ double rs[256], rs2[256];
  rs[threadIdx.x] = s; rs2[threadIdx.x] = s2;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { rs[threadIdx.x] += rs[threadIdx.x + o]; rs2[threadIdx.x] += rs2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double mean = rs[0] / (double)M;
    double var = rs2[0] / (double)M - mean * mean;
    out[c] = mean;
    out[C + c] = sqrt(var > 0.0 ? var : 0.0);
  }
}

__global__ void k_scale(float* x, float a, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= a;
}

__global__ void k_actnorm(const float* __restrict__ x, const float* __restrict__ ls, const float* __restrict__ sh,
                          float* __restrict__ y, long long total, int C, int inverse) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = i % C;
  y[i] = inverse ? (x[i] - sh[c]) / expf(ls[c]) : x[i] * expf(ls[c]) + sh[c];
}

__global__ void k_chanmix(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                          long long M, int C) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  int o = i % C;
  long long p = i / C;
  float a = 0.f;
  for (int k = 0; k < C; ++k) a = fmaf(x[p * C + k], w[k * C + o], a);
  y[i] = a;
}

}  // namespace

#define DISPATCH_C(C, ...)                                                                   \
  switch (C) {                                                                               \
    case 2: { constexpr int kC = 2; __VA_ARGS__; } break;                                    \
    case 4: { constexpr int kC = 4; __VA_ARGS__; } break;                                    \
    case 8: { constexpr int kC = 8; __VA_ARGS__; } break;                                    \
    case 16: { constexpr int kC = 16; __VA_ARGS__; } break;                                  \
    case 32: { constexpr int kC = 32; __VA_ARGS__; } break;                                  \
    default: throw Error(ASEP_ERR_UNSUPPORTED, strfmt("channel count %d not built (2,4,8,16,32)", C)); \
  }

void launch_pre(const float* x, float* u, const float* sc, long long M, int C, cudaStream_t s) {
  if (M == 0) return;
  HbmScope prof(kHbmFlowStep, 8.0 * C * (double)M, s);            // read x, write u
  DISPATCH_C(C, (k_pre<kC><<<cdiv(M, kThreads), kThreads, 0, s>>>(x, u, sc, M)));
  ASEP_LAUNCH_CHECK();
}

void launch_post_pre(const float* u, const float* r, float* out, const float* sc_next, double* acc, long long M,
                     int HW, int C, cudaStream_t s) {
  if (M == 0) return;
  HbmScope prof(kHbmFlowStep, 12.0 * C * (double)M, s);           // read u, read r, write the next state
  if (sc_next) {
    DISPATCH_C(C, (k_post_pre<kC, true><<<cdiv(M, kThreads), kThreads, 0, s>>>(u, r, out, sc_next, acc, M, HW)));
  } else {
    DISPATCH_C(C, (k_post_pre<kC, false><<<cdiv(M, kThreads), kThreads, 0, s>>>(u, r, out, nullptr, acc, M, HW)));
  }
  ASEP_LAUNCH_CHECK();
}

void launch_inv_step(const float* y, const float* r, float* x, const float* sc, double* acc, long long M, int HW,
                     int C, cudaStream_t s) {
  if (M == 0) return;
  HbmScope prof(kHbmFlowStep, 12.0 * C * (double)M, s);
  DISPATCH_C(C, (k_inv_step<kC><<<cdiv(M, kThreads), kThreads, 0, s>>>(y, r, x, sc, acc, M, HW)));
  ASEP_LAUNCH_CHECK();
}

namespace {
GArg to_arg(const GatherSrc& g) { return GArg{g.G, g.const3, g.c3, g.n3p, g.nparts, g.H, g.W, g.part_stride}; }
// the fused gather reads whole float4 columns of G: C (forward) must be a multiple of 4, C/2 (backward) 2 or a multiple of 4
#define DISPATCH_CG(C, EXPR)                                                                           \
  do {                                                                                                 \
    switch (C) {                                                                                       \
      case 4: { constexpr int kC = 4; EXPR; } break;                                                   \
      case 8: { constexpr int kC = 8; EXPR; } break;                                                   \
      case 16: { constexpr int kC = 16; EXPR; } break;                                                 \
      default: ASEP_CHECK(false, ASEP_ERR_UNSUPPORTED, "fused gather: channel count %d", C);           \
    }                                                                                                  \
  } while (0)
}  // namespace

bool fused_gather_supported(int C) { return C == 4 || C == 8 || C == 16; }

void launch_post_pre_g(const float* u, const GatherSrc& g, float* r_out, float* out, const float* sc_next, double* acc,
                       long long M, int HW, int C, cudaStream_t s) {
  if (M == 0) return;
  // algorithmic bytes as for the unfused step (state in, network output, state out: 12 C per pixel; + r when it is kept);
  // the fused col2im really reads the 9 per-tap partials (36 C per pixel) instead of r
  HbmScope prof(kHbmFlowStep, (r_out ? 16.0 : 12.0) * C * (double)M, s);
  const GArg a = to_arg(g);
  if (sc_next) {
    DISPATCH_CG(C, (k_post_pre<kC, true, true><<<cdiv(M, kThreads), kThreads, 0, s>>>(u, nullptr, out, sc_next, acc, M, HW, a, r_out)));
  } else {
    DISPATCH_CG(C, (k_post_pre<kC, false, true><<<cdiv(M, kThreads), kThreads, 0, s>>>(u, nullptr, out, nullptr, acc, M, HW, a, r_out)));
  }
  ASEP_LAUNCH_CHECK();
}

void launch_inv_step_g(const float* y, const GatherSrc& g, float* x, const float* sc, double* acc, long long M, int HW,
                       int C, cudaStream_t s) {
  if (M == 0) return;
  HbmScope prof(kHbmFlowStep, 12.0 * C * (double)M, s);
  const GArg a = to_arg(g);
  DISPATCH_CG(C, (k_inv_step<kC, true><<<cdiv(M, kThreads), kThreads, 0, s>>>(y, nullptr, x, sc, acc, M, HW, a)));
  ASEP_LAUNCH_CHECK();
}

void launch_bwd_pre_g(const float* gu, const GatherSrc& g, float* gx, const float* sc, long long M, int C, cudaStream_t s) {
  if (M == 0) return;
  HbmScope prof(kHbmFlowStep, 10.0 * C * (double)M, s);           // read gu (C), gxb (C/2), write gx (C)
  const GArg a = to_arg(g);
  DISPATCH_CG(C, (k_bwd_pre<kC, true><<<cdiv(M, kThreads), kThreads, 0, s>>>(gu, nullptr, gx, sc, M, a)));
  ASEP_LAUNCH_CHECK();
}

void launch_bwd_coupling(const float* gy, const float* u, const float* r, float* gr, float* gu, long long M, int C,
                         cudaStream_t s) {
  if (M == 0) return;
  HbmScope prof(kHbmFlowStep, 20.0 * C * (double)M, s);           // read gy, u, r; write gr, gu
  DISPATCH_C(C, (k_bwd_coupling<kC><<<cdiv(M, kThreads), kThreads, 0, s>>>(gy, u, r, gr, gu, M)));
  ASEP_LAUNCH_CHECK();
}

void launch_bwd_pre(const float* gu, const float* gxb, float* gx, const float* sc, long long M, int C,
                    cudaStream_t s) {
  if (M == 0) return;
  HbmScope prof(kHbmFlowStep, 10.0 * C * (double)M, s);
  DISPATCH_C(C, (k_bwd_pre<kC><<<cdiv(M, kThreads), kThreads, 0, s>>>(gu, gxb, gx, sc, M)));
  ASEP_LAUNCH_CHECK();
}

void launch_squeeze(const float* x, float* y, int N, int H, int W, int C, int mode, float p0, float p1, int inverse,
                    cudaStream_t s) {
  long long total = (long long)N * H * W * C;
  if (total == 0) return;
  k_squeeze<<<cdiv(total, 256), 256, 0, s>>>(x, y, N, H, W, C, mode, p0, p1, inverse);
  ASEP_LAUNCH_CHECK();
}

void launch_split_merge(float* o, float* z, float* next, int N, int H, int W, int C, int Cz, int nb, int CL,
                        int coff, long long Dl, int merge, cudaStream_t s) {
  long long total = (long long)N * H * W * C;
  if (total == 0) return;
  k_split_merge<<<cdiv(total, 256), 256, 0, s>>>(o, z, next, N, H, W, C, Cz, nb, CL, coff, Dl, merge);
  ASEP_LAUNCH_CHECK();
}

void launch_prior(const float* z, const float* loc, const float* log_scale, double* acc, float* gz, int N, int D,
                  cudaStream_t s) {
  if (N == 0) return;
  k_prior<<<N, 256, 0, s>>>(z, loc, log_scale, acc, gz, D);
  ASEP_LAUNCH_CHECK();
}

void launch_prior_sample(const float* eps, const float* loc, const float* log_scale, float* z, int N, int D,
                         cudaStream_t s) {
  long long total = (long long)N * D;
  if (total == 0) return;
  k_prior_sample<<<cdiv(total, 256), 256, 0, s>>>(eps, loc, log_scale, z, total, D);
  ASEP_LAUNCH_CHECK();
}

void launch_finish(const double* acc, float* out, double add, double scale, int N, cudaStream_t s, const double* dev_add) {
  if (N == 0) return;
  k_finish<<<cdiv(N, 128), 128, 0, s>>>(acc, out, add, dev_add, scale, N);
  ASEP_LAUNCH_CHECK();
}

void launch_channel_stats(const float* x, double* mean_std, long long M, int C, cudaStream_t s) {
  k_channel_stats<<<C, 256, 0, s>>>(x, mean_std, M, C);
  ASEP_LAUNCH_CHECK();
}

void launch_scale(float* x, float a, long long n, cudaStream_t s) {
  if (n == 0) return;
  k_scale<<<cdiv(n, 256), 256, 0, s>>>(x, a, n);
  ASEP_LAUNCH_CHECK();
}

void launch_actnorm(const float* x, const float* log_scale, const float* shift, float* y, long long M, int C,
                    int inverse, cudaStream_t s) {
  if (M == 0) return;
  k_actnorm<<<cdiv(M * C, 256), 256, 0, s>>>(x, log_scale, shift, y, M * C, C, inverse);
  ASEP_LAUNCH_CHECK();
}

void launch_chanmix(const float* x, const float* w, float* y, long long M, int C, cudaStream_t s) {
  if (M == 0) return;
  k_chanmix<<<cdiv(M * C, 256), 256, 0, s>>>(x, w, y, M, C);
  ASEP_LAUNCH_CHECK();
}

}  // namespace asep
