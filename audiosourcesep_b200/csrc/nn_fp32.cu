// CUDA-core fp32 implementation of ShiftAndLogScaleConvNet (flow_tfk_layers.py:73-84) and of its
// data gradient.  This is the ASEP_PREC_FP32 "exact" mode: plain fp32 FMAs in the reference's
// operator order (conv+bias -> ReLU -> inference BatchNorm affine -> ...).  It exists so that the
// tcgen05 path can be checked on the device at full size and so that tight-tolerance parity
// (round trip <= 1e-4) has an arithmetic that is not rounded to bf16.  It is NOT tuned.
#include "kernels.h"

namespace asep {
namespace {

// a1[p][f] = relu(c1[f] + sum_{tap,ci} xb[p+off(tap)][ci] * K1[tap][ci][f])
// grid (pixel, F / 256): a block owns one pixel (its tap inputs are warp-uniform loads) and 256 output channels
// (coalesced weight reads and stores); pixel coordinates from 32-bit divisions (M < 2^31).
__global__ void __launch_bounds__(256) k_conv1(const float* __restrict__ state, const float* __restrict__ k1,
                                               const float* __restrict__ c1, float* __restrict__ a1, int H, int W, int C,
                                               int F) {
  const int Ch = C / 2;
  const unsigned p = blockIdx.x;
  const int f = blockIdx.y * 256 + threadIdx.x;
  if (f >= F) return;
  const int w = (int)(p % (unsigned)W), h = (int)((p / (unsigned)W) % (unsigned)H);
  float acc = c1[f];
  for (int dy = -1; dy <= 1; ++dy) {
    const int hh = h + dy;
    if (hh < 0 || hh >= H) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int ww = w + dx;
      if (ww < 0 || ww >= W) continue;
      const float* xin = state + ((long long)p + (long long)dy * W + dx) * C + Ch;
      const float* kk = k1 + ((dy + 1) * 3 + (dx + 1)) * Ch * F + f;
      for (int ci = 0; ci < Ch; ++ci) acc = fmaf(__ldg(xin + ci), kk[ci * F], acc);
    }
  }
  a1[(long long)p * F + f] = fmaxf(acc, 0.f);
}

// Tiled SGEMM  out[M,Nn] = epi( (sa[k]*A[m][k]+oa[k]) . B[k][n] )
//   kEpi 0: relu(acc + bias[n])        kEpi 1: acc * scale[n] * (mask[m][n] > 0)
// 128 x 128 x 16 tiles, 256 threads, 8 x 8 outputs per thread (two 4-wide strips in each direction so that every
// shared-memory read is a float4), next tile prefetched into registers while the current one is multiplied.  Every
// output is one fmaf chain over ascending k (the summation order of the reference's fp32 matmul restatement).
template <int kEpi>
__global__ void __launch_bounds__(256) k_sgemm(const float* __restrict__ A, const float* __restrict__ sa,
                                               const float* __restrict__ oa, const float* __restrict__ B,
                                               const float* __restrict__ vec, const float* __restrict__ mask,
                                               float* __restrict__ out, long long M, int Nn, int K) {
  constexpr int BM = 128, BN = 128, BK = 16, LDA = BM + 4;
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const long long m0 = (long long)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  // loaders: A tile = 128 rows x 4 float4 (k), B tile = 16 rows (k) x 32 float4 (n); two of each per thread
  const int ar = tid / 4, ak = (tid % 4) * 4;           // rows ar and ar + 64
  const int bk = tid / 32, bn = (tid % 32) * 4;         // k rows bk and bk + 8
  float4 pa[2], pb[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long m = m0 + ar + 64 * i;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M) {
        v = *reinterpret_cast<const float4*>(A + m * K + k0 + ak);
        if (sa != nullptr) {
          const float4 s4 = *reinterpret_cast<const float4*>(sa + k0 + ak), o4 = *reinterpret_cast<const float4*>(oa + k0 + ak);
          v.x = s4.x * v.x + o4.x; v.y = s4.y * v.y + o4.y; v.z = s4.z * v.z + o4.z; v.w = s4.w * v.w + o4.w;
        }
      }
      pa[i] = v;
      pb[i] = n0 + bn < Nn ? *reinterpret_cast<const float4*>(B + (long long)(k0 + bk + 8 * i) * Nn + n0 + bn)
                           : make_float4(0.f, 0.f, 0.f, 0.f);      // partial column tile (Nn % 4 == 0)
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      As[buf][ak + 0][ar + 64 * i] = pa[i].x; As[buf][ak + 1][ar + 64 * i] = pa[i].y;
      As[buf][ak + 2][ar + 64 * i] = pa[i].z; As[buf][ak + 3][ar + 64 * i] = pa[i].w;
      *reinterpret_cast<float4*>(&Bs[buf][bk + 8 * i][bn]) = pb[i];
    }
  };
  float acc[8][8] = {};
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += BK) {
    const bool more = k0 + BK < K;
    if (more) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tiles(buf ^ 1);        // the other buffer was last read one iteration ago (before the previous barrier)
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      if (n >= Nn) continue;
      float4 v = make_float4(acc[i][4 * jh], acc[i][4 * jh + 1], acc[i][4 * jh + 2], acc[i][4 * jh + 3]);
      const float4 w4 = *reinterpret_cast<const float4*>(vec + n);
      if constexpr (kEpi == 0) {
        v.x = fmaxf(v.x + w4.x, 0.f); v.y = fmaxf(v.y + w4.y, 0.f); v.z = fmaxf(v.z + w4.z, 0.f); v.w = fmaxf(v.w + w4.w, 0.f);
      } else {
        const float4 mk = *reinterpret_cast<const float4*>(mask + m * Nn + n);
        v.x = mk.x > 0.f ? v.x * w4.x : 0.f; v.y = mk.y > 0.f ? v.y * w4.y : 0.f;
        v.z = mk.z > 0.f ? v.z * w4.z : 0.f; v.w = mk.w > 0.f ? v.w * w4.w : 0.f;
      }
      *reinterpret_cast<float4*>(out + m * Nn + n) = v;
    }
  }
}

// r[p][c] = c3[c] + sum_{tap valid} sum_k (g2[k]*a2[q][k]+b2[k]) * K3[tap][k][c]    (one warp per pixel)
template <int C>
__global__ void __launch_bounds__(256) k_conv3(const float* __restrict__ a2, const float* __restrict__ g2,
                                               const float* __restrict__ b2, const float* __restrict__ k3,
                                               const float* __restrict__ c3, float* __restrict__ r, int N, int H,
                                               int W, int F) {
  const int lane = threadIdx.x & 31;
  long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  long long M = (long long)N * H * W;
  if (p >= M) return;
  int w = p % W;
  int h = (p / W) % H;
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  for (int dy = -1; dy <= 1; ++dy) {
    int hh = h + dy;
    if (hh < 0 || hh >= H) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      int ww = w + dx;
      if (ww < 0 || ww >= W) continue;
      const float* hin = a2 + (p + (long long)dy * W + dx) * F;
      const float* kk = k3 + (long long)((dy + 1) * 3 + (dx + 1)) * F * C;
      for (int k = lane; k < F; k += 32) {
        float hv = g2[k] * hin[k] + b2[k];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = fmaf(hv, kk[k * C + c], acc[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float v = warp_sum(acc[c]);
    if (lane == 0) r[p * C + c] = v + c3[c];
  }
}

// gp2[p][k] = g2[k] * [a2[p][k] > 0] * sum_{tap} sum_c gr[p - off(tap)][c] * K3[tap][k][c]
__global__ void __launch_bounds__(256) k_conv3_bwd(const float* __restrict__ gr, const float* __restrict__ k3,
                                                   const float* __restrict__ g2, const float* __restrict__ a2,
                                                   float* __restrict__ gp2, int H, int W, int C, int F) {
  // grid (pixel, F / 256), as k_conv1
  const unsigned pu = blockIdx.x;
  const int k = blockIdx.y * 256 + threadIdx.x;
  if (k >= F) return;
  const long long p = pu, idx = p * F + k;
  if (!(a2[idx] > 0.f)) { gp2[idx] = 0.f; return; }
  const int w = (int)(pu % (unsigned)W), h = (int)((pu / (unsigned)W) % (unsigned)H);
  float acc = 0.f;
  for (int dy = -1; dy <= 1; ++dy) {
    int hh = h - dy;
    if (hh < 0 || hh >= H) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      int ww = w - dx;
      if (ww < 0 || ww >= W) continue;
      const float* g = gr + (p - (long long)dy * W - dx) * C;
      const float* kk = k3 + ((long long)((dy + 1) * 3 + (dx + 1)) * F + k) * C;
      for (int c = 0; c < C; ++c) acc = fmaf(g[c], kk[c], acc);
    }
  }
  gp2[idx] = acc * g2[k];
}

// gxb[p][ci] = sum_{tap} sum_f gp1[p - off(tap)][f] * K1[tap][ci][f]     (one warp per pixel)
template <int Ch>
__global__ void __launch_bounds__(256) k_conv1_bwd(const float* __restrict__ gp1, const float* __restrict__ k1,
                                                   float* __restrict__ gxb, int N, int H, int W, int F) {
  const int lane = threadIdx.x & 31;
  long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  long long M = (long long)N * H * W;
  if (p >= M) return;
  int w = p % W;
  int h = (p / W) % H;
  float acc[Ch];
#pragma unroll
  for (int c = 0; c < Ch; ++c) acc[c] = 0.f;
  for (int dy = -1; dy <= 1; ++dy) {
    int hh = h - dy;
    if (hh < 0 || hh >= H) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      int ww = w - dx;
      if (ww < 0 || ww >= W) continue;
      const float* g = gp1 + (p - (long long)dy * W - dx) * F;
      const float* kk = k1 + (long long)((dy + 1) * 3 + (dx + 1)) * Ch * F;
      for (int f = lane; f < F; f += 32) {
        float gv = g[f];
#pragma unroll
        for (int c = 0; c < Ch; ++c) acc[c] = fmaf(gv, kk[c * F + f], acc[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < Ch; ++c) {
    float v = warp_sum(acc[c]);
    if (lane == 0) gxb[p * Ch + c] = v;
  }
}

}  // namespace

#define DISPATCH_NNC(C, ...)                                                                 \
  switch (C) {                                                                               \
    case 1: { constexpr int kC = 1; __VA_ARGS__; } break;                                    \
    case 2: { constexpr int kC = 2; __VA_ARGS__; } break;                                    \
    case 4: { constexpr int kC = 4; __VA_ARGS__; } break;                                    \
    case 8: { constexpr int kC = 8; __VA_ARGS__; } break;                                    \
    case 16: { constexpr int kC = 16; __VA_ARGS__; } break;                                  \
    case 32: { constexpr int kC = 32; __VA_ARGS__; } break;                                  \
    default: throw Error(ASEP_ERR_UNSUPPORTED, strfmt("channel count %d not built", C));     \
  }

void nn_fp32_forward(const NNWeightsF32& w, const float* state, float* a1, float* a2, float* r, int N, int H, int W,
                     int C, int F, cudaStream_t s) {
  long long M = (long long)N * H * W;
  if (M == 0) return;
  ASEP_CHECK(F % 64 == 0, ASEP_ERR_UNSUPPORTED, "n_filters must be a multiple of 64 (got %d)", F);
  ASEP_CHECK(M < (1ll << 31), ASEP_ERR_UNSUPPORTED, "more than 2^31 pixels in one coupling-network launch");
  k_conv1<<<dim3((unsigned)M, (unsigned)cdiv(F, 256)), 256, 0, s>>>(state, w.k1, w.c1, a1, H, W, C, F);
  ASEP_LAUNCH_CHECK();
  dim3 grid((unsigned)cdiv(F, 128), (unsigned)cdiv(M, 128));
  k_sgemm<0><<<grid, 256, 0, s>>>(a1, w.g1, w.b1, w.k2, w.c2, nullptr, a2, M, F, F);
  ASEP_LAUNCH_CHECK();
  DISPATCH_NNC(C, (k_conv3<kC><<<cdiv(M * 32, 256), 256, 0, s>>>(a2, w.g2, w.b2, w.k3, w.c3, r, N, H, W, F)));
  ASEP_LAUNCH_CHECK();
}

void nn_fp32_backward(const NNWeightsF32& w, const float* a1, const float* a2, const float* gr, float* t1, float* t2,
                      float* gxb, int N, int H, int W, int C, int F, cudaStream_t s) {
  long long M = (long long)N * H * W;
  if (M == 0) return;
  // t2 = gp2 = conv3^T(gr) * g2 * [p2 > 0]
  ASEP_CHECK(M < (1ll << 31), ASEP_ERR_UNSUPPORTED, "more than 2^31 pixels in one coupling-network launch");
  k_conv3_bwd<<<dim3((unsigned)M, (unsigned)cdiv(F, 256)), 256, 0, s>>>(gr, w.k3, w.g2, a2, t2, H, W, C, F);
  ASEP_LAUNCH_CHECK();
  // t1 = gp1 = (gp2 . K2^T) * g1 * [p1 > 0]
  dim3 grid((unsigned)cdiv(F, 128), (unsigned)cdiv(M, 128));
  k_sgemm<1><<<grid, 256, 0, s>>>(t2, nullptr, nullptr, w.k2t, w.g1, a1, t1, M, F, F);
  ASEP_LAUNCH_CHECK();
  DISPATCH_NNC(C / 2, (k_conv1_bwd<kC><<<cdiv(M * 32, 256), 256, 0, s>>>(t1, w.k1, gxb, N, H, W, F)));
  ASEP_LAUNCH_CHECK();
}

}  // namespace asep
