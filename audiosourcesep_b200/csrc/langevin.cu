// Fused BASIS Langevin update (run_basis_sep.py:163-181, dB mixing :131-147): both sources, the
// mixture-consistency gradient, noise injection and the step-size scaling in ONE HBM-bound pass:
// 5 reads + 2 writes of 4 B per element with in-kernel Philox noise (7 reads with injected noise).
#include "kernels.h"

namespace asep {
namespace {

// Philox4x32-10 (Salmon et al. 2011), counter = (elem_lo, elem_hi, step, stream), key = seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}

__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  // u1 in (0,1], u2 in [0,1)
  float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
  float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
  float rad = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(rad * c, rad * s);
}

// Four standard normals for the group of 4 consecutive elements starting at global index 4*g.
__device__ __forceinline__ float4 normal4(uint64_t seed, uint64_t step, uint32_t stream_id, uint64_t group) {
  uint4 ctr = make_uint4((uint32_t)group, (uint32_t)(group >> 32), (uint32_t)step,
                         (uint32_t)(step >> 32) ^ (stream_id * 0x9E3779B9u));
  uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
  return make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ void mix_db(float x1, float x2, float& g, float& w1, float& w2) {
  const float k = 0.23025850929940458f;      // ln(10)/10
  float a1 = x1 * k, a2 = x2 * k;
  float m = fmaxf(a1, a2);
  float e1 = expf(a1 - m), e2 = expf(a2 - m);
  float sum = e1 + e2;
  g = (1.0f / k) * (m + logf(sum) - 0.69314718055994531f);   // (10/ln10)(logsumexp - ln 2)
  w1 = e1 / sum;
  w2 = e2 / sum;
}

__device__ __forceinline__ float get4(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// kDev: the scalars that change from one Langevin step to the next (step size, Philox step number, position in the injected
// noise / per-step dump tensors) are read from device memory, so that ONE captured CUDA graph of a whole BASIS step can
// be replayed for every step of every noise level (api.cu: basis graphs).
template <bool kInjected, bool kDev = false>
__global__ void __launch_bounds__(256) k_langevin(float* __restrict__ x1, float* __restrict__ x2,
                                                  const float* __restrict__ s1, const float* __restrict__ s2,
                                                  const float* __restrict__ mixed, const float* __restrict__ n1,
                                                  const float* __restrict__ n2, float eta, float lambda,
                                                  float noise_scale, uint64_t seed, uint64_t step,
                                                  uint64_t elem_offset, int* __restrict__ nan_count, long long n4,
                                                  const LangevinDev* __restrict__ dev = nullptr, float* __restrict__ dump = nullptr) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  if constexpr (kDev) {
    eta = dev->eta; lambda = dev->lambda; noise_scale = dev->noise_scale; step = dev->step;
    const long long t = (long long)dev->t;
    if constexpr (kInjected) { n1 += t * n4 * 4; n2 += t * n4 * 4; }
    if (dump != nullptr) dump += 2 * t * n4 * 4;
  }
  float4 a1 = reinterpret_cast<const float4*>(x1)[i], a2 = reinterpret_cast<const float4*>(x2)[i];
  float4 g1 = reinterpret_cast<const float4*>(s1)[i], g2 = reinterpret_cast<const float4*>(s2)[i];
  float4 mx = reinterpret_cast<const float4*>(mixed)[i];
  float4 z1, z2;
  if constexpr (kInjected) {
    z1 = reinterpret_cast<const float4*>(n1)[i];
    z2 = reinterpret_cast<const float4*>(n2)[i];
  } else {
    uint64_t group = (elem_offset >> 2) + (uint64_t)i;
    z1 = normal4(seed, step, 1u, group);
    z2 = normal4(seed, step, 2u, group);
  }
  float o1[4], o2[4];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float g, w1, w2;
    float v1 = get4(a1, j), v2 = get4(a2, j);
    mix_db(v1, v2, g, w1, w2);
    float resid = get4(mx, j) - g;
    // evaluation order of run_basis_sep.py:180-181
    o1[j] = v1 + eta * (get4(g1, j) + lambda * w1 * resid) + noise_scale * get4(z1, j);
    o2[j] = v2 + eta * (get4(g2, j) + lambda * w2 * resid) + noise_scale * get4(z2, j);
    bad |= (o1[j] != o1[j]) || (o2[j] != o2[j]);
  }
  reinterpret_cast<float4*>(x1)[i] = make_float4(o1[0], o1[1], o1[2], o1[3]);
  reinterpret_cast<float4*>(x2)[i] = make_float4(o2[0], o2[1], o2[2], o2[3]);
  if constexpr (kDev) {
    if (dump != nullptr) {                       // per-step state dump [T, 2, ...]
      reinterpret_cast<float4*>(dump)[i] = make_float4(o1[0], o1[1], o1[2], o1[3]);
      reinterpret_cast<float4*>(dump)[n4 + i] = make_float4(o2[0], o2[1], o2[2], o2[3]);
    }
  }
  if (nan_count != nullptr && bad) atomicAdd(nan_count, 1);
}

__global__ void k_langevin_advance(LangevinDev* dev) {
  dev->step += 1;
  dev->t += 1;
}

__global__ void k_mixing_db(const float* __restrict__ x1, const float* __restrict__ x2, float* __restrict__ g,
                            float* __restrict__ w1, float* __restrict__ w2, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gg, a, b;
  mix_db(x1[i], x2[i], gg, a, b);
  g[i] = gg; w1[i] = a; w2[i] = b;
}

__global__ void k_philox_normal(float* __restrict__ out, uint64_t seed, uint64_t step, uint32_t stream_id,
                                uint64_t elem_offset, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  reinterpret_cast<float4*>(out)[i] = normal4(seed, step, stream_id, (elem_offset >> 2) + (uint64_t)i);
}

}  // namespace

void launch_langevin(float* x1, float* x2, const float* s1, const float* s2, const float* mixed, const float* n1,
                     const float* n2, float eta, float lambda, float noise_scale, uint64_t seed, uint64_t step,
                     uint64_t elem_offset, int* nan_count, long long n, cudaStream_t s) {
  if (n == 0) return;
  ASEP_CHECK(n % 4 == 0 && elem_offset % 4 == 0, ASEP_ERR_BAD_SHAPE,
             "langevin: element count %lld and offset must be multiples of 4", n);
  // 5 reads + 2 writes of 4 bytes per element with in-kernel noise, 7 reads + 2 writes with injected noise
  HbmScope prof(kHbmLangevin, (n1 ? 36.0 : 28.0) * (double)n, s);
  long long n4 = n / 4;
  if (n1 != nullptr && n2 != nullptr)
    k_langevin<true><<<cdiv(n4, 256), 256, 0, s>>>(x1, x2, s1, s2, mixed, n1, n2, eta, lambda, noise_scale, seed,
                                                    step, elem_offset, nan_count, n4);
  else
    k_langevin<false><<<cdiv(n4, 256), 256, 0, s>>>(x1, x2, s1, s2, mixed, nullptr, nullptr, eta, lambda,
                                                     noise_scale, seed, step, elem_offset, nan_count, n4);
  ASEP_LAUNCH_CHECK();
}

void launch_langevin_dev(float* x1, float* x2, const float* s1, const float* s2, const float* mixed, const float* n1,
                         const float* n2, float* dump, LangevinDev* dev, uint64_t seed, uint64_t elem_offset, int* nan_count,
                         long long n, cudaStream_t s) {
  if (n == 0) return;
  ASEP_CHECK(n % 4 == 0 && elem_offset % 4 == 0, ASEP_ERR_BAD_SHAPE,
             "langevin: element count %lld and offset must be multiples of 4", n);
  const long long n4 = n / 4;
  if (n1 != nullptr && n2 != nullptr)
    k_langevin<true, true><<<cdiv(n4, 256), 256, 0, s>>>(x1, x2, s1, s2, mixed, n1, n2, 0.f, 0.f, 0.f, seed, 0, elem_offset, nan_count,
                                                          n4, dev, dump);
  else
    k_langevin<false, true><<<cdiv(n4, 256), 256, 0, s>>>(x1, x2, s1, s2, mixed, nullptr, nullptr, 0.f, 0.f, 0.f, seed, 0, elem_offset,
                                                           nan_count, n4, dev, dump);
  ASEP_LAUNCH_CHECK();
  k_langevin_advance<<<1, 1, 0, s>>>(dev);
  ASEP_LAUNCH_CHECK();
}

void launch_mixing_db(const float* x1, const float* x2, float* g, float* w1, float* w2, long long n, cudaStream_t s) {
  if (n == 0) return;
  k_mixing_db<<<cdiv(n, 256), 256, 0, s>>>(x1, x2, g, w1, w2, n);
  ASEP_LAUNCH_CHECK();
}

void launch_philox_normal(float* out, uint64_t seed, uint64_t step, uint64_t stream_id, uint64_t elem_offset,
                          long long n, cudaStream_t s) {
  if (n == 0) return;
  ASEP_CHECK(n % 4 == 0 && elem_offset % 4 == 0, ASEP_ERR_BAD_SHAPE, "philox: counts must be multiples of 4");
  k_philox_normal<<<cdiv(n / 4, 256), 256, 0, s>>>(out, seed, step, (uint32_t)stream_id, elem_offset, n / 4);
  ASEP_LAUNCH_CHECK();
}

}  // namespace asep
