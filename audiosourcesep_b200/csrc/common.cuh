// Shared host/device helpers of libasep.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/asep.h"

namespace asep {

// ---- error plumbing: C++ exceptions inside, status codes at the ABI
struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

std::string strfmt(const char* fmt, ...);
void set_last_error(const std::string& m);
extern std::atomic<long long> g_launch_count;

#define ASEP_CHECK(cond, code, ...)                                   \
  do {                                                                \
    if (!(cond)) throw ::asep::Error((code), ::asep::strfmt(__VA_ARGS__)); \
  } while (0)

#define CUDA_CHECK(expr)                                                                           \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      throw ::asep::Error(ASEP_ERR_CUDA, ::asep::strfmt("%s failed: %s (%s:%d)", #expr,            \
                                                        cudaGetErrorString(_e), __FILE__, __LINE__)); \
  } while (0)

// Every kernel launch of the library goes through this so that `gpu_launches` is a count, not a guess.
#define ASEP_LAUNCH_CHECK()                      \
  do {                                           \
    ::asep::g_launch_count.fetch_add(1);         \
    CUDA_CHECK(cudaGetLastError());              \
  } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- measurement aid for the HBM-bound kernels (bench.py `roofline_hbm`): while switched on (asep_hbm_profile), every
// launch of a profiled category is bracketed by a CUDA event pair on its own stream and its ALGORITHMIC bytes are
// summed.  Off (the default) a scope costs one branch.
enum HbmCat {
  kHbmFlowStep = 0,   // k_pre / k_post_pre / k_inv_step / k_bwd_*: ActNorm + 1x1 + coupling (+ fused col2im)
  kHbmLangevin = 1,   // k_langevin
  kHbmPrep = 2,       // k_prep: normalise (+ ELU) + bf16 cast of a convolution input
  kHbmPool = 3,       // k_pool5_1d / k_avgpool2 / k_resize2x_add
  kHbmGather = 4,     // k_gather_* (col2im of the per-tap outputs; only where not fused)
  kHbmCatCount = 5
};
struct HbmScope {
  HbmScope(int cat, double bytes, cudaStream_t s);
  ~HbmScope();
  int cat_; cudaStream_t s_; void* rec_;
};
void hbm_profile(bool on);
bool hbm_profile_enabled();
void hbm_profile_read(int cat, double* ms, long long* launches, double* bytes);

// ---- DLTensor validation
struct TView {
  float* f32 = nullptr;
  void* raw = nullptr;
  int ndim = 0;
  int64_t shape[8] = {0};
  int64_t numel = 0;
  bool on_device = false;
};
TView view_f32(const DLTensor* t, const char* what, int device, bool allow_host = false);
TView view_i32(const DLTensor* t, const char* what, int device);
void expect_shape(const TView& v, const char* what, std::initializer_list<int64_t> shp);

// ---- device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace asep
