// NCSN v1 (CondRefineNetDilated, ncsn/score_network.py:224-296) and v2 (RefineNetDilated,
// ncsn/score_network_v2.py:202-278) score networks: parameters, prepared tcgen05 weight images and the forward
// graph `model([x, sigma_idx], training=True) -> score`.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "conv_tc.h"
#include "ncsn_kernels.h"

namespace asep {

struct NcsnParam {
  std::vector<int64_t> shape;
  std::vector<float> host;
  float* dev = nullptr;
};

class NcsnModel {
 public:
  NcsnModel(const asep_ncsn_cfg& cfg, int device);
  ~NcsnModel();
  void set_param(const std::string& name, const float* src, const std::vector<int64_t>& shape, bool on_device);
  void set_sigmas(const float* sigmas, int n);
  void prepare();
  // split-bf16 mode (ASEP_PREC_BF16X3): takes effect at the next prepare()
  void set_precision(bool x3) { if (x3 != x3_) { x3_ = x3; prepared_ = false; } }
  // x [N,H,W,1] fp32, idx [N] int32 -> score [N,H,W,1] fp32
  void forward(const float* x, const int* idx, float* score, int N, cudaStream_t s);
  const asep_ncsn_cfg& cfg() const { return cfg_; }
  int device() const { return device_; }
  int64_t num_params() const;
  // persistent scratch of the BASIS inner loop: a score tensor [N,H,W,1] and a constant sigma index vector [N]
  float* score_scratch(int N, int slots = 1);   // slots = 2: room for both sources when one handle serves as both priors
  const int* index_scratch(int N, int sigma_idx, cudaStream_t s);
  // identity of the device allocations a captured graph of this model's kernels would bake in (api.cu: BASIS step graphs)
  long long uid() const { return uid_; }
  long long generation() const { return generation_; }

 private:
  // fp32 NHWC activation (batch = N_); sums != NULL: per-(n,c) sum / sum of squares already accumulated by the producer
  // bf != NULL: bf16 copy written by the producer (saves the cast pass when no norm intervenes)
  struct T { float* p = nullptr; int H = 0, W = 0, C = 0; double* sums = nullptr; __nv_bfloat16* bf = nullptr; };
  const NcsnParam& param(const std::string& name) const;
  bool has(const std::string& name) const { return params_.count(name) != 0; }
  void* take(size_t bytes);
  T new_t(int H, int W, int C);
  __nv_bfloat16* new_bf(int H, int W, int C);
  const float2* norm_coef(const T& x, const std::string& name);
  struct BF { __nv_bfloat16* hi = nullptr; __nv_bfloat16* lo = nullptr; };   // convolution operand (lo only in the x3 mode)
  BF prep(const T& x, const float2* coef, bool elu);
  T conv(const std::string& name, const BF& xin, int H, int W, const float* add, bool stats, bool bf16_copy = false);
  T res_block(const T& x, const std::string& name, int cout, bool down, int dilation);
  T rcu(T x, const std::string& prefix, int n_blocks, int n_stages);
  T crp(T x, const std::string& prefix);
  T msf(const std::vector<T>& xs, const std::string& prefix, int H, int W, int features);
  T refine(const std::vector<T>& xs, const std::string& name, int features, bool end, int H, int W);
  void run(const float* x, const int* idx, float* score);

  asep_ncsn_cfg cfg_;
  int device_;
  bool v1_;
  std::map<std::string, NcsnParam> params_;
  std::map<std::string, ConvWeightsTC> convs_;
  std::map<std::string, ConvWeightsTC> convs_lo_;  // x3 mode: tile images of w - bf16(w)
  bool x3_ = false;
  std::map<std::string, float*> gab_;            // v2: packed [gamma|alpha|beta] rows per norm layer
  float* sigmas_dev_ = nullptr;
  int n_sigmas_ = 0;
  bool prepared_ = false;
  // per-call state
  bool dry_ = false;
  int N_ = 0;
  const int* idx_ = nullptr;
  cudaStream_t s_ = nullptr;
  float* score_buf_ = nullptr;
  int* idx_buf_ = nullptr;
  size_t score_cap_ = 0, idx_cap_ = 0;
  std::vector<int> idx_host_;
  char* arena_ = nullptr;
  size_t arena_cap_ = 0, arena_off_ = 0;
  long long uid_ = 0, generation_ = 0;
};

}  // namespace asep
