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
  long long flat_off = -1;        // >= 0 once training is enabled: dev points into the flat parameter vector
};

class NcsnModel {
 public:
  NcsnModel(const asep_ncsn_cfg& cfg, int device);
  ~NcsnModel();
  void set_param(const std::string& name, const float* src, const std::vector<int64_t>& shape, bool on_device);
  void set_sigmas(const float* sigmas, int n);
  void prepare();
  // split-bf16 mode (ASEP_PREC_BF16X3): takes effect at the next prepare()
  void set_precision(bool x3) {
    ASEP_CHECK(!training_ || x3 == x3_, ASEP_ERR_STATE, "the precision mode is fixed once training is enabled");
    if (x3 != x3_) { x3_ = x3; prepared_ = false; }
  }
  // x [N,H,W,1] fp32, idx [N] int32 -> score [N,H,W,1] fp32
  void forward(const float* x, const int* idx, float* score, int N, cudaStream_t s);
  const asep_ncsn_cfg& cfg() const { return cfg_; }
  int device() const { return device_; }
  int64_t num_params() const;
  // persistent scratch of the BASIS inner loop: a score tensor [N,H,W,1] and a constant sigma index vector [N]
  float* score_scratch(int N, int slots = 1);   // slots = 2: room for both sources when one handle serves as both priors
  const int* index_scratch(int N, int sigma_idx, cudaStream_t s);
  // ---- training: denoising score matching (train_ncsn.py:26-57) -- ncsn_train.cu
  // Moves every parameter into one flat fp32 device vector (name order; every tensor starts on a 16-byte boundary),
  // allocates the Adam moments and the data-gradient weight images.
  void enable_training();
  bool training() const { return training_; }
  long long num_trainable() const { return n_flat_; }
  // offset / element count of a parameter inside the flat vector (checkpointing, tests)
  void param_span(const std::string& name, long long* offset, long long* numel) const;
  // x [N,H,W,1], noise [N,H,W,1] standard normal, idx [N] int32 (all device).  grads [num_trainable] <- d loss / d theta,
  // loss [1] <- sum_n 1/2 ||score(x + sigma_n noise, idx_n) + noise / sigma_n||^2 sigma_n^2 / global_batch
  void train_grads(const float* x, const float* noise, const int* idx, int N, int global_batch, float* grads, float* loss,
                   cudaStream_t s);
  void adam_step(const float* grads, float lr, float beta1, float beta2, float eps, cudaStream_t s);   // Keras Adam
  void copy_flat(float* dst, cudaStream_t s) const;
  void set_flat(const float* src, cudaStream_t s);
  void sync_host();                                // flat vector -> host copies (get_param / prepare see trained values)
  const NcsnParam& get_param(const std::string& name) const { return param(name); }

  // identity of the device allocations a captured graph of this model's kernels would bake in (api.cu: BASIS step graphs)
  long long uid() const { return uid_; }
  long long generation() const { return generation_; }

 private:
  // fp32 NHWC activation (batch = N_); sums != NULL: per-(n,c) sum / sum of squares already accumulated by the producer
  // bf != NULL: bf16 copy written by the producer (saves the cast pass when no norm intervenes)
  // g (training): gradient of the loss w.r.t. this tensor, accumulated by every consumer's backward pass
  struct T { float* p = nullptr; int H = 0, W = 0, C = 0; double* sums = nullptr; __nv_bfloat16* bf = nullptr; float* g = nullptr; };
  const NcsnParam& param(const std::string& name) const;
  bool has(const std::string& name) const { return params_.count(name) != 0; }
  void* take(size_t bytes);
  T new_t(int H, int W, int C);
  __nv_bfloat16* new_bf(int H, int W, int C);
  // folded instance-norm++ of one layer: per-(n,c) coefficients, the statistics they came from, the layer name
  struct Norm { const float2* coef = nullptr; const double* sums = nullptr; std::string name; };
  Norm norm_coef(const T& x, const std::string& name);
  // convolution operand (lo only in the x3 mode); gy (training): fp32 gradient w.r.t. the operand
  struct BF { __nv_bfloat16* hi = nullptr; __nv_bfloat16* lo = nullptr; float* gy = nullptr; };
  // stat_src: the tensor the statistics of `norm` were taken from when it is not x itself (CRP: pool first, normalise after)
  BF prep(const T& x, const Norm& norm, bool elu, const T* stat_src = nullptr);
  T conv(const std::string& name, const BF& xin, int H, int W, const float* add, bool stats, bool bf16_copy = false,
         const float* add2 = nullptr);
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

  // ---- training state (ncsn_train.cu)
  struct Op {
    enum Kind { kBegin, kPrep, kConv, kAvgPool2, kPool5, kResizeAdd, kElu, kAdd, kEnd } kind;
    T a, b, out;               // kPrep: a = value tensor, b = statistics source; kResizeAdd: a = low, b = add; kAdd: a + b
    BF bf;                     // kPrep: produced operand; kConv / kEnd: consumed operand
    Norm norm;                 // kPrep
    bool elu = false;
    std::string name;          // kConv: layer name
    const float* add = nullptr;   // kConv: tensor summed in the epilogue
    const float* add2 = nullptr;  // kConv: second tensor summed in the epilogue
  };
  void* take_g(size_t bytes);
  void record(const Op& op) { if (train_ && !dry_) tape_.push_back(op); }
  float* grad_of(const float* p) const;
  float* G(const std::string& name) const;       // slice of the flat gradient vector
  void backward(float* grads);
  void refresh_images(cudaStream_t s);
  bool training_ = false, train_ = false;        // train_: the current run records the tape and allocates gradients
  std::vector<Op> tape_;
  std::map<const float*, float*> grad_of_;
  std::map<std::string, ConvWeightsTC> convs_t_, convs_t_lo_;   // data-gradient operands (transposed, flipped kernels)
  float *theta_ = nullptr, *adam_m_ = nullptr, *adam_v_ = nullptr, *grads_cur_ = nullptr;
  long long n_flat_ = 0, adam_t_ = 0;
  char* garena_ = nullptr;
  size_t garena_cap_ = 0, garena_off_ = 0;
  float *xt_ = nullptr, *tscore_ = nullptr, *gscore_ = nullptr;   // perturbed input, score, d loss / d score
  int train_cap_ = 0;
  double* loss_acc_ = nullptr;
  bool images_dirty_ = false;
};

}  // namespace asep
