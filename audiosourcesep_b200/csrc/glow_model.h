// Glow prior: parameters, derived per-step constants, workspace and the forward / inverse /
// log_prob / grad_log_prob orchestration over the kernels (reference: flow_models/flow_glow.py,
// flow_models/flow_builder.py:60-146, run_basis_sep.py:73-79).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "kernels.h"
#include "nn_tc.h"

namespace asep {

struct Param {
  std::vector<int64_t> shape;
  std::vector<float> host;
  float* dev = nullptr;
  int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

struct Level { int H, W, C; };

struct StepDerived {
  float* sc = nullptr;          // element-wise constants (step_const_floats(C))
  double logdet_const = 0.0;    // H*W*(sum log_scale + sum log_S)
  float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr, *k2t = nullptr;
  NNWeightsF32 w32{};
  NNWeightsTC wtc{};
};

class DeviceArena {
 public:
  ~DeviceArena() { release(); }
  void reserve(size_t bytes);
  void reset() { off_ = 0; }
  void* take(size_t bytes);
  template <typename T> T* take_n(size_t n) { return reinterpret_cast<T*>(take(n * sizeof(T))); }
  void release();
  size_t capacity() const { return cap_; }
 private:
  char* base_ = nullptr;
  size_t cap_ = 0, off_ = 0;
};

class GlowModel {
 public:
  explicit GlowModel(const asep_glow_cfg& cfg, int device);
  ~GlowModel();

  void set_param(const std::string& name, const float* src, const std::vector<int64_t>& shape, bool src_on_device);
  const Param& get_param(const std::string& name) const;
  void prepare(int precision);
  void init_actnorm(const float* minibatch, int N, cudaStream_t s);

  void forward(const float* x, float* z, float* fldj, int N, cudaStream_t s);
  void inverse(const float* z, float* x, int N, cudaStream_t s);
  void log_prob(const float* x, float* logp, int N, cudaStream_t s);
  void grad_log_prob(const float* x, float* grad, float* logp, int N, cudaStream_t s);
  void sample(const float* eps, float* x, int N, cudaStream_t s);
  void coupling_nn(int block, int step, const float* state, float* r, int N, cudaStream_t s);
  void coupling_nn_backward(int block, int step, const float* state, const float* gr, float* gxb, int N,
                            cudaStream_t s);

  const asep_glow_cfg& cfg() const { return cfg_; }
  int device() const { return device_; }
  const Level& level(int b) const { return levels_[b]; }
  int latent_dims() const { return Dl_; }
  int precision() const { return precision_; }

 private:
  struct Work {
    int N = 0;
    bool save = false;
    std::vector<float*> X, O;               // per level: block input / output state
    std::vector<std::vector<float*>> U, R;  // per level, per step (save) or 2 ping-pong / 1 (no save)
    float *z = nullptr, *gz = nullptr, *gA = nullptr, *gB = nullptr, *gr = nullptr, *gu = nullptr, *gxb = nullptr;
    double *acc_ld = nullptr, *acc_prior = nullptr;
    float *a1 = nullptr, *a2 = nullptr, *t1 = nullptr, *t2 = nullptr;   // fp32 NN scratch
    NNScratchTC tc{};
  };
  void ensure_work(int N, bool save);
  StepDerived& step(int b, int k) { return steps_[(size_t)b * cfg_.K + k]; }
  void run_forward(const float* x, int N, bool save, cudaStream_t s);
  void nn_forward(int b, int k, const float* state, float* r, int N, bool save, cudaStream_t s);
  void nn_backward(int b, int k, const float* state, const float* gr, float* gxb, int N, cudaStream_t s);
  void build_step_consts(int b, int k);
  void require_prepared() const;
  double const_logdet() const;
  void latent_slice(int b, int& Cz, int& nb, int& coff) const;

  asep_glow_cfg cfg_;
  int device_;
  std::vector<Level> levels_;
  int Hl_, Wl_, CL_, Dl_;
  std::map<std::string, Param> params_;
  std::vector<StepDerived> steps_;
  bool prepared_ = false;
  int precision_ = ASEP_PREC_FP32;
  DeviceArena arena_;
  Work work_;
};

}  // namespace asep
