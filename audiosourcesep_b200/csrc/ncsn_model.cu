#include "ncsn_model.h"

#include <cstring>

#include <atomic>

namespace asep {

namespace { std::atomic<long long> g_next_ncsn_uid{1}; }

NcsnModel::NcsnModel(const asep_ncsn_cfg& cfg, int device) : cfg_(cfg), device_(device), v1_(cfg.version == 1) {
  uid_ = g_next_ncsn_uid.fetch_add(1);
  ASEP_CHECK(cfg.version == 1 || cfg.version == 2, ASEP_ERR_BAD_ARG, "NCSN version must be 1 or 2");
  ASEP_CHECK(cfg.C == 1, ASEP_ERR_UNSUPPORTED, "the score networks of the separation path take 1-channel patches");
  ASEP_CHECK(cfg.ngf % 64 == 0, ASEP_ERR_UNSUPPORTED, "n_filters must be a multiple of 64 (got %d)", cfg.ngf);
  ASEP_CHECK(conv_tc_supported(cfg.ngf, cfg.ngf, cfg.H, cfg.W) && cfg.H % 2 == 0 && cfg.W % 2 == 0 &&
                 conv_tc_supported(2 * cfg.ngf, 2 * cfg.ngf, cfg.H / 2, cfg.W / 2),
             ASEP_ERR_UNSUPPORTED, "patch %dx%d with %d filters is outside the tcgen05 convolution tiling", cfg.H, cfg.W,
             cfg.ngf);
}

NcsnModel::~NcsnModel() {
  for (auto& kv : params_)
    if (kv.second.dev && kv.second.flat_off < 0) cudaFree(kv.second.dev);
  for (auto& kv : convs_) conv_tc_release(kv.second);
  for (auto& kv : convs_lo_) conv_tc_release(kv.second);
  for (auto& kv : convs_t_) conv_tc_release(kv.second);
  for (auto& kv : convs_t_lo_) conv_tc_release(kv.second);
  for (void* p : {(void*)theta_, (void*)adam_m_, (void*)adam_v_, (void*)garena_, (void*)xt_, (void*)tscore_, (void*)gscore_,
                  (void*)loss_acc_})
    if (p) cudaFree(p);
  if (sigmas_dev_) cudaFree(sigmas_dev_);
  if (arena_) cudaFree(arena_);
  if (score_buf_) cudaFree(score_buf_);
  if (idx_buf_) cudaFree(idx_buf_);
}

float* NcsnModel::score_scratch(int N, int slots) {
  const size_t need = (size_t)slots * N * cfg_.H * cfg_.W * cfg_.C * sizeof(float);
  if (need > score_cap_) {
    CUDA_CHECK(cudaDeviceSynchronize());
    if (score_buf_) cudaFree(score_buf_);
    score_buf_ = nullptr;
    CUDA_CHECK(cudaMalloc(&score_buf_, need));
    score_cap_ = need;
    ++generation_;
  }
  return score_buf_;
}

const int* NcsnModel::index_scratch(int N, int sigma_idx, cudaStream_t s) {
  if ((size_t)N > idx_cap_) {
    CUDA_CHECK(cudaDeviceSynchronize());
    if (idx_buf_) cudaFree(idx_buf_);
    idx_buf_ = nullptr;
    CUDA_CHECK(cudaMalloc(&idx_buf_, (size_t)N * sizeof(int)));
    idx_cap_ = (size_t)N;
    ++generation_;
  }
  CUDA_CHECK(cudaStreamSynchronize(s));            // the host staging vector may still feed an earlier copy
  idx_host_.assign((size_t)N, sigma_idx);
  CUDA_CHECK(cudaMemcpyAsync(idx_buf_, idx_host_.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, s));
  return idx_buf_;
}

void NcsnModel::set_param(const std::string& name, const float* src, const std::vector<int64_t>& shape, bool on_device) {
  int64_t n = 1;
  for (auto v : shape) n *= v;
  if (training_) {                                 // the parameter lives in the flat vector: overwrite its slice
    auto it = params_.find(name);
    ASEP_CHECK(it != params_.end() && (int64_t)it->second.host.size() == n, ASEP_ERR_BAD_SHAPE,
               "'%s': a training handle only accepts values of the existing shape", name.c_str());
    NcsnParam& q = it->second;
    if (on_device) CUDA_CHECK(cudaMemcpy(q.host.data(), src, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    else std::memcpy(q.host.data(), src, (size_t)n * sizeof(float));
    CUDA_CHECK(cudaMemcpy(q.dev, q.host.data(), (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    images_dirty_ = true;
    return;
  }
  NcsnParam& p = params_[name];
  if (p.dev && (int64_t)p.host.size() != n) { cudaFree(p.dev); p.dev = nullptr; }
  p.shape = shape;
  p.host.resize((size_t)n);
  if (on_device) CUDA_CHECK(cudaMemcpy(p.host.data(), src, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  else std::memcpy(p.host.data(), src, (size_t)n * sizeof(float));
  if (!p.dev) CUDA_CHECK(cudaMalloc(&p.dev, std::max<int64_t>(n, 1) * sizeof(float)));
  CUDA_CHECK(cudaMemcpy(p.dev, p.host.data(), (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
  prepared_ = false;
}

void NcsnModel::set_sigmas(const float* sigmas, int n) {
  if (sigmas_dev_) cudaFree(sigmas_dev_);
  ++generation_;
  CUDA_CHECK(cudaMalloc(&sigmas_dev_, (size_t)n * sizeof(float)));
  CUDA_CHECK(cudaMemcpy(sigmas_dev_, sigmas, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
  n_sigmas_ = n;
}

int64_t NcsnModel::num_params() const {
  int64_t n = 0;
  for (auto& kv : params_) n += (int64_t)kv.second.host.size();
  return n;
}

const NcsnParam& NcsnModel::param(const std::string& name) const {
  auto it = params_.find(name);
  ASEP_CHECK(it != params_.end(), ASEP_ERR_STATE, "score network parameter '%s' has not been set", name.c_str());
  return it->second;
}

void NcsnModel::prepare() {
  CUDA_CHECK(cudaSetDevice(device_));
  if (training_) {                                 // weights live on the device: only the tile images are derived
    CUDA_CHECK(cudaDeviceSynchronize());
    refresh_images(nullptr);
    return;
  }
  ++generation_;                                   // tile images / packed rows below are re-allocated
  CUDA_CHECK(cudaDeviceSynchronize());             // a graph replay may still be reading them
  for (auto& kv : convs_) conv_tc_release(kv.second);
  convs_.clear();
  for (auto& kv : convs_lo_) conv_tc_release(kv.second);
  convs_lo_.clear();
  for (auto& kv : params_) {
    const std::string& name = kv.first;
    const size_t pos = name.rfind('/');
    const std::string leaf = name.substr(pos + 1), layer = name.substr(0, pos);
    if (leaf == "kernel") {
      if (layer == "begin_conv" || layer == "end_conv") continue;      // CUDA-core kernels (1 channel on one side)
      const auto& sh = kv.second.shape;
      ASEP_CHECK(sh.size() == 4 && sh[0] == sh[1], ASEP_ERR_BAD_SHAPE, "%s: expected a [k,k,Cin,Cout] kernel", name.c_str());
      const int dil = layer.rfind("Res3_", 0) == 0 ? 2 : (layer.rfind("Res4_", 0) == 0 ? 4 : 1);   // score_network.py:259-268
      const float* bias = has(layer + "/bias") ? param(layer + "/bias").host.data() : nullptr;
      conv_tc_prepare(convs_[layer], kv.second.host.data(), bias, (int)sh[0], (int)sh[2], (int)sh[3], dil);
      if (x3_) {                                                        // second term of the split-bf16 weights
        std::vector<float> lo(kv.second.host.size());
        for (size_t i = 0; i < lo.size(); ++i) lo[i] = kv.second.host[i] - __bfloat162float(__float2bfloat16(kv.second.host[i]));
        conv_tc_prepare(convs_lo_[layer], lo.data(), nullptr, (int)sh[0], (int)sh[2], (int)sh[3], dil);
      }
    }
  }
  if (!v1_) ASEP_CHECK(sigmas_dev_ != nullptr, ASEP_ERR_STATE, "NCSN v2 needs the noise levels (asep_ncsn_set_sigmas)");
  prepared_ = true;
  // dry run: checks that every parameter the graph needs is present
  dry_ = true; N_ = 1; arena_off_ = 0;
  run(nullptr, nullptr, nullptr);
  dry_ = false;
}

// ------------------------------------------------------------------ per-call allocation (bump arena sized by a dry run)
void* NcsnModel::take(size_t bytes) {
  const size_t al = (bytes + 1023) & ~(size_t)1023;
  void* p = dry_ ? nullptr : arena_ + arena_off_;
  if (!dry_) ASEP_CHECK(arena_off_ + al <= arena_cap_, ASEP_ERR_STATE, "score-network arena overflow");
  arena_off_ += al;
  return p;
}
void* NcsnModel::take_g(size_t bytes) {
  const size_t al = (bytes + 1023) & ~(size_t)1023;
  void* p = dry_ ? nullptr : garena_ + garena_off_;
  if (!dry_) ASEP_CHECK(garena_off_ + al <= garena_cap_, ASEP_ERR_STATE, "score-network gradient arena overflow");
  garena_off_ += al;
  return p;
}
NcsnModel::T NcsnModel::new_t(int H, int W, int C) {
  T t;
  t.H = H; t.W = W; t.C = C;
  t.p = static_cast<float*>(take((size_t)N_ * H * W * C * sizeof(float)));
  if (train_) {
    t.g = static_cast<float*>(take_g((size_t)N_ * H * W * C * sizeof(float)));
    if (!dry_) grad_of_[t.p] = t.g;
  }
  return t;
}
__nv_bfloat16* NcsnModel::new_bf(int H, int W, int C) {
  return static_cast<__nv_bfloat16*>(take((size_t)N_ * H * W * C * sizeof(__nv_bfloat16)));
}

// ------------------------------------------------------------------ layers
NcsnModel::Norm NcsnModel::norm_coef(const T& x, const std::string& name) {
  double* sums = x.sums ? x.sums : static_cast<double*>(take((size_t)N_ * x.C * 2 * sizeof(double)));
  float2* coef = static_cast<float2*>(take((size_t)N_ * x.C * sizeof(float2)));
  const NcsnParam& ig = param(name + "/in_gamma");
  const NcsnParam& ib = param(name + "/in_beta");
  const float *gamma = nullptr, *alpha = nullptr, *beta = nullptr;
  int stride = 0;
  if (v1_) {
    const NcsnParam& e = param(name + "/embed");
    ASEP_CHECK(e.shape.size() == 2 && e.shape[1] == 3 * x.C, ASEP_ERR_BAD_SHAPE, "%s/embed: expected [classes, %d]",
               name.c_str(), 3 * x.C);
    gamma = e.dev; alpha = e.dev + x.C; beta = e.dev + 2 * x.C;      // Embedding row = [gamma | alpha | beta]
    stride = 3 * x.C;
  } else {
    const NcsnParam &g = param(name + "/gamma"), &a = param(name + "/alpha"), &b = param(name + "/beta");
    ASEP_CHECK((int)g.host.size() == x.C && (int)a.host.size() == x.C && (int)b.host.size() == x.C, ASEP_ERR_BAD_SHAPE,
               "%s: channel mismatch", name.c_str());
    gamma = g.dev; alpha = a.dev; beta = b.dev;
  }
  ASEP_CHECK((int)ig.host.size() == x.C && (int)ib.host.size() == x.C, ASEP_ERR_BAD_SHAPE, "%s: channel mismatch", name.c_str());
  Norm out;
  out.coef = coef; out.sums = sums; out.name = name;
  if (dry_) return out;
  if (!x.sums) launch_in_stats(x.p, sums, N_, x.H * x.W, x.C, s_);   // conv outputs arrive with their statistics
  launch_in_coef(sums, gamma, alpha, beta, stride, v1_ ? idx_ : nullptr, ig.dev, ib.dev, coef, N_, x.H * x.W, x.C, s_);
  return out;
}

NcsnModel::BF NcsnModel::prep(const T& x, const Norm& norm, bool elu, const T* stat_src) {
  BF y;
  const size_t n = (size_t)N_ * x.H * x.W * x.C;
  if (train_) y.gy = static_cast<float*>(take_g(n * sizeof(float)));
  Op op;
  op.kind = Op::kPrep; op.a = x; op.b = stat_src ? *stat_src : x; op.norm = norm; op.elu = elu;
  if (norm.coef == nullptr && !elu && x.bf != nullptr) {      // plain cast already done by the producing convolution
    y.hi = x.bf;
    op.bf = y;
    record(op);
    return y;
  }
  y.hi = new_bf(x.H, x.W, x.C);
  if (x3_) y.lo = new_bf(x.H, x.W, x.C);
  if (!dry_) launch_prep(x.p, norm.coef, y.hi, y.lo, N_, x.H * x.W, x.C, elu ? 1 : 0, s_);
  op.bf = y;
  record(op);
  return y;
}

NcsnModel::T NcsnModel::conv(const std::string& name, const BF& xin, int H, int W, const float* add, bool stats,
                             bool bf16_copy, const float* add2) {
  auto it = convs_.find(name);
  ASEP_CHECK(it != convs_.end(), ASEP_ERR_STATE, "convolution '%s' has no kernel parameter", name.c_str());
  const ConvWeightsTC& w = it->second;
  T out = new_t(H, W, w.Cout);
  if (stats) out.sums = static_cast<double*>(take((size_t)N_ * w.Cout * 2 * sizeof(double)));
  if (bf16_copy && !x3_) out.bf = new_bf(H, W, w.Cout);            // (the x3 mode needs the lo word too: k_prep writes both)
  if (dry_) return out;
  {
    Op op;
    op.kind = Op::kConv; op.out = out; op.bf = xin; op.name = name; op.add = add; op.add2 = add2;
    op.a.H = H; op.a.W = W;
    record(op);
  }
  if (!x3_) {
    conv_tc_forward(w, xin.hi, add, out.p, N_, H, W, s_, out.sums, out.bf, add2);
    return out;
  }
  // split-bf16: x.w = xhi.whi + xlo.whi + xhi.wlo (+ xlo.wlo ~ 2^-18, dropped), small terms first, fp32 accumulation in
  // the epilogues (`add` may alias `out`); bias and the statistics ride on the last pass
  auto lo = convs_lo_.find(name);
  ASEP_CHECK(lo != convs_lo_.end() && xin.lo != nullptr, ASEP_ERR_STATE, "'%s': split-bf16 operands missing", name.c_str());
  ConvWeightsTC w_nobias = w;
  w_nobias.bias = nullptr;
  conv_tc_forward(lo->second, xin.hi, add, out.p, N_, H, W, s_, nullptr, nullptr, add2);
  conv_tc_forward(w_nobias, xin.lo, out.p, out.p, N_, H, W, s_);
  conv_tc_forward(w, xin.hi, out.p, out.p, N_, H, W, s_, out.sums);
  return out;
}

// (Conditional)ResidualBlock: score_network.py:165-178 / score_network_v2.py:156-171
NcsnModel::T NcsnModel::res_block(const T& x, const std::string& name, int cout, bool down, int dilation) {
  (void)cout; (void)dilation;
  const Norm c1 = norm_coef(x, name + "/norm1");
  const BF h = prep(x, c1, true);
  T o1 = conv(name + "/conv1", h, x.H, x.W, nullptr, true);                   // norm2 follows
  const Norm c2 = norm_coef(o1, name + "/norm2");
  const BF h2 = prep(o1, c2, true);
  const float* sc = x.p;
  if (has(name + "/shortcut/kernel")) {
    const BF xr = prep(x, Norm{}, false);
    sc = conv(name + "/shortcut", xr, x.H, x.W, nullptr, false).p;
  }
  const bool pool = down && name.rfind("Res2_", 0) == 0;  // only the undilated 'down' block pools (score_network.py:141-144)
  // shortcut + output fused into the epilogue; an unpooled block output feeds norm layers (next block / RCU / MSF)
  T o2 = conv(name + "/conv2", h2, x.H, x.W, sc, !pool, !pool);
  if (!pool) return o2;
  // avg_pool2(shortcut) + avg_pool2(output) == avg_pool2(shortcut + output)
  T out = new_t(x.H / 2, x.W / 2, o2.C);
  if (!dry_) launch_avgpool2(o2.p, out.p, N_, out.H, out.W, out.C, s_);
  { Op op; op.kind = Op::kAvgPool2; op.a = o2; op.out = out; record(op); }
  return out;
}

// (Cond)RCUBlock: score_network.py:47-54 (norm -> conv, no activation: quirk Q8) / score_network_v2.py:41-47
NcsnModel::T NcsnModel::rcu(T x, const std::string& prefix, int n_blocks, int n_stages) {
  for (int i = 0; i < n_blocks; ++i) {
    const T residual = x;
    for (int j = 0; j < n_stages; ++j) {
      const std::string sfx = "_" + std::to_string(i + 1) + "_" + std::to_string(j + 1);
      const Norm c = v1_ ? norm_coef(x, prefix + "/norm" + sfx) : Norm{};
      const BF h = prep(x, c, false);
      x = conv(prefix + "/conv" + sfx, h, x.H, x.W, j == n_stages - 1 ? residual.p : nullptr, v1_, !v1_);
    }
  }
  return x;
}

// (Cond)CRPBlock: score_network.py:20-28 (norm -> 5x5 avg-pool -> conv) / score_network_v2.py:15-25 (5x5 max-pool -> conv)
NcsnModel::T NcsnModel::crp(T x, const std::string& prefix) {
  T acc = new_t(x.H, x.W, x.C);
  if (v1_) {
    // the first stage normalises the ELU output: its statistics are accumulated while it is written
    acc.sums = static_cast<double*>(take((size_t)N_ * x.C * 2 * sizeof(double)));
    if (!dry_) launch_elu_stats(x.p, acc.p, acc.sums, N_, x.H * x.W, x.C, s_);
  } else if (!dry_) {
    launch_elu(x.p, acc.p, (long long)N_ * x.H * x.W * x.C, s_);
  }
  { Op op; op.kind = Op::kElu; op.a = x; op.out = acc; record(op); }
  // x + path_1 + path_2 (score_network.py:24-27): path_1 is needed raw by the second stage's pooling, so the two
  // accumulations ride in the epilogue of the second convolution (out = conv_2 + acc + path_1) instead of two add passes
  T path = acc, path1;
  for (int i = 0; i < 2; ++i) {
    const std::string sfx = "_" + std::to_string(i + 1);
    // avg over the in-bounds taps commutes with the per-(n,c) affine of the norm: pool first, normalise while casting
    const Norm c = v1_ ? norm_coef(path, prefix + "/norm" + sfx) : Norm{};
    T pooled = new_t(x.H, x.W, x.C), ptmp = new_t(x.H, x.W, x.C);
    if (!dry_) launch_pool5(path.p, ptmp.p, pooled.p, N_, x.H, x.W, x.C, v1_ ? 0 : 1, s_);
    { Op op; op.kind = Op::kPool5; op.a = path; op.b = ptmp; op.out = pooled; record(op); }
    const BF h = prep(pooled, c, false, &path);
    if (i == 0) {
      path = conv(prefix + "/conv" + sfx, h, x.H, x.W, nullptr, v1_);
      path1 = path;
    } else {
      // (v1: the block output feeds the norm of RCU_output: its statistics ride in the same epilogue)
      path = conv(prefix + "/conv" + sfx, h, x.H, x.W, acc.p, v1_, false, path1.p);
    }
  }
  return path;
}

// (Cond)MSFBlock: score_network.py:70-79 / score_network_v2.py:61-69
NcsnModel::T NcsnModel::msf(const std::vector<T>& xs, const std::string& prefix, int H, int W, int features) {
  (void)features;
  T sums;
  // same-resolution branches first so their sum can ride in the convolution epilogue
  for (int pass = 0; pass < 2; ++pass)
    for (size_t i = 0; i < xs.size(); ++i) {
      const T& xi = xs[i];
      const bool same = xi.H == H && xi.W == W;
      if (same != (pass == 0)) continue;
      const std::string sfx = "_" + std::to_string(i + 1);
      const Norm c = v1_ ? norm_coef(xi, prefix + "/norm" + sfx) : Norm{};
      const BF h = prep(xi, c, false);
      if (same) {
        sums = conv(prefix + "/conv" + sfx, h, xi.H, xi.W, sums.p, false);
      } else {
        ASEP_CHECK(2 * xi.H == H && 2 * xi.W == W, ASEP_ERR_UNSUPPORTED, "MSF resize other than x2");
        T low = conv(prefix + "/conv" + sfx, h, xi.H, xi.W, nullptr, false);
        T up = new_t(H, W, low.C);
        if (!dry_) launch_resize2x_add(low.p, sums.p, up.p, N_, xi.H, xi.W, low.C, s_);
        { Op op; op.kind = Op::kResizeAdd; op.a = low; op.b = sums; op.out = up; record(op); }
        sums = up;
      }
    }
  return sums;
}

// (Cond)RefineBlock: score_network.py:103-118 / score_network_v2.py:93-107
NcsnModel::T NcsnModel::refine(const std::vector<T>& xs, const std::string& name, int features, bool end, int H, int W) {
  std::vector<T> hs;
  for (size_t i = 0; i < xs.size(); ++i) hs.push_back(rcu(xs[i], name + "/RCU_" + std::to_string(i + 1), 2, 2));
  T h = xs.size() > 1 ? msf(hs, name + "/MSF", H, W, features) : hs[0];
  h = crp(h, name + "/CRP");
  return rcu(h, name + "/RCU_output", end ? 3 : 1, 2);
}

void NcsnModel::run(const float* x, const int* idx, float* score) {
  const int H = cfg_.H, W = cfg_.W, ngf = cfg_.ngf;
  idx_ = idx;
  T out = new_t(H, W, ngf);
  const NcsnParam& bk = param("begin_conv/kernel");
  const NcsnParam& bb = param("begin_conv/bias");
  ASEP_CHECK((int64_t)bk.host.size() == 9 * ngf && (int)bb.host.size() == ngf, ASEP_ERR_BAD_SHAPE, "begin_conv shape");
  if (!dry_) launch_begin_conv(x, bk.dev, bb.dev, out.p, N_, H, W, ngf, v1_ ? 1 : 0, s_);   // 2x-1 only in v1 (Q10)
  { Op op; op.kind = Op::kBegin; op.out = out; record(op); }
  T l1 = res_block(res_block(out, "Res1_1", ngf, false, 0), "Res1_2", ngf, false, 0);
  T l2 = res_block(res_block(l1, "Res2_1", 2 * ngf, true, 0), "Res2_2", 2 * ngf, false, 0);
  T l3 = res_block(res_block(l2, "Res3_1", 2 * ngf, true, 2), "Res3_2", 2 * ngf, false, 2);
  T l4 = res_block(res_block(l3, "Res4_1", 2 * ngf, true, 4), "Res4_2", 2 * ngf, false, 4);
  T r1 = refine({l4}, "refine1", 2 * ngf, false, l4.H, l4.W);
  T r2 = refine({l3, r1}, "refine2", 2 * ngf, false, l3.H, l3.W);
  T r3 = refine({l2, r2}, "refine3", ngf, false, l2.H, l2.W);
  T o = refine({l1, r3}, "refine4", ngf, true, l1.H, l1.W);
  const Norm c = norm_coef(o, "normalizer");
  const BF h = prep(o, c, true);
  const NcsnParam& ek = param("end_conv/kernel");
  const NcsnParam& eb = param("end_conv/bias");
  ASEP_CHECK((int64_t)ek.host.size() == 9 * ngf && eb.host.size() == 1, ASEP_ERR_BAD_SHAPE, "end_conv shape");
  if (!dry_)
    launch_end_conv(h.hi, h.lo, ek.dev, eb.dev, v1_ ? nullptr : sigmas_dev_, v1_ ? nullptr : idx, score, N_, H, W, ngf, s_);
  { Op op; op.kind = Op::kEnd; op.bf = h; op.a = o; record(op); }
}

void NcsnModel::forward(const float* x, const int* idx, float* score, int N, cudaStream_t s) {
  if (N == 0) return;
  ASEP_CHECK(prepared_, ASEP_ERR_STATE, "asep_ncsn_prepare() must be called after setting parameters");
  CUDA_CHECK(cudaSetDevice(device_));
  s_ = s;
  N_ = N;
  if (training_ && images_dirty_) refresh_images(s);
  dry_ = true; arena_off_ = 0;
  run(nullptr, nullptr, nullptr);
  const size_t need = arena_off_;
  dry_ = false;
  if (need > arena_cap_) {
    CUDA_CHECK(cudaDeviceSynchronize());
    if (arena_) cudaFree(arena_);
    arena_ = nullptr;
    CUDA_CHECK(cudaMalloc(&arena_, need));
    arena_cap_ = need;
    ++generation_;
  }
  arena_off_ = 0;
  run(x, idx, score);
}

}  // namespace asep
