// Glow training step: loss, gradients of every trainable parameter, Adamax update (reference: train_glow.py:29-44,
// train_noisy_glow.py:30-33, train_utils.py:23-41).  The reverse sweep is grad_log_prob's; per flow step it adds the
// weight gradients (train_kernels.cu).  Data-parallel training sums `grads` over ranks with an NCCL all-reduce between
// train_grads() and adamax_step() (host side: audiosourcesep_b200/train_glow.py).
#include <cmath>
#include <cstring>

#include "glow_model.h"

namespace asep {

namespace {
bool trainable_name(const std::string& n) {
  auto ends = [&](const char* suf) {
    const size_t l = std::strlen(suf);
    return n.size() >= l && n.compare(n.size() - l, l, suf) == 0;
  };
  return !(ends("inv1x1/P") || ends("inv1x1/sign_S") || ends("moving_mean") || ends("moving_variance"));
}
}  // namespace

void GlowModel::enable_training() {
  if (training_) return;
  ASEP_CHECK(prepared_, ASEP_ERR_STATE, "call asep_glow_prepare() before asep_glow_enable_training()");
  ASEP_CHECK(!is_tcx(), ASEP_ERR_UNSUPPORTED,
             "the split-precision modes have no weight-gradient path: prepare with ASEP_PREC_BF16 / FP16 / FP32 to train");
  CUDA_CHECK(cudaSetDevice(device_));
  CUDA_CHECK(cudaDeviceSynchronize());
  long long n = 0;
  for (const auto& name : order_)
    if (trainable_name(name)) { params_.at(name).flat_off = n; n += params_.at(name).numel(); }
  n_trainable_ = n;
  CUDA_CHECK(cudaMalloc(&theta_, (size_t)n * sizeof(float)));
  CUDA_CHECK(cudaMalloc(&adam_m_, (size_t)n * sizeof(float)));
  CUDA_CHECK(cudaMalloc(&adam_u_, (size_t)n * sizeof(float)));
  CUDA_CHECK(cudaMemset(adam_m_, 0, (size_t)n * sizeof(float)));
  CUDA_CHECK(cudaMemset(adam_u_, 0, (size_t)n * sizeof(float)));
  for (const auto& name : order_) {
    Param& p = params_.at(name);
    if (p.flat_off < 0) continue;
    CUDA_CHECK(cudaMemcpy(theta_ + p.flat_off, p.dev, (size_t)p.numel() * sizeof(float), cudaMemcpyDeviceToDevice));
    cudaFree(p.dev);
    p.dev = theta_ + p.flat_off;
    p.in_flat = true;
  }
  const int F = cfg_.n_filters;
  {
    // one block for every per-step accumulator, so that a single memset clears them (doubles first: alignment)
    const size_t n_stats = (size_t)(2 * 64 + 64 * 64), n_q2 = (size_t)F * F, n_dc = (size_t)F, n_r3 = (size_t)9 * F * 64,
                 n_s3 = (size_t)9 * 64, n_d1 = (size_t)F * 256;
    tscratch_bytes_ = n_stats * sizeof(double) + (n_q2 + 2 * n_dc + n_r3 + n_s3 + n_d1) * sizeof(float);
    CUDA_CHECK(cudaMalloc(&tscratch_, tscratch_bytes_));
    tstats_ = reinterpret_cast<double*>(tscratch_);
    tq2_ = reinterpret_cast<float*>(tstats_ + n_stats);
    tdc2_ = tq2_ + n_q2;
    tdc1_ = tdc2_ + n_dc;
    tr3_ = tdc1_ + n_dc;
    ts3_ = tr3_ + n_r3;
    td1_ = ts3_ + n_s3;
  }
  CUDA_CHECK(cudaMalloc(&ldc_, steps_.size() * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&ld_total_, sizeof(double)));
  adam_t_ = 0;
  training_ = true;
  prepare(precision_);              // re-reads the parameter pointers (they moved into the flat vector)
  derive_on_device(nullptr);
  CUDA_CHECK(cudaDeviceSynchronize());
}

StepTrainPtrs GlowModel::step_ptrs(int b, int k) {
  const std::string pre = "b" + std::to_string(b) + "/s" + std::to_string(k) + "/";
  auto P = [&](const char* n) -> const Param& { return params_.at(pre + n); };
  StepDerived& sd = step(b, k);
  StepTrainPtrs sp{};
  sp.C = levels_[b].C; sp.F = cfg_.n_filters;
  sp.an_ls = P("actnorm/log_scale").dev; sp.an_shift = P("actnorm/shift").dev;
  sp.P = P("inv1x1/P").dev; sp.L = P("inv1x1/L").dev; sp.U = P("inv1x1/U").dev;
  sp.logS = P("inv1x1/log_S").dev; sp.signS = P("inv1x1/sign_S").dev;
  sp.k1 = P("nn/conv1/kernel").dev; sp.c1 = P("nn/conv1/bias").dev;
  sp.bn1_gamma = P("nn/bn1/gamma").dev; sp.bn1_beta = P("nn/bn1/beta").dev;
  sp.bn1_mean = P("nn/bn1/moving_mean").dev; sp.bn1_var = P("nn/bn1/moving_variance").dev;
  sp.k2 = P("nn/conv2/kernel").dev; sp.c2 = P("nn/conv2/bias").dev;
  sp.bn2_gamma = P("nn/bn2/gamma").dev; sp.bn2_beta = P("nn/bn2/beta").dev;
  sp.bn2_mean = P("nn/bn2/moving_mean").dev; sp.bn2_var = P("nn/bn2/moving_variance").dev;
  sp.k3 = P("nn/conv3/kernel").dev; sp.c3 = P("nn/conv3/bias").dev;
  sp.sc = sd.sc; sp.g1f = sd.g1; sp.b1f = sd.b1; sp.g2f = sd.g2; sp.b2f = sd.b2; sp.k2t = sd.k2t;
  sp.o_an_ls = P("actnorm/log_scale").flat_off; sp.o_an_shift = P("actnorm/shift").flat_off;
  sp.o_L = P("inv1x1/L").flat_off; sp.o_U = P("inv1x1/U").flat_off; sp.o_logS = P("inv1x1/log_S").flat_off;
  sp.o_k1 = P("nn/conv1/kernel").flat_off; sp.o_c1 = P("nn/conv1/bias").flat_off;
  sp.o_bn1_gamma = P("nn/bn1/gamma").flat_off; sp.o_bn1_beta = P("nn/bn1/beta").flat_off;
  sp.o_k2 = P("nn/conv2/kernel").flat_off; sp.o_c2 = P("nn/conv2/bias").flat_off;
  sp.o_bn2_gamma = P("nn/bn2/gamma").flat_off; sp.o_bn2_beta = P("nn/bn2/beta").flat_off;
  sp.o_k3 = P("nn/conv3/kernel").flat_off; sp.o_c3 = P("nn/conv3/bias").flat_off;
  return sp;
}

// Refresh of every derived constant after a parameter update: ONE launch per kernel for all L*K flow steps (the rows of
// the device-side table carry each step's pointers), instead of three launches per step.
void GlowModel::derive_on_device(cudaStream_t s) {
  const int n_steps = (int)steps_.size();
  if (refresh_dirty_) {                                  // pointers changed (prepare / enable_training): rebuild the table
    std::vector<StepRefresh> rows((size_t)n_steps);
    for (int b = 0; b < cfg_.L; ++b)
      for (int k = 0; k < cfg_.K; ++k) {
        StepRefresh& r = rows[(size_t)b * cfg_.K + k];
        r = StepRefresh{};
        r.sp = step_ptrs(b, k);
        ASEP_CHECK(r.sp.C * r.sp.C <= 256, ASEP_ERR_UNSUPPORTED, "derive: C > 16");
        r.HW = (double)levels_[b].H * levels_[b].W;
        if (is_tc()) {
          const NNWeightsTC& w = step(b, k).wtc;
          r.fwd_img = w.fwd.img; r.bwd_img = w.bwd.img;
          r.k1p_f = w.fwd.k1_panels; r.n3p_f = w.fwd.n3p; r.k1p_b = w.bwd.k1_panels; r.n3p_b = w.bwd.n3p;
          r.bias1 = w.bias1; r.bias2 = w.bias2; r.const3 = w.const3; r.c3 = w.c3;
          r.f16 = w.f16 ? 1 : 0;
        }
      }
    if (!refresh_table_) CUDA_CHECK(cudaMalloc(&refresh_table_, rows.size() * sizeof(StepRefresh)));
    CUDA_CHECK(cudaStreamSynchronize(s));                // an earlier refresh may still read the old rows
    CUDA_CHECK(cudaMemcpy(refresh_table_, rows.data(), rows.size() * sizeof(StepRefresh), cudaMemcpyHostToDevice));
    refresh_dirty_ = false;
  }
  launch_derive_all(refresh_table_, n_steps, ldc_, precision_ == ASEP_PREC_FP32 ? 1 : 0, s);
  if (is_tc()) launch_build_tc_all(refresh_table_, n_steps, s);      // tcgen05 operands: tile images + folded biases
  launch_sum_doubles(ldc_, n_steps, ld_total_, s);
}

void GlowModel::ensure_train_dumps(long long rows) {
  if (rows <= dump_rows_) return;
  CUDA_CHECK(cudaDeviceSynchronize());
  invalidate_graphs();
  for (__nv_bfloat16** p : {&da1_, &da2_, &dgp2_, &dgp1_, &dcol_})
    if (*p) { cudaFree(*p); *p = nullptr; }
  const size_t n = (size_t)rows * cfg_.n_filters;
  for (__nv_bfloat16** p : {&da1_, &da2_, &dgp2_, &dgp1_}) CUDA_CHECK(cudaMalloc(p, n * sizeof(__nv_bfloat16)));
  CUDA_CHECK(cudaMalloc(&dcol_, (size_t)rows * 256 * sizeof(__nv_bfloat16)));
  dump_rows_ = rows;
}

void GlowModel::carve_scratch(char* base, TrainSlot& t) const {
  const int F = cfg_.n_filters;
  const size_t n_stats = (size_t)(2 * 64 + 64 * 64), n_q2 = (size_t)F * F, n_dc = (size_t)F, n_r3 = (size_t)9 * F * 64, n_s3 = (size_t)9 * 64;
  t.scratch = base;
  t.stats = reinterpret_cast<double*>(base);
  t.q2 = reinterpret_cast<float*>(t.stats + n_stats);
  t.dc2 = t.q2 + n_q2;
  t.dc1 = t.dc2 + n_dc;
  t.r3 = t.dc1 + n_dc;
  t.s3 = t.r3 + n_r3;
  t.d1 = t.s3 + n_s3;
}

void GlowModel::ensure_train_slots(long long rows) {
  if (rows <= slot_rows_) return;
  CUDA_CHECK(cudaDeviceSynchronize());
  invalidate_graphs();
  const int Cmax = levels_.back().C;
  for (TrainSlot& t : tslots_) {
    for (void* p : {(void*)t.gp2, (void*)t.gp1, (void*)t.col, (void*)t.gr, (void*)t.gu, (void*)t.gxb, (void*)t.scratch})
      if (p) cudaFree(p);
    const size_t n = (size_t)rows * cfg_.n_filters;
    CUDA_CHECK(cudaMalloc(&t.gp2, n * sizeof(__nv_bfloat16)));
    CUDA_CHECK(cudaMalloc(&t.gp1, n * sizeof(__nv_bfloat16)));
    CUDA_CHECK(cudaMalloc(&t.col, (size_t)rows * 256 * sizeof(__nv_bfloat16)));
    // [pixels of a level, C of that level] never exceeds rows * C_level0 floats (pixels shrink 4x per level, C grows 2x)
    const size_t st = (size_t)rows * std::max(levels_[0].C, Cmax / 4 + 1);
    CUDA_CHECK(cudaMalloc(&t.gr, st * sizeof(float)));
    CUDA_CHECK(cudaMalloc(&t.gu, st * sizeof(float)));
    CUDA_CHECK(cudaMalloc(&t.gxb, st * sizeof(float)));
    char* sc = nullptr;
    CUDA_CHECK(cudaMalloc(&sc, tscratch_bytes_));
    carve_scratch(sc, t);
    if (!t.ev_fork) {
      CUDA_CHECK(cudaEventCreateWithFlags(&t.ev_fork, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&t.ev_a, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&t.ev_join, cudaEventDisableTiming));
    }
    t.busy = false;
  }
  for (cudaStream_t& q : tside_)
    if (!q) CUDA_CHECK(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
  slot_rows_ = rows;
}

void GlowModel::train_grads(const float* x, const float* noise, float sigma, int N, int global_batch, float* grads,
                            float* loss, cudaStream_t s) {
  ASEP_CHECK(training_, ASEP_ERR_STATE, "asep_glow_enable_training() has not been called");
  ASEP_CHECK(N >= 1 && global_batch >= N, ASEP_ERR_BAD_ARG, "bad batch sizes (local %d, global %d)", N, global_batch);
  CUDA_CHECK(cudaSetDevice(device_));
  const bool use_graph = is_tc() && !nn_tc_profile_enabled() && getenv("ASEP_NO_GRAPH") == nullptr &&
                         getenv("ASEP_TC_DBG_TIMING") == nullptr;
  if (!use_graph || N > tg_calls_) {             // the first call of a batch size runs eagerly: it allocates the scratch
    train_grads_body(x, noise, sigma, N, global_batch, grads, loss, s);
    if (use_graph) tg_calls_ = N;                 // (largest batch whose scratch exists)
    return;
  }
  const size_t nx = (size_t)N * cfg_.H * cfg_.W * cfg_.C;
  if (tg_stream_ == nullptr) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&tg_stream_, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&tg_ev_in_, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&tg_ev_out_, cudaEventDisableTiming));
    CUDA_CHECK(cudaMalloc(&tg_grads_, (size_t)n_trainable_ * sizeof(float)));
    CUDA_CHECK(cudaMalloc(&tg_loss_, sizeof(float)));
  }
  if (nx > tg_x_cap_) {
    CUDA_CHECK(cudaDeviceSynchronize());
    if (tg_x_) cudaFree(tg_x_);
    if (tg_noise_) cudaFree(tg_noise_);
    CUDA_CHECK(cudaMalloc(&tg_x_, nx * sizeof(float)));
    CUDA_CHECK(cudaMalloc(&tg_noise_, nx * sizeof(float)));
    tg_x_cap_ = nx;
    if (tgraph_.exec) { cudaGraphExecDestroy(tgraph_.exec); tgraph_.exec = nullptr; }     // staging pointers changed
  }
  // inputs -> staging (on the caller's stream), then the private stream takes over
  CUDA_CHECK(cudaMemcpyAsync(tg_x_, x, nx * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (noise) CUDA_CHECK(cudaMemcpyAsync(tg_noise_, noise, nx * sizeof(float), cudaMemcpyDeviceToDevice, s));
  CUDA_CHECK(cudaEventRecord(tg_ev_in_, s));
  CUDA_CHECK(cudaStreamWaitEvent(tg_stream_, tg_ev_in_, 0));
  TrainGraph& g = tgraph_;
  if (g.exec == nullptr || g.N != N || g.global_batch != global_batch || g.sigma != sigma || g.noisy != (noise != nullptr)) {
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    cudaGraph_t graph = nullptr;
    const long long c0 = g_launch_count.load();
    CUDA_CHECK(cudaStreamBeginCapture(tg_stream_, cudaStreamCaptureModeRelaxed));
    try {
      train_grads_body(tg_x_, noise ? tg_noise_ : nullptr, sigma, N, global_batch, tg_grads_, tg_loss_, tg_stream_);
    } catch (...) {
      cudaStreamEndCapture(tg_stream_, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    CUDA_CHECK(cudaStreamEndCapture(tg_stream_, &graph));
    g.launches = g_launch_count.load() - c0;
    CUDA_CHECK(cudaGraphInstantiate(&g.exec, graph, 0));
    CUDA_CHECK(cudaGraphDestroy(graph));
    g.N = N; g.global_batch = global_batch; g.sigma = sigma; g.noisy = noise != nullptr;
  } else {
    g_launch_count.fetch_add(g.launches);      // the replay launches the same kernels
  }
  CUDA_CHECK(cudaGraphLaunch(g.exec, tg_stream_));
  CUDA_CHECK(cudaMemcpyAsync(grads, tg_grads_, (size_t)n_trainable_ * sizeof(float), cudaMemcpyDeviceToDevice, tg_stream_));
  CUDA_CHECK(cudaMemcpyAsync(loss, tg_loss_, sizeof(float), cudaMemcpyDeviceToDevice, tg_stream_));
  CUDA_CHECK(cudaEventRecord(tg_ev_out_, tg_stream_));
  CUDA_CHECK(cudaStreamWaitEvent(s, tg_ev_out_, 0));
}

void GlowModel::train_grads_body(const float* x, const float* noise, float sigma, int N, int global_batch, float* grads,
                                 float* loss, cudaStream_t s) {
  const int L = cfg_.L, K = cfg_.K, F = cfg_.n_filters;
  const float gs = -1.0f / (float)global_batch;      // loss = -sum log_prob / global_batch (train_glow.py:30-31)
  const bool tc = is_tc();
  ensure_work(N, true, tc);                 // tcgen05: the forward pass keeps bf16 copies of relu(p1), relu(p2) of every step
  if (tc) {
    ensure_train_dumps((long long)N * levels_[0].H * levels_[0].W);
    ensure_train_slots((long long)N * levels_[0].H * levels_[0].W);
    for (TrainSlot& t : tslots_) t.busy = false;
  }
  int step_no = 0;
  const float* xin = x;
  if (noise != nullptr) {                              // train_noisy_glow.py:31-32: X + sigma*N(0,1) in raw data units
    launch_axpy(x, noise, sigma, work_.gB, (long long)N * cfg_.H * cfg_.W * cfg_.C, s);
    xin = work_.gB;
  }
  run_forward(xin, N, true, s, tc);
  // learntop = False: standard-normal prior without trainable parameters (flow_builder.py:143-144)
  const float* loc = cfg_.learntop ? params_.at("prior/loc").dev : nullptr;
  const float* ls = cfg_.learntop ? params_.at("prior/log_scale").dev : nullptr;
  launch_prior(work_.z, loc, ls, work_.acc_prior, work_.gz, N, Dl_, s);
  // SpecPreprocessing log-det constant (flow_tfp_bijectors.py:390-396); the per-step constants come from the device
  const double pre_const = (double)cfg_.H * cfg_.W * cfg_.C * std::log(1.0 / ((double)cfg_.maxval - (double)cfg_.minval));
  launch_loss(work_.acc_ld, work_.acc_prior, ld_total_, pre_const, N, 1.0 / (double)global_batch, loss, s);
  CUDA_CHECK(cudaMemsetAsync(grads, 0, (size_t)n_trainable_ * sizeof(float), s));
  if (cfg_.learntop)
    launch_prior_grads(work_.z, loc, ls, grads + params_.at("prior/loc").flat_off, grads + params_.at("prior/log_scale").flat_off,
                       N, Dl_, gs, s);
  float* gX_next = nullptr;
  for (int b = L - 1; b >= 0; --b) {
    const Level& lv = levels_[b];
    const long long M = (long long)N * lv.H * lv.W;
    int Cz, nb, coff;
    latent_slice(b, Cz, nb, coff);
    float* gy = work_.gA;
    float* other = work_.gB;
    if (gX_next == gy) std::swap(gy, other);
    launch_split_merge(gy, work_.gz, gX_next, N, lv.H, lv.W, lv.C, Cz, nb, CL_, coff, Dl_, 1, s);
    for (int k = 0; k < K; ++k) {
      StepDerived& sd = step(b, k);
      const StepTrainPtrs sp = step_ptrs(b, k);
      // tcgen05 path: this step's buffers come from the ring (its side work may still be running when the next step starts)
      TrainSlot* slot = tc ? &tslots_[step_no % kTrainRing] : nullptr;
      ++step_no;
      if (slot && slot->busy) CUDA_CHECK(cudaStreamWaitEvent(s, slot->ev_join, 0));
      float* gr_k = slot ? slot->gr : work_.gr;
      float* gu_k = slot ? slot->gu : work_.gu;
      float* gxb_k = slot ? slot->gxb : work_.gxb;
      launch_bwd_coupling(gy, work_.U[b][k], work_.R[b][k], gr_k, gu_k, M, lv.C, s);
      // every per-step accumulator (Q2, dc2, R3, S3, D1, dc1, step statistics) lives in ONE block: one memset per step
      CUDA_CHECK(cudaMemsetAsync(slot ? slot->scratch : tscratch_, 0, tscratch_bytes_, s));
      if (!tc) {
        // recompute a1, a2 from the saved step input, then gp2 (t2), gp1 (t1), gxb
        nn_fp32_forward(sd.w32, work_.U[b][k], work_.a1, work_.a2, work_.gxb, N, lv.H, lv.W, lv.C, F, s);
        nn_fp32_backward(sd.w32, work_.a1, work_.a2, work_.gr, work_.t1, work_.t2, work_.gxb, N, lv.H, lv.W, lv.C, F, s);
        // ---- weight gradients of this step (CUDA-core fp32)
        launch_wgrad_tn(work_.a1, work_.t2, tq2_, M, F, s);
        launch_colsum(work_.t2, tdc2_, M, F, s);
        launch_wgrad_conv3(work_.a2, work_.gr, tr3_, ts3_, N, lv.H, lv.W, lv.C, F, s);
        launch_wgrad_conv1(work_.U[b][k], work_.t1, grads + sp.o_k1, grads + sp.o_c1, N, lv.H, lv.W, lv.C, F, gs, s);
        launch_step_stats(work_.gu, work_.gxb, work_.U[b][k], sd.sc, tstats_, M, lv.C, s);
        launch_finalize_step(sp, tq2_, tdc2_, tr3_, ts3_, tstats_, grads, (double)M, gs, s);
      } else {
        // tcgen05: recompute the forward with bf16 dumps of relu(p1), relu(p2); the backward dumps dL/dp2, dL/dp1
        const bool saved = !work_.M1.empty() && !work_.M1[b].empty();
        uint32_t* m1 = saved ? work_.M1[b][k] : work_.tc.mask1;
        uint32_t* m2 = saved ? work_.M2[b][k] : work_.tc.mask2;
        // relu(p1), relu(p2) of this step: kept by the forward pass (work_.dumps) or recomputed from the saved step input
        const bool kept = saved && work_.dumps && !work_.D1[b].empty();      // (written by this call's run_forward)
        const __nv_bfloat16* a1 = kept ? work_.D1[b][k] : da1_;
        const __nv_bfloat16* a2 = kept ? work_.D2[b][k] : da2_;
        if (!kept)
          nn_tc_forward(sd.wtc, work_.tc, work_.U[b][k], work_.gxb, saved ? nullptr : m1, saved ? nullptr : m2, N, lv.H,
                        lv.W, lv.C, s, da1_, da2_);
        // main stream: data gradient of the network; the gradients of this step's weights are left to the side streams
        nn_tc_backward(sd.wtc, work_.tc, slot->gr, m1, m2, slot->gxb, N, lv.H, lv.W, lv.C, s, slot->gp2, slot->gp1);
        CUDA_CHECK(cudaEventRecord(slot->ev_fork, s));
        const int ld3 = (9 * lv.C + 63) / 64 * 64, ld1 = (9 * (lv.C / 2) + 63) / 64 * 64;
        cudaStream_t qa = tside_[0], qb = tside_[1];
        // side A: Q2 = a1^T gp2 and the two bias gradients
        CUDA_CHECK(cudaStreamWaitEvent(qa, slot->ev_fork, 0));
        wgrad_tc(a1, slot->gp2, F, F, slot->q2, F, M, qa);
        launch_colsum_bf16(slot->gp2, slot->dc2, M, F, qa);
        launch_colsum_bf16(slot->gp1, slot->dc1, M, F, qa);
        CUDA_CHECK(cudaEventRecord(slot->ev_a, qa));
        // side B: conv3 / conv1 weight gradients, the element-wise statistics, then the chain rule into `grads`
        CUDA_CHECK(cudaStreamWaitEvent(qb, slot->ev_fork, 0));
        launch_im2col_gr(slot->gr, slot->col, N, lv.H, lv.W, lv.C, ld3, qb);
        wgrad_tc(a2, slot->col, ld3, 9 * lv.C, slot->r3, ld3, M, qb);                  // R3t[k][tap,c]
        launch_s3(slot->gr, slot->s3, N, lv.H, lv.W, lv.C, qb);
        launch_im2col_xb(work_.U[b][k], slot->col, N, lv.H, lv.W, lv.C, ld1, qb);
        wgrad_tc(slot->gp1, slot->col, ld1, 9 * (lv.C / 2), slot->d1, ld1, M, qb);     // D1t[f][tap,ci]
        launch_step_stats(slot->gu, slot->gxb, work_.U[b][k], sd.sc, slot->stats, M, lv.C, qb);
        CUDA_CHECK(cudaStreamWaitEvent(qb, slot->ev_a, 0));
        launch_finalize_step_tc(sp, slot->q2, slot->dc2, slot->r3, ld3, slot->s3, slot->d1, ld1, slot->dc1, slot->stats, grads,
                                (double)M, gs, qb);
        CUDA_CHECK(cudaEventRecord(slot->ev_join, qb));
        slot->busy = true;
        if (!kept) {            // recomputed activations live in ONE scratch pair: the next step may not overwrite them early
          CUDA_CHECK(cudaStreamWaitEvent(s, slot->ev_join, 0));
          slot->busy = false;
        }
      }
      // ---- data gradient to the previous step
      launch_bwd_pre(gu_k, gxb_k, other, sd.sc, M, lv.C, s);
      std::swap(gy, other);
    }
    gX_next = gy;
  }
  // join: the gradient vector is complete once every side stream has drained (also required to end a stream capture)
  if (tc)
    for (TrainSlot& t : tslots_)
      if (t.busy) { CUDA_CHECK(cudaStreamWaitEvent(s, t.ev_join, 0)); t.busy = false; }
}

void GlowModel::adamax_step(const float* grads, float lr, float beta1, float beta2, float eps, cudaStream_t s) {
  ASEP_CHECK(training_, ASEP_ERR_STATE, "asep_glow_enable_training() has not been called");
  CUDA_CHECK(cudaSetDevice(device_));
  ++adam_t_;
  const float lr_t = (float)((double)lr / (1.0 - std::pow((double)beta1, (double)adam_t_)));
  launch_adamax(theta_, grads, adam_m_, adam_u_, n_trainable_, lr_t, beta1, beta2, eps, s);
  derive_on_device(s);
}

void GlowModel::adam_step(const float* grads, float lr, float beta1, float beta2, float eps, cudaStream_t s) {
  ASEP_CHECK(training_, ASEP_ERR_STATE, "asep_glow_enable_training() has not been called");
  CUDA_CHECK(cudaSetDevice(device_));
  ++adam_t_;
  const double t = (double)adam_t_;
  const float lr_t = (float)((double)lr * std::sqrt(1.0 - std::pow((double)beta2, t)) / (1.0 - std::pow((double)beta1, t)));
  launch_adam(theta_, grads, adam_m_, adam_u_, n_trainable_, lr_t, beta1, beta2, eps, s);
  derive_on_device(s);
}

void GlowModel::copy_flat(float* dst, cudaStream_t s) const {
  ASEP_CHECK(training_, ASEP_ERR_STATE, "asep_glow_enable_training() has not been called");
  CUDA_CHECK(cudaMemcpyAsync(dst, theta_, (size_t)n_trainable_ * sizeof(float), cudaMemcpyDeviceToDevice, s));
}

void GlowModel::set_flat(const float* src, cudaStream_t s) {
  ASEP_CHECK(training_, ASEP_ERR_STATE, "asep_glow_enable_training() has not been called");
  CUDA_CHECK(cudaMemcpyAsync(theta_, src, (size_t)n_trainable_ * sizeof(float), cudaMemcpyDeviceToDevice, s));
  derive_on_device(s);
}

void GlowModel::sync_host() {
  if (!training_) return;
  CUDA_CHECK(cudaDeviceSynchronize());
  for (const auto& name : order_) {
    Param& p = params_.at(name);
    if (p.flat_off < 0) continue;
    CUDA_CHECK(cudaMemcpy(p.host.data(), p.dev, (size_t)p.numel() * sizeof(float), cudaMemcpyDeviceToHost));
  }
  prepare(precision_);
}

}  // namespace asep
