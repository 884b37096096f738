#!/usr/bin/env python
"""Benchmark of the separation hot path (driver contract: one JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

Headline metric (BASELINE.json): Glow ``log_prob`` samples/s on synthetic mel-spectrogram patches of
the ``configs/melspec_glow.yml`` shape (96x64x1, L=3, K=40, 512 filters, learnable top prior),
random-init ("perturbed") weights.  One *step* = one ``log_prob`` pass over one batch of B patches per
GPU.  The second half of the metric, BASIS Langevin segment-steps/s with two Glow priors, is measured in
the same run and reported under ``"basis"``.

* ``value``  : whole-job samples/s, inputs resident in HBM, device time by CUDA events (per step, with
               an L2 flush between steps outside the events), max over ranks.
* ``e2e``    : the same metric through the public API (``Glow.log_prob`` -> C ABI) with the batch in
               PINNED HOST memory: H2D copy + compute + D2H of the [B] log-probabilities inside the
               timed region, every step.
* ``roofline``: the dominant kernel (k_nn_tc4, the fused tcgen05 coupling network): algorithmic conv
               FLOPs of the recorded launches / their summed CUDA-event time (events on the launching
               stream, recorded during the timed region) vs the measured sustained bf16 peak.
* ``cpu_baseline`` / ``--impl reference``: the reference-faithful CPU graph (oracle/glow_oracle.py,
               torch-CPU fp32, coupling network evaluated twice as the TFP graph does) on the host
               cores, on a bounded sample of the same workload.  TensorFlow 2.2 cannot be installed in
               this image, so kind = "port".

Multi-GPU (torchrun, one rank per GPU): the batch dimension shards with no data-path collective
("weak": B patches per GPU); NCCL is used only for the barrier and the max-over-ranks timing.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

F_GLOW = 48.22e9          # FLOP per patch per coupling-network pass
D_PATCH = 96 * 64


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_sustained": float(p.get("bf16_tflops_sustained", 1378.1)), "bf16_burst": float(p.get("bf16_tflops", 1643.9)),
                "hbm": float(p.get("hbm_gbs", 6551.0)), "source": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi SM clock / throttle reasons sampled every 200 ms while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.strip().split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU reference arm
def _cpu_model(K: int):
    import torch
    from audiosourcesep_b200 import GlowConfig
    from audiosourcesep_b200.weights import init_glow_params
    from oracle.glow_oracle import GlowOracle
    cfg = GlowConfig(K=K)
    return cfg, GlowOracle(cfg, init_glow_params(cfg, seed=2), dtype=torch.float32)


def cpu_log_prob_rate(sample: int, K: int = 40, repeats: int = 1):
    """samples/s of the reference-faithful CPU graph on `sample` patches (all host threads)."""
    import torch
    from audiosourcesep_b200 import synthetic
    _, oracle = _cpu_model(K)
    x = torch.as_tensor(synthetic.mel_patches_db(sample, seed=0))
    best = None
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            oracle.log_prob_reference_graph(x)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return sample / best, best


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path.  The TF graph cannot run
    here (TensorFlow 2.2 / TFP 0.9 absent, no wheel, Python 3.12), so the reference-faithful port in
    oracle/ is timed with every host thread torch-CPU will use; rank 0 only."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 for its workers; the reference arm runs on rank 0 alone and takes every core
    try:
        torch.set_num_threads(max(torch.get_num_threads(), os.cpu_count() or 1))
    except RuntimeError:
        pass
    cores = torch.get_num_threads()
    sample = args.cpu_sample
    _, oracle = _cpu_model(args.K)
    from audiosourcesep_b200 import synthetic
    x = torch.as_tensor(synthetic.mel_patches_db(sample, seed=0))
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            oracle.log_prob_reference_graph(x)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms * 1e-3)
    desc = (f"{sample} patches per step, reference-faithful graph (coupling network evaluated twice), torch-CPU fp32, "
            f"{cores} threads of {os.cpu_count()} logical cores")
    line = {
        "impl": "reference", "metric": "glow_log_prob_samples_per_s", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(args, per_gpu_batch=sample, note="bounded CPU sample of the same workload"),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _config(args, per_gpu_batch, note=None):
    c = {"workload": f"glow_log_prob melspec_glow.yml (96x64x1, L=3, K={args.K}, n_filters=512, learntop), "
                     f"{per_gpu_batch} patches per GPU per step",
         "per_gpu_batch": per_gpu_batch, "global_batch": per_gpu_batch * args.gpus,
         "weights": "random-init (perturbed generator, seed 2)",
         "cache": "L2 flushed (512 MiB write) between timed steps, outside the CUDA-event brackets",
         "sharding": "patches sharded over ranks, no data-path collective"}
    if note:
        c["note"] = note
    return c


# ------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: audiosourcesep_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from audiosourcesep_b200 import GlowConfig, _lib, ops, synthetic
    from audiosourcesep_b200.glow import Glow
    from audiosourcesep_b200.weights import init_glow_params
    # (the product path never touches oracle/: the sigma schedule and step constants come from the package)
    from audiosourcesep_b200.ncsn import utils as bo

    dev = torch.device("cuda", local_rank)
    peaks = _peaks()
    cfg = GlowConfig(K=args.K)
    params = init_glow_params(cfg, seed=2)
    model = Glow(cfg, params, precision=_lib.PREC_BF16, device=local_rank)
    ops.set_tc_cluster(args.cluster)
    B = args.batch
    base = synthetic.mel_patches_db(min(B, 64), seed=100 + rank)
    reps = (B + base.shape[0] - 1) // base.shape[0]
    # distinct patches: cyclic shifts of the seeded base set along time
    host = np.concatenate([np.roll(base, 3 * i, axis=2) for i in range(reps)], axis=0)[:B]
    x_host = torch.as_tensor(np.ascontiguousarray(host)).pin_memory()
    x_dev = x_host.to(dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(step_fn, steps, warmup, profile=False):
        for _ in range(warmup):
            step_fn()
        barrier()
        evs = []
        if profile == "conv":
            _lib.conv_profile(True)
        elif profile:
            _lib.tc_profile(True)
        n0 = _lib.launch_count()
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn()
            b.record()
            evs.append((a, b))
        barrier()
        launches = _lib.launch_count() - n0
        prof = None
        if profile == "conv":
            prof = _lib.conv_profile_read()
            _lib.conv_profile(False)
        elif profile:
            prof = _lib.tc_profile_read()
            _lib.tc_profile(False)
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, prof

    # ---- 1. device-resident throughput (value) with the kernel-level CUDA events for the roofline
    out = {}

    def step_dev():
        out["lp"] = model.log_prob(x_dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, launches, prof = timed_loop(step_dev, args.steps, args.warmup, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms * 1e-3)
    lp = out["lp"].float().cpu().numpy()
    if not np.all(np.isfinite(lp)):
        raise SystemExit("non-finite log_prob in the benchmark batch")

    # ---- 2. end to end through the public API with host buffers
    def step_e2e():
        out["lp_host"] = model.log_prob(x_host.to(dev, non_blocking=True)).cpu()

    e2e_ms, _, _ = timed_loop(step_e2e, args.steps, args.warmup)
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)

    # ---- 2b. the other directions of config 1 (forward+inverse): inverse(z) and grad_log_prob(x) samples/s, device-resident
    z_dev = model.forward(x_dev)

    def step_inv():
        out["xr"] = model.inverse(z_dev)

    inv_ms, inv_launches, _ = timed_loop(step_inv, max(2, args.steps // 2), 2)
    rt = float((out["xr"] - x_dev).abs().max().item())

    def step_grad():
        out["g"] = model.grad_log_prob(x_dev)

    g_ms, g_launches, _ = timed_loop(step_grad, max(2, args.steps // 2), 2)
    nsub = max(2, args.steps // 2)
    directions = {
        "inverse": {"value": world * B * nsub / (inv_ms * 1e-3), "unit": "samples/s", "ms_per_step": inv_ms / nsub,
                    "gpu_launches": inv_launches, "alg_tflops": world * B * nsub * F_GLOW * (args.K / 40.0) / (inv_ms * 1e-3) / 1e12,
                    "round_trip_max_abs_dB": rt,
                    "note": "bf16 tensor-core mode; the <= 1e-4 (normalised units) round-trip gate is met by ASEP_PREC_FP32, see DESIGN.md section 4"},
        "grad_log_prob": {"value": world * B * nsub / (g_ms * 1e-3), "unit": "samples/s", "ms_per_step": g_ms / nsub,
                          "gpu_launches": g_launches,
                          "alg_tflops": world * B * nsub * 2 * F_GLOW * (args.K / 40.0) / (g_ms * 1e-3) / 1e12},
    }
    del z_dev

    # ---- 3. BASIS Langevin segment-steps/s with two Glow priors (second half of the metric)
    basis = None
    if args.basis_segments > 0:
        nseg = args.basis_segments
        bcfg = GlowConfig(K=args.K, minval=0.0, maxval=1.0)
        m1 = Glow(bcfg, init_glow_params(bcfg, seed=2), precision=_lib.PREC_BF16, device=local_rank)
        m2 = Glow(bcfg, init_glow_params(bcfg, seed=3), precision=_lib.PREC_BF16, device=local_rank)
        mixed, _, _ = synthetic.basis_problem(min(nseg, 32), seed1=10 + rank, seed2=50 + rank)
        mixed = np.concatenate([mixed] * ((nseg + mixed.shape[0] - 1) // mixed.shape[0]))[:nseg]
        x1, x2 = synthetic.langevin_init(nseg, seed=4 + rank)
        mixed_d = torch.as_tensor(mixed).to(dev)
        t1, t2 = torch.as_tensor(x1).to(dev), torch.as_tensor(x2).to(dev)
        sig = bo.get_sigmas(1.0, 0.01, 10, "logarithmic")
        eta, lam, ns = bo.langevin_step_constants(sig, 9)
        T = args.basis_T
        stepno = [0]

        def step_basis():
            ops.basis_glow_inner(m1, m2, mixed_d, t1, t2, T, float(eta), float(lam), float(ns), seed=1,
                                 step0=stepno[0], elem_offset=rank * nseg * D_PATCH)
            stepno[0] += T

        b_ms, b_launches, _ = timed_loop(step_basis, max(1, args.steps // 2), 1)
        nsteps = max(1, args.steps // 2)
        basis = {"metric": "basis_glow_segment_steps_per_s", "value": world * nseg * T * nsteps / (b_ms * 1e-3),
                 "unit": "segment-steps/s", "segments_per_gpu": nseg, "langevin_steps_per_call": T,
                 "ms_per_langevin_step": b_ms / (nsteps * T), "gpu_launches": b_launches,
                 "alg_tflops": world * nseg * T * nsteps * 4 * F_GLOW * (args.K / 40.0) / (b_ms * 1e-3) / 1e12,
                 "note": "2 priors x (forward + data-gradient) = 4 F_glow algorithmic FLOP per segment-step; in-kernel "
                         "Philox noise; sigma index 9 of the 10-level schedule"}
        if not (torch.isfinite(t1).all() and torch.isfinite(t2).all()):
            basis["note"] += "; WARNING non-finite state"

    # ---- 4. BASIS with NCSN v1 / v2 score networks (configs 4 and 5 of BASELINE.json), n_mixed = 30 segments per GPU
    ncsn = {}
    if args.ncsn_segments > 0:
        from audiosourcesep_b200 import NCSNConfig
        from audiosourcesep_b200.ncsn.score_model import ScoreModel
        from audiosourcesep_b200.weights import init_ncsn_params
        nseg = args.ncsn_segments
        mixed, _, _ = synthetic.basis_problem(nseg, seed1=20 + rank, seed2=60 + rank)
        mixed_d = torch.as_tensor(mixed).to(dev)
        for ver, ncfg, gflop in (("v1", NCSNConfig(version="v1", ngf=192, num_classes=10, sigma1=1.0), 533.9),
                                 ("v2", NCSNConfig(version="v2", ngf=128, num_classes=200, sigma1=30.0), 237.3)):
            sig_n = bo.get_sigmas(ncfg.sigma1, ncfg.sigmaL, ncfg.num_classes, "logarithmic")
            s1 = ScoreModel(ncfg, init_ncsn_params(ncfg, seed=11), sigmas=sig_n, device=local_rank)
            s2 = ScoreModel(ncfg, init_ncsn_params(ncfg, seed=12), sigmas=sig_n, device=local_rank)
            a1, a2 = synthetic.langevin_init(nseg, seed=7 + rank)
            u1, u2 = torch.as_tensor(a1).to(dev), torch.as_tensor(a2).to(dev)
            idx = ncfg.num_classes - 1
            eta_n, lam_n, ns_n = bo.langevin_step_constants(sig_n, idx)
            cnt = [0]

            def step_ncsn():
                ops.basis_ncsn_inner(s1, s2, mixed_d, u1, u2, idx, args.ncsn_T, float(eta_n), float(lam_n), float(ns_n),
                                     seed=2, step0=cnt[0], elem_offset=rank * nseg * D_PATCH)
                cnt[0] += args.ncsn_T

            nst = max(2, args.steps // 2)
            n_ms, n_launches, (c_ms, c_n, c_fl) = timed_loop(step_ncsn, nst, 2, profile="conv")
            rate = world * nseg * args.ncsn_T * nst / (n_ms * 1e-3)
            conv_tf = c_fl / (c_ms * 1e-3) / 1e12 if c_ms > 0 else 0.0
            ncsn[ver] = {"metric": f"basis_ncsn_{ver}_segment_steps_per_s", "value": rate, "unit": "segment-steps/s",
                         "segments_per_gpu": nseg, "ms_per_langevin_step": n_ms / (nst * args.ncsn_T),
                         "gpu_launches": n_launches, "alg_tflops": rate * gflop / 1e3,
                         "roofline": {"bound": "tensor", "kernel": "k_conv_tc (TMA-fed tcgen05 implicit-GEMM convolution)",
                                      "achieved": conv_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                                      "frac": conv_tf / peaks["bf16_sustained"], "launches": c_n,
                                      "kernel_share_of_step": c_ms / n_ms if n_ms > 0 else None},
                         "finite": bool(torch.isfinite(u1).all() and torch.isfinite(u2).all())}
            del s1, s2

    # ---- 5. Glow training step (config 2): fp32 exact mode, global batch 32 per GPU, Adamax, NCCL gradient all-reduce
    train = None
    if args.train_batch > 0:
        from audiosourcesep_b200 import train_glow as tg
        tcfg = GlowConfig(K=args.K)
        tb = args.train_batch
        xt = torch.as_tensor(synthetic.mel_patches_db(tb, seed=300 + rank)).to(dev)
        # the reference's own initialisation: QR/LU 1x1, Glorot conv1/conv2, zero conv3, data-dependent ActNorm
        # (flow_builder.py:96-100); identical on every rank (same seed, rank-0-shaped minibatch)
        tprec = _lib.PREC_FP32 if args.train_fp32 else _lib.PREC_BF16
        tm = Glow(tcfg, init_glow_params(tcfg, seed=2, mode="faithful"), precision=tprec, device=local_rank)
        tm.init_actnorm(torch.as_tensor(synthetic.mel_patches_db(tb, seed=300)).to(dev))
        tm.enable_training()
        opt = dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7)
        last = [0.0]

        def step_train():
            last[0] = tg.distributed_train_step(tm, opt, xt, tb * world)

        nst = max(2, args.steps // 3)
        t_ms, t_launches, _ = timed_loop(step_train, nst, 3)     # eager (sizes the scratch), graph capture, first replay
        train = {"metric": "glow_train_samples_per_s", "value": world * tb * nst / (t_ms * 1e-3), "unit": "samples/s",
                 "steps_per_s": nst / (t_ms * 1e-3), "per_gpu_batch": tb, "global_batch": tb * world, "dtype": "f32" if args.train_fp32 else "bf16",
                 "ms_per_step": t_ms / nst, "gpu_launches": t_launches, "allreduce_bytes_per_step": int(tm.num_trainable * 4),
                 "alg_tflops": world * tb * nst * 3 * F_GLOW * (args.K / 40.0) / (t_ms * 1e-3) / 1e12,
                 "loss": float(last[0].item()), "loss_finite": bool(torch.isfinite(last[0]).all()),
                 "note": ("CUDA-core fp32 exact mode" if args.train_fp32 else
                          "tcgen05 forward / data-gradient / weight-gradient GEMMs (bf16 operands, fp32 accumulate), "
                          "fp32 master weights + Adamax, tile images rebuilt on the device every step, gradient pass replayed as a CUDA graph; "
                          "the loss value is that of the reference's own (quirk Q1/Q7) initialisation on synthetic patches, see DESIGN.md section 5")
                         + "; NCCL all-reduce of the flat gradient vector"}
        del tm

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    k_ms, k_launches, k_flops = prof
    achieved = k_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "k_nn_tc4<fwd> (fused conv3x3 -> conv1x1 -> conv3x3 coupling network, K-pipelined tcgen05)",
                "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_sustained"], "peak_source": f"{peaks['source']} sustained bf16 (kernel timed inside a long step)",
                "traffic": None, "launches": k_launches, "avg_launch_ms": k_ms / max(1, k_launches),
                "kernel_share_of_step": k_ms / total_ms if total_ms > 0 else None,
                "alg_flops_per_launch": k_flops / max(1, k_launches)}
    prof_path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof_path):
        try:
            with open(prof_path) as f:
                prof_js = json.load(f)
            # ncu dram bytes per pixel row of the captured block-1 launch x the average rows per launch of this run
            per_row = prof_js.get("dram_bytes_per_pixel_row")
            rows_per_launch = B * (48 * 32 + 24 * 16 + 12 * 8) / 3.0
            roofline["traffic"] = None if per_row is None else per_row * rows_per_launch
            roofline["traffic_note"] = "ncu dram__bytes_read+write per pixel row of a captured block-1 launch (profiles/ncu_summary.json) x mean rows per launch"
        except Exception:
            pass

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
    cpu = None
    if world == 1 and args.cpu_sample > 0:
        import torch as _t
        _, probe = cpu_log_prob_rate(2, args.K)                      # size the sample for ~15 s of CPU work
        args.cpu_sample = int(min(64, max(args.cpu_sample, round(15.0 / (probe / 2.0)))))
        rate, secs = cpu_log_prob_rate(args.cpu_sample, args.K)
        cpu = {"value": rate, "unit": "samples/s", "cores": _t.get_num_threads(), "kind": "port",
               "sample": f"{args.cpu_sample} patches, one pass of the reference-faithful graph (coupling network evaluated "
                         f"twice as in the TFP graph), torch-CPU fp32, {secs:.1f} s on {_t.get_num_threads()} threads "
                         f"({os.cpu_count()} logical cores)"}

    line = {
        "metric": "glow_log_prob_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": _config(args, per_gpu_batch=B),
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(x_host.numel() * 4),
                "d2h_bytes_per_step": int(B * 4), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "alg_tflops": value * F_GLOW * (args.K / 40.0) / 1e12,
        "directions": directions,
        "basis": basis,
        "basis_ncsn": ncsn or None,
        "train": train,
        "tc_cluster": args.cluster,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=2048, help="patches per GPU per step")
    ap.add_argument("--K", type=int, default=40, help="flow steps per block (40 = configs/melspec_glow.yml)")
    ap.add_argument("--cluster", type=int, default=1, help="TMA-multicast cluster size of the tcgen05 kernel")
    ap.add_argument("--basis-segments", type=int, default=256, help="segments per GPU of the BASIS leg (0 = skip)")
    ap.add_argument("--basis-T", type=int, default=2)
    ap.add_argument("--ncsn-segments", type=int, default=30, help="segments per GPU of the NCSN-BASIS legs (0 = skip)")
    ap.add_argument("--ncsn-T", type=int, default=2)
    ap.add_argument("--train-batch", type=int, default=32, help="per-GPU batch of the Glow train-step leg (0 = skip)")
    ap.add_argument("--train-fp32", action="store_true", help="train leg in the CUDA-core fp32 exact mode")
    ap.add_argument("--cpu-sample", type=int, default=4, help="patches of the CPU baseline sample (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
