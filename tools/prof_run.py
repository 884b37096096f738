"""Profiling driver (round 2): one region of interest (NVTX range "roi") per invocation, for `ncu --nvtx --nvtx-include "roi/"`.

    python tools/prof_run.py logprob --prec bf16 --n 2048
    python tools/prof_run.py ncsn --version v1 --n 30
    python tools/prof_run.py langevin --n 4096
    python tools/prof_run.py train --n 32            (run with ASEP_NO_GRAPH=1 so that the kernels are launched eagerly)
    python tools/prof_run.py basis --prec fp16x3 --n 30
    python tools/prof_run.py gemm                    (torch.matmul bf16 8192^3: the GEMM that defines the measured peak)
"""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from audiosourcesep_b200 import GlowConfig, NCSNConfig, _lib, ops, synthetic  # noqa: E402
from audiosourcesep_b200.glow import Glow  # noqa: E402
from audiosourcesep_b200.weights import init_glow_params, init_ncsn_params  # noqa: E402

P = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16, "bf16x2": _lib.PREC_BF16X2, "fp16x2": _lib.PREC_FP16X2,
     "fp16x3": _lib.PREC_FP16X3}


def patches(n, seed=0):
    base = synthetic.mel_patches_db(min(n, 64), seed=seed)
    return torch.as_tensor(np.concatenate([np.roll(base, 3 * i, axis=2) for i in range((n + 63) // 64)], 0)[:n]).cuda()


def roi(fn, reps=1, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    t0 = time.time()
    torch.cuda.nvtx.range_push("roi")
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print(f"roi: {reps} rep(s), {(time.time() - t0) * 1e3 / reps:.3f} ms each, {(_lib.launch_count() - n0) // reps} libasep launches per rep", flush=True)


ap = argparse.ArgumentParser()
ap.add_argument("what")
ap.add_argument("--prec", default="bf16")
ap.add_argument("--n", type=int, default=2048)
ap.add_argument("--version", default="v1")
ap.add_argument("--K", type=int, default=40)
ap.add_argument("--op", default="log_prob")
a = ap.parse_args()

if a.what == "logprob":
    m = Glow(GlowConfig(K=a.K), init_glow_params(GlowConfig(K=a.K), seed=2), precision=P[a.prec])
    x = patches(a.n)
    if a.op == "log_prob":
        roi(lambda: m.log_prob(x))
    elif a.op == "grad":
        roi(lambda: m.grad_log_prob(x))
    else:
        z = m.forward(x)
        roi(lambda: m.inverse(z))
elif a.what == "basis":
    cfg = GlowConfig(K=a.K, minval=0.0, maxval=1.0)
    m1 = Glow(cfg, init_glow_params(cfg, seed=2), precision=P[a.prec])
    m2 = Glow(cfg, init_glow_params(cfg, seed=3), precision=P[a.prec])
    mixed, _, _ = synthetic.basis_problem(a.n)
    x1, x2 = synthetic.langevin_init(a.n, seed=4)
    t1, t2, md = torch.as_tensor(x1).cuda(), torch.as_tensor(x2).cuda(), torch.as_tensor(mixed).cuda()
    roi(lambda: ops.basis_glow_inner(m1, m2, md, t1, t2, 1, 2e-5, 1e4, 6.3e-3, seed=1))
elif a.what == "ncsn":
    from audiosourcesep_b200.ncsn import utils as bo
    from audiosourcesep_b200.ncsn.score_model import ScoreModel
    cfg = NCSNConfig(version="v1", ngf=192, num_classes=10) if a.version == "v1" else NCSNConfig(version="v2", ngf=128, num_classes=200, sigma1=30.0)
    sig = bo.get_sigmas(cfg.sigma1, cfg.sigmaL, cfg.num_classes, "logarithmic")
    prec = _lib.PREC_BF16X3 if a.prec == "bf16x3" else _lib.PREC_BF16
    m = ScoreModel(cfg, init_ncsn_params(cfg, seed=11), sigmas=sig, precision=prec)
    x = torch.as_tensor(synthetic.langevin_init(a.n, seed=7)[0]).cuda()
    idx = torch.full((a.n,), cfg.num_classes - 1, dtype=torch.int32, device="cuda")
    roi(lambda: m([x, idx], training=True))
elif a.what == "langevin":
    g = torch.Generator(device="cuda").manual_seed(0)
    t = [torch.rand((a.n, 96, 64, 1), device="cuda", generator=g) for _ in range(5)]
    roi(lambda: ops.langevin_step(t[0], t[1], t[2], t[3], t[4], 2e-5, 1e4, 6.3e-3, seed=1, step=0), reps=2)
elif a.what == "train":
    from audiosourcesep_b200 import train_glow as tg
    cfg = GlowConfig(K=a.K)
    tm = Glow(cfg, init_glow_params(cfg, seed=2, mode="faithful"), precision=_lib.PREC_BF16)
    tm.init_actnorm(torch.as_tensor(synthetic.mel_patches_db(a.n, seed=300)).cuda())
    tm.enable_training()
    xt = torch.as_tensor(synthetic.mel_patches_db(a.n, seed=300)).cuda()
    opt = dict(kind="adamax", lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7)
    roi(lambda: tg.distributed_train_step(tm, opt, xt, a.n), warm=3)
elif a.what == "ncsn_train":
    from audiosourcesep_b200.ncsn import utils as bo
    from audiosourcesep_b200.ncsn.score_model import ScoreModel
    cfg = NCSNConfig(version="v1", ngf=192, num_classes=10) if a.version == "v1" else NCSNConfig(version="v2", ngf=128, num_classes=200, sigma1=30.0)
    sig = bo.get_sigmas(cfg.sigma1, cfg.sigmaL, cfg.num_classes, "logarithmic")
    m = ScoreModel(cfg, init_ncsn_params(cfg, seed=11, mode="faithful"), sigmas=sig, precision=_lib.PREC_BF16X3 if a.prec == "bf16x3" else _lib.PREC_BF16)
    m.enable_training()
    x = torch.as_tensor(synthetic.normalise(synthetic.mel_patches_db(a.n, seed=0))).cuda()
    z = torch.randn(x.shape, device="cuda")
    idx = torch.full((a.n,), cfg.num_classes // 2, dtype=torch.int32, device="cuda")
    opt = dict(kind="adam", lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-7)

    def tstep():
        g, _ = m.train_grads(x, z, idx, a.n)
        m.apply_gradients(g, opt)

    roi(tstep, warm=2)
elif a.what == "gemm":
    A = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    B = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        A @ B
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        A @ B
    e1.record()
    torch.cuda.synchronize()
    print(f"cuBLAS bf16 8192^3: {2 * 8192 ** 3 * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e12:.1f} TFLOP/s", flush=True)
    torch.cuda.nvtx.range_push("roi")
    A @ B
    A @ B
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
else:
    raise SystemExit("unknown target")
