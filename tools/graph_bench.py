"""Langevin step time at the reference's n_mixed = 30 with eager launches vs CUDA-graph replay (Glow priors, NCSN v1/v2)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiosourcesep_b200 import GlowConfig, NCSNConfig, ops, _lib, synthetic
from audiosourcesep_b200.glow import Glow
from audiosourcesep_b200.weights import init_glow_params, init_ncsn_params
from audiosourcesep_b200.ncsn import utils as bo
from audiosourcesep_b200.ncsn.score_model import ScoreModel
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
T = 8
which = sys.argv[2] if len(sys.argv) > 2 else "glow,v1,v2"
mixed, _, _ = synthetic.basis_problem(N)
mixed = torch.as_tensor(mixed).cuda()
def timeit(fn, it=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
def run(name, call):
    for graphs in (False, True):
        _lib.basis_graphs(graphs)
        x1, x2 = synthetic.langevin_init(N, seed=4)
        t1, t2 = torch.as_tensor(x1).cuda(), torch.as_tensor(x2).cuda()
        ms = timeit(lambda: call(t1, t2))
        print(f"{name:18s} graphs={graphs!s:5s} {ms / T:8.2f} ms per Langevin step  ({N * T / ms * 1e3:8.1f} segment-steps/s) finite={bool(torch.isfinite(t1).all())}", flush=True)
if "glow" in which:
    cfg = GlowConfig(K=40, minval=0.0, maxval=1.0)
    sig = bo.get_sigmas(1.0, 0.01, 10, "logarithmic"); eta, lam, ns = bo.langevin_step_constants(sig, 9)
    for mode, prec in (("bf16", _lib.PREC_BF16), ("fp16x3", _lib.PREC_FP16X3)):
        m1 = Glow(cfg, init_glow_params(cfg, seed=2), precision=prec)
        m2 = Glow(cfg, init_glow_params(cfg, seed=3), precision=prec)
        run(f"glow {mode}", lambda a, b: ops.basis_glow_inner(m1, m2, mixed, a, b, T, float(eta), float(lam), float(ns), seed=1))
        del m1, m2
for ver, ncfg in (("v1", NCSNConfig(version="v1", ngf=192, num_classes=10, sigma1=1.0)),
                  ("v2", NCSNConfig(version="v2", ngf=128, num_classes=200, sigma1=30.0))):
    if ver not in which:
        continue
    sig = bo.get_sigmas(ncfg.sigma1, ncfg.sigmaL, ncfg.num_classes, "logarithmic")
    idx = ncfg.num_classes - 1
    eta, lam, ns = bo.langevin_step_constants(sig, idx)
    for mode, prec in (("bf16", _lib.PREC_BF16), ("bf16x3", _lib.PREC_BF16X3)):
        s1 = ScoreModel(ncfg, init_ncsn_params(ncfg, seed=11), sigmas=sig, precision=prec)
        s2 = ScoreModel(ncfg, init_ncsn_params(ncfg, seed=12), sigmas=sig, precision=prec)
        run(f"ncsn {ver} {mode}", lambda a, b: ops.basis_ncsn_inner(s1, s2, mixed, a, b, idx, T, float(eta), float(lam), float(ns), seed=2))
        del s1, s2
