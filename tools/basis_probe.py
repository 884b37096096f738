import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiosourcesep_b200 import GlowConfig, ops, _lib, synthetic
from audiosourcesep_b200.glow import Glow
from audiosourcesep_b200.weights import init_glow_params
from audiosourcesep_b200.ncsn import utils as bo
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = GlowConfig(K=40, minval=0.0, maxval=1.0)
m1 = Glow(cfg, init_glow_params(cfg, seed=2), precision=_lib.PREC_BF16)
m2 = Glow(cfg, init_glow_params(cfg, seed=3), precision=_lib.PREC_BF16)
mixed, _, _ = synthetic.basis_problem(32)
mixed = torch.as_tensor(np.concatenate([mixed] * (N // 32))).cuda()
x1, x2 = synthetic.langevin_init(N, seed=4)
t1, t2 = torch.as_tensor(x1).cuda(), torch.as_tensor(x2).cuda()
sig = bo.get_sigmas(1.0, 0.01, 10, "logarithmic"); eta, lam, ns = bo.langevin_step_constants(sig, 9)
def timeit(fn, it=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
print("grad m1 only       %.1f ms" % timeit(lambda: m1.grad_log_prob(t1)))
print("grad m2 only       %.1f ms" % timeit(lambda: m2.grad_log_prob(t2)))
print("grad m1 then m2    %.1f ms" % timeit(lambda: (m1.grad_log_prob(t1), m2.grad_log_prob(t2))))
w0 = time.time()
ms = timeit(lambda: ops.basis_glow_inner(m1, m2, mixed, t1, t2, 2, float(eta), float(lam), float(ns), seed=1))
print("basis_glow_inner T=2  %.1f ms per call (%.1f per Langevin step); wall %.2f s for 4 calls" % (ms, ms / 2, time.time() - w0))
print("finite", bool(torch.isfinite(t1).all()), float(t1.abs().max()))
def py_loop():
    for t in range(2):
        g1 = m1.grad_log_prob(t1); g2 = m2.grad_log_prob(t2)
        ops.langevin_step(t1, t2, g1, g2, mixed, float(eta), float(lam), float(ns), seed=1, step=t)
print("python loop T=2       %.1f ms per call" % timeit(py_loop))
ms = timeit(lambda: ops.basis_glow_inner(m1, m2, mixed, t1, t2, 2, float(eta), float(lam), float(ns), seed=1))
print("basis_glow_inner again %.1f ms per call" % ms)
ms = timeit(lambda: ops.basis_glow_inner(m1, m2, mixed, t1, t2, 1, float(eta), float(lam), float(ns), seed=1))
print("basis_glow_inner T=1   %.1f ms per call" % ms)
