"""log_prob / inverse / grad_log_prob at the reference's batch sizes: eager launches vs graph replay."""
import os, sys, subprocess, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiosourcesep_b200 import GlowConfig, _lib, synthetic
from audiosourcesep_b200.glow import Glow
from audiosourcesep_b200.weights import init_glow_params
cfg = GlowConfig(K=40)
m = Glow(cfg, init_glow_params(cfg, seed=2), precision=_lib.PREC_BF16)
def timeit(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for N in (30, 32, 64, 128):
    x = torch.as_tensor(synthetic.mel_patches_db(N, seed=0)).cuda()
    z = m.forward(x)
    print(f"N={N:4d} graphs={'off' if os.environ.get('ASEP_NO_GRAPH') else 'on '} log_prob {timeit(lambda: m.log_prob(x)):7.3f} ms  inverse {timeit(lambda: m.inverse(z)):7.3f} ms  grad {timeit(lambda: m.grad_log_prob(x)):7.3f} ms", flush=True)
