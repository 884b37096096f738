"""Scratch probe: eager vs CUDA-graph replay of log_prob / grad_log_prob at small batch."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from audiosourcesep_b200 import GlowConfig, _lib, synthetic
from audiosourcesep_b200.glow import Glow
from audiosourcesep_b200.weights import init_glow_params

def timeit(fn, n=20, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

cfg = GlowConfig()
p = init_glow_params(cfg, seed=2)
for mode, prec in (("bf16", _lib.PREC_BF16), ("fp16x3", _lib.PREC_FP16X3)):
    m = Glow(cfg, p, precision=prec)
    for N in (30, 32, 256):
        x = torch.as_tensor(synthetic.mel_patches_db(N, seed=0)).cuda()
        for name, fn in (("log_prob", lambda: m.log_prob(x)), ("grad", lambda: m.grad_log_prob(x))):
            fn(); torch.cuda.synchronize()
            n0 = _lib.launch_count(); fn(); nl = _lib.launch_count() - n0
            eager = timeit(fn)
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                fn()
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=s):
                    out = fn()
            graph = timeit(lambda: g.replay())
            print(f"{mode:7s} N={N:4d} {name:9s}: {nl:5d} launches, eager {eager:7.3f} ms, graph {graph:7.3f} ms ({eager/graph:4.2f}x)", flush=True)
