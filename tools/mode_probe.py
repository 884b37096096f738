"""Scratch probe: throughput of every Glow precision mode (log_prob / inverse / grad_log_prob) at a few batch sizes."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from audiosourcesep_b200 import GlowConfig, _lib, synthetic
from audiosourcesep_b200.glow import Glow
from audiosourcesep_b200.weights import init_glow_params

F = 48.22e9
def timeit(fn, n=3, w=1):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

cfg = GlowConfig()
p = init_glow_params(cfg, seed=2)
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["bf16", "bf16x2", "fp16x2", "fp16x3"]
batches = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["30", "2048"])]
P = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16, "bf16x2": _lib.PREC_BF16X2, "fp16x2": _lib.PREC_FP16X2, "fp16x3": _lib.PREC_FP16X3}
for mode in modes:
    m = Glow(cfg, p, precision=P[mode])
    for N in batches:
        base = synthetic.mel_patches_db(min(N, 64), seed=0)
        x = torch.as_tensor(np.concatenate([np.roll(base, 3 * i, axis=2) for i in range((N + 63) // 64)], 0)[:N]).cuda()
        n = 2 if (mode == "fp32" and N > 256) else 5
        ms = timeit(lambda: m.log_prob(x), n)
        _lib.tc_profile(True); m.log_prob(x); k_ms, k_n, k_fl = _lib.tc_profile_read(); _lib.tc_profile(False)
        z = m.forward(x)
        ms_i = timeit(lambda: m.inverse(z), n)
        rt = float((m.inverse(z) - x).abs().max()) / 120.0
        ms_g = timeit(lambda: m.grad_log_prob(x), max(2, n // 2))
        print(f"{mode:7s} N={N:5d}: log_prob {ms:8.2f} ms {N/ms*1e3:9.0f}/s {N*F/ms/1e9:7.0f} TF | kernel {k_ms:8.2f} ms ({k_n} launches, {k_fl/k_ms/1e9 if k_ms else 0:6.0f} TF) | "
              f"inverse {ms_i:8.2f} ms rt {rt:.1e} | grad {ms_g:8.2f} ms {N/ms_g*1e3:9.0f}/s", flush=True)
    del m
