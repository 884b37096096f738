"""Quick device-side timing of the Glow paths (development aid; bench.py is the contract)."""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from audiosourcesep_b200 import GlowConfig, synthetic, _lib, ops          # noqa: E402
from audiosourcesep_b200.glow import Glow                                   # noqa: E402
from audiosourcesep_b200.weights import init_glow_params                    # noqa: E402

F_GLOW = 48.22e9  # FLOP per sample per NN pass


def timeit(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=40)
    ap.add_argument("--batches", type=int, nargs="+", default=[32, 256, 1024])
    ap.add_argument("--clusters", type=int, nargs="+", default=[1])
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--grad", action="store_true")
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--pair", type=int, nargs="+", default=[3])
    a = ap.parse_args()
    cfg = GlowConfig(K=a.K)
    t0 = time.time()
    p = init_glow_params(cfg, seed=2)
    m = Glow(cfg, p, precision=_lib.PREC_BF16)
    print(f"model ready in {time.time() - t0:.1f}s", flush=True)
    scale = a.K / 40.0
    for N in a.batches:
        x = torch.as_tensor(synthetic.mel_patches_db(min(N, 64), seed=0)).cuda()
        x = x.repeat((N + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:N].contiguous()
        for cs, pair in [(c, q) for q in a.pair for c in (a.clusters if q == 0 else [1])]:
            ms = timeit(lambda: m.log_prob(x), a.iters)
            tf = N * F_GLOW * scale / (ms * 1e-3) / 1e12
            print(f"log_prob  N={N:5d} cluster={cs} pair={pair}: {ms:9.3f} ms  {N / ms * 1e3:10.1f} samples/s  {tf:8.1f} TFLOP/s", flush=True)
            if a.grad:
                ms = timeit(lambda: m.grad_log_prob(x), a.iters)
                tf = N * 2 * F_GLOW * scale / (ms * 1e-3) / 1e12
                print(f"grad_logp N={N:5d} cluster={cs} pair={pair}: {ms:9.3f} ms  {N / ms * 1e3:10.1f} samples/s  {tf:8.1f} TFLOP/s (2F alg.)", flush=True)
    if a.fp32:
        m.prepare(_lib.PREC_FP32)
        x = torch.as_tensor(synthetic.mel_patches_db(32, seed=0)).cuda()
        ms = timeit(lambda: m.log_prob(x), 2, warmup=1)
        print(f"fp32 log_prob N=32: {ms:9.3f} ms  {32 / ms * 1e3:10.1f} samples/s")
    print("launches", _lib.launch_count())


if __name__ == "__main__":
    main()
