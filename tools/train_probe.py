import sys, os, time, argparse, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiosourcesep_b200 import GlowConfig, _lib, synthetic
from audiosourcesep_b200.glow import Glow
from audiosourcesep_b200.weights import init_glow_params
ap = argparse.ArgumentParser(); ap.add_argument("--K", type=int, default=40); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=4); ap.add_argument("--lr", type=float, default=1e-3); ap.add_argument("--fp32", action="store_true")
a = ap.parse_args()
cfg = GlowConfig(K=a.K)
x = torch.as_tensor(synthetic.mel_patches_db(a.batch, seed=300)).cuda()
m = Glow(cfg, init_glow_params(cfg, seed=2, mode="faithful"), precision=_lib.PREC_FP32 if a.fp32 else _lib.PREC_BF16)
m.init_actnorm(x); m.enable_training()
for it in range(a.steps):
    torch.cuda.synchronize(); t0 = time.time()
    g, loss = m.train_grads(x, global_batch=a.batch)
    torch.cuda.synchronize(); t1 = time.time()
    m.adamax_step(g, lr=a.lr)
    torch.cuda.synchronize(); t2 = time.time()
    print(f"step {it}: loss {loss.item():.4e}  grads {1e3*(t1-t0):.1f} ms  adamax+refresh {1e3*(t2-t1):.1f} ms  |g| {float(g.norm()):.3e}", flush=True)
