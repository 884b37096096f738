import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_ncsn_train as T
from oracle import train_ncsn_oracle as to
ver = sys.argv[1] if len(sys.argv) > 1 else "v2"
cfg, params, sig, model, x, z, idx = T._setup(ver, "x3")
loss_ref, g_ref = to.dsm_loss_and_grads(cfg, params, sig, x, z, idx, 3)
runs = []
for r in range(3):
    grads, loss = model.train_grads(torch.as_tensor(x), torch.as_tensor(z), torch.as_tensor(idx), 3)
    runs.append((model.unflatten(grads), loss.item()))
print("loss", [r[1] for r in runs], loss_ref)
names = list(g_ref)
def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))
print("%-40s %10s %10s %10s" % ("tensor", "vs oracle", "run1-run0", "run2-run0"))
order = ["end_conv", "normalizer", "refine4", "refine3", "refine2", "refine1", "Res4_2", "Res4_1", "Res3_2", "Res3_1", "Res2_2", "Res2_1", "Res1_2", "Res1_1", "begin_conv"]
for pre in order:
    for n in names:
        if n.startswith(pre) and (n.endswith("kernel") or n.endswith("beta")):
            print("%-40s %10.2e %10.2e %10.2e" % (n, rel(runs[0][0][n], g_ref[n]), rel(runs[1][0][n], runs[0][0][n]), rel(runs[2][0][n], runs[0][0][n])))
