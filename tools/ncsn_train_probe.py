"""NCSN train step at the reference sizes (96 x 64, batch 32): time per step, memory, loss trajectory."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiosourcesep_b200 import NCSNConfig, _lib, synthetic
from audiosourcesep_b200.weights import init_ncsn_params
from audiosourcesep_b200.ncsn import utils as bo
from audiosourcesep_b200.ncsn.score_model import ScoreModel
ver = sys.argv[1] if len(sys.argv) > 1 else "v1"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
cfg = NCSNConfig(version="v1", ngf=192, num_classes=10, sigma1=1.0) if ver == "v1" else NCSNConfig(version="v2", ngf=128, num_classes=200, sigma1=30.0)
sig = bo.get_sigmas(cfg.sigma1, cfg.sigmaL, cfg.num_classes, "logarithmic")
m = ScoreModel(cfg, init_ncsn_params(cfg, seed=11, mode="faithful"), sigmas=sig, precision=_lib.PREC_BF16X3 if mode == "x3" else _lib.PREC_BF16)
m.enable_training()
x = torch.as_tensor(synthetic.normalise(synthetic.mel_patches_db(N, seed=0))).cuda()
g = torch.Generator(device="cuda").manual_seed(0)
opt = dict(kind="adam", lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-7)
gflop = {"v1": 266.96, "v2": 118.66}[ver]
def step():
    idx = torch.randint(0, cfg.num_classes, (1,), generator=g, device="cuda").repeat(N).to(torch.int32)
    z = torch.randn(x.shape, generator=g, device="cuda")
    grads, loss = m.train_grads(x, z, idx, N)
    m.apply_gradients(grads, opt)
    return loss
losses = [step().item() for _ in range(2)]
torch.cuda.synchronize()
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps): losses.append(step().item())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"{ver} {mode} batch {N}: {ms:.1f} ms per train step, {N / ms * 1e3:.1f} samples/s, {(_lib.launch_count() - n0) // steps} launches per step, "
      f"{3 * gflop * N / ms / 1e0 * 1e-3:.1f} TFLOP/s algorithmic (3 GEMMs per conv), mem {torch.cuda.mem_get_info()[0] / 2**30:.1f} GiB free of {torch.cuda.mem_get_info()[1] / 2**30:.1f}")
print("losses", [round(v, 2) for v in losses])
