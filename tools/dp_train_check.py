"""Data-parallel training check: run under torchrun with any world size; writes the flat parameters after 2 steps."""
import argparse, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiosourcesep_b200 import GlowConfig, _lib
from audiosourcesep_b200.glow import Glow
from audiosourcesep_b200.weights import init_glow_params
from audiosourcesep_b200 import train_glow as tg

ap = argparse.ArgumentParser(); ap.add_argument("--out", required=True); a = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = GlowConfig(H=32, W=16, C=1, L=3, K=2, n_filters=64)
m = Glow(cfg, init_glow_params(cfg, seed=9), precision=_lib.PREC_FP32, device=lr); m.enable_training()
x = np.random.default_rng(0).uniform(-90, 10, (4, 32, 16, 1)).astype(np.float32)
local = 4 // world
opt = dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7)
losses = []
for step in range(2):
    xb = torch.as_tensor(x[rank * local:(rank + 1) * local]).cuda()
    noise = tg.noise_for(xb.shape, seed=1, step=step, rank_offset=rank * local, device=m.device)
    losses.append(float(tg.distributed_train_step(m, opt, xb, 4, noise=noise, sigma=0.3).item()))
if rank == 0:
    np.save(a.out, np.concatenate([m.get_flat().cpu().numpy(), np.array(losses, np.float32)]))
    print("world", world, "losses", losses, flush=True)
if world > 1: dist.destroy_process_group()
