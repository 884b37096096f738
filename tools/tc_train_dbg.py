import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiosourcesep_b200 import GlowConfig, _lib
from audiosourcesep_b200.glow import Glow
from audiosourcesep_b200.weights import init_glow_params
cfg = GlowConfig(H=32, W=16, C=1, L=3, K=2, n_filters=512, minval=0.0, maxval=1.0)
p = init_glow_params(cfg, seed=6, mode="perturbed")
x = torch.as_tensor(np.random.default_rng(2).uniform(0, 1, (4, 32, 16, 1)).astype(np.float32))
for prec in (_lib.PREC_FP32, _lib.PREC_BF16):
    m = Glow(cfg, p, precision=prec); m.enable_training()
    for it in range(3):
        g, loss = m.train_grads(x, global_batch=4)
        th0 = m.get_flat().clone()
        m.adamax_step(g, lr=1e-3)
        _, loss_dev = m.train_grads(x, global_batch=4)
        th = m.get_flat().clone()
        m.sync_host()
        lp_host = float(m.log_prob(x).sum().item())
        _, loss_dev2 = m.train_grads(x, global_batch=4)
        print(f"prec {prec} it {it}: loss before {loss.item():.3f} after(dev consts) {loss_dev.item():.3f} host log_prob-> {-lp_host/4:.3f} after(host consts) {loss_dev2.item():.3f} |dtheta|max {float((th-th0).abs().max()):.2e} gnorm {float(g.norm()):.3e}", flush=True)
