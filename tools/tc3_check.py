"""Development check: N-quarter TS-form kernel (mode 3) vs the 8-warp single-CTA kernel (mode 0), forward and data gradient."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiosourcesep_b200 import GlowConfig, ops, _lib
from audiosourcesep_b200.glow import Glow
from audiosourcesep_b200.weights import init_glow_params
cfg = GlowConfig(H=32, W=16, C=1, L=3, K=2, n_filters=512, minval=0.0, maxval=1.0)
p = init_glow_params(cfg, seed=21, mode="perturbed")
m = Glow(cfg, p, precision=_lib.PREC_BF16)
ok = True
for block in range(3):
    Hb, Wb, Cb = cfg.level_shape(block)
    for N in (5, 64, 301, 2500):
        g = torch.Generator().manual_seed(block)
        state = torch.randn(N, Hb, Wb, Cb, generator=g) * 0.5
        gr = torch.randn(N, Hb, Wb, Cb, generator=g)
        ops.set_tc_pair_mode(0)
        r0 = m.coupling_nn(block, 0, state).cpu().numpy(); b0 = m.coupling_nn_backward(block, 0, state, gr).cpu().numpy()
        ops.set_tc_pair_mode(3)
        r1 = m.coupling_nn(block, 0, state).cpu().numpy(); b1 = m.coupling_nn_backward(block, 0, state, gr).cpu().numpy()
        torch.cuda.synchronize()
        e = (np.abs(r1 - r0).max() / np.abs(r0).max(), np.abs(b1 - b0).max() / np.abs(b0).max())
        good = e[0] < 1e-5 and e[1] < 1e-5
        ok &= good
        print(f"block {block} N={N}: fwd rel max|d|={e[0]:.3e} bwd rel max|d|={e[1]:.3e} identical={np.array_equal(r0, r1) and np.array_equal(b0, b1)}", flush=True)
print("TC3 CHECK", "OK" if ok else "MISMATCH")
