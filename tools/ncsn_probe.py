"""Development probe: NCSN forward parity vs the oracle and device timing."""
import argparse, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiosourcesep_b200 import NCSNConfig, synthetic, _lib
from audiosourcesep_b200.weights import init_ncsn_params
from audiosourcesep_b200.ncsn.score_model import ScoreModel
from audiosourcesep_b200.ncsn import utils as bo

ap = argparse.ArgumentParser()
ap.add_argument("--version", default="v1")
ap.add_argument("--batches", type=int, nargs="+", default=[30])
ap.add_argument("--check", action="store_true")
ap.add_argument("--x3", action="store_true", help="split-bf16 mode (ASEP_PREC_BF16X3)")
a = ap.parse_args()
cfg = NCSNConfig(version="v1", ngf=192, num_classes=10) if a.version == "v1" else NCSNConfig(version="v2", ngf=128, num_classes=200, sigma1=30.0)
sig = bo.get_sigmas(cfg.sigma1, cfg.sigmaL, cfg.num_classes, cfg.progression)
p = init_ncsn_params(cfg, seed=5, mode="perturbed")
t0 = time.time(); m = ScoreModel(cfg, p, sigmas=sig, precision=_lib.PREC_BF16X3 if a.x3 else _lib.PREC_BF16); print(f"model ready {time.time()-t0:.1f}s", flush=True)
if a.check:
    x = synthetic.normalise(synthetic.mel_patches_db(2, seed=1))
    idx = np.array([0, cfg.num_classes - 1], dtype=np.int32)
    got = m([torch.as_tensor(x), torch.as_tensor(idx)]).cpu().numpy()
    from oracle.ncsn_oracle import NCSNOracle          # development check only
    o = NCSNOracle(cfg, p, sigmas=sig, dtype=torch.float32)
    want = o.score(x, idx).numpy()
    print("finite", np.isfinite(got).all(), "rel err", np.linalg.norm(got - want) / np.linalg.norm(want),
          [float(np.linalg.norm(got[i] - want[i]) / np.linalg.norm(want[i])) for i in range(2)], flush=True)
GF = 266.96e9 if a.version == "v1" else 118.66e9
for N in a.batches:
    x = torch.as_tensor(synthetic.normalise(synthetic.mel_patches_db(min(N, 32), seed=0))).cuda()
    x = x.repeat((N + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:N].contiguous()
    idx = torch.full((N,), 3, dtype=torch.int32, device="cuda")
    for _ in range(2): m([x, idx])
    torch.cuda.synchronize()
    _lib.conv_profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 3
    e0.record()
    for _ in range(it): m([x, idx])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    kms, kn, kfl = _lib.conv_profile_read(); _lib.conv_profile(False)
    print(f"N={N}: {ms:.2f} ms/eval  {N/ms*1e3:.1f} evals/s  {N*GF/ms/1e9:.1f} TFLOP/s alg;  conv kernels {kms/it:.2f} ms "
          f"({kn//it} launches) = {kfl/kms/1e9:.1f} TFLOP/s; launches total {_lib.launch_count()}", flush=True)
