import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiosourcesep_b200 import GlowConfig, ops, _lib, synthetic
from audiosourcesep_b200.glow import Glow
from audiosourcesep_b200.weights import init_glow_params
cfg = GlowConfig(K=1)
m = Glow(cfg, init_glow_params(cfg, seed=2), precision=_lib.PREC_BF16)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = torch.as_tensor(synthetic.mel_patches_db(64, seed=0)).cuda().repeat(N // 64, 1, 1, 1).contiguous()
for _ in range(2): m.log_prob(x)
torch.cuda.synchronize()
os.environ["ASEP_TC_DBG_TIMING"] = "1"
m.log_prob(x)
if len(sys.argv) > 2: m.grad_log_prob(x)
torch.cuda.synchronize()
