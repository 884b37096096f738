"""Pins the Glow oracle against the reference's own known answers
(unittest_flow_models.py:25-51 factory, :122-186 cases) and against independent checks."""
import math

import numpy as np
import pytest
import torch

from audiosourcesep_b200 import GlowConfig
from audiosourcesep_b200.weights import init_glow_params, count_trainable
from oracle.glow_oracle import GlowOracle, squeeze, unsqueeze, conv2d_same, inv1x1_weight, inv1x1_weight_inverse

LOG2 = math.log(2.0)


def toy_nn(xb):  # unittest_flow_models.py:76-79
    return LOG2 * torch.ones_like(xb), torch.ones_like(xb)


def tiny_cfg(H, W, C, L=2, K=2, F=8):
    return GlowConfig(H=H, W=W, C=C, L=L, K=K, n_filters=F, learntop=True, minval=0.0, maxval=1.0)


def test_param_count_known_answer():
    assert count_trainable(GlowConfig()) == 39_598_720 + 2 * 6_144


def test_squeeze_channel_order_and_roundtrip():
    x = torch.arange(2 * 4 * 6 * 3, dtype=torch.float64).reshape(2, 4, 6, 3)
    y = squeeze(x)
    for c in range(3):
        for dh in range(2):
            for dw in range(2):
                assert torch.equal(y[:, :, :, c * 4 + dh * 2 + dw], x[:, dh::2, dw::2, c])
    assert torch.equal(unsqueeze(y), x)


def test_coupling_split_logdet_known_answer():
    # AffineCouplingLayerSplit on (2,2,2) with the toy net: 4*log2 (unittest_flow_models.py:141-146)
    cfg = GlowConfig(H=4, W=4, C=2, L=2, K=1, n_filters=8, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=0, mode="faithful")
    o = GlowOracle(cfg, p, coupling_nn=toy_nn)
    x = torch.randn(1, 2, 2, 8, dtype=torch.float64)
    pre = "b0/s0/"
    # neutralise actnorm / 1x1 so that only the coupling contributes
    o.p[pre + "inv1x1/log_S"] = torch.zeros(8, dtype=torch.float64)
    y, ld = o.step_forward(x, pre)
    assert ld.item() == pytest.approx(2 * 2 * 4 * LOG2, rel=1e-12)   # H*W*(C/2) elements scaled
    xr = o.step_inverse(y, pre)
    assert torch.allclose(xr, x, atol=1e-12)


def test_actnorm_data_init_known_answer():
    # minibatch of 2s and 1s -> scale exactly 2 -> logdet H*W*log 2 (unittest_flow_models.py:149-154)
    cfg = GlowConfig(H=4, W=4, C=1, L=2, K=1, n_filters=8, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=0, mode="faithful")
    o = GlowOracle(cfg, p, coupling_nn=toy_nn)
    mb = np.concatenate([2 * np.ones((1, 4, 4, 1)), np.ones((1, 4, 4, 1))], 0)
    o.init_actnorm(mb)
    ls = o.p["b0/s0/actnorm/log_scale"]
    assert torch.allclose(ls, torch.full((4,), LOG2, dtype=torch.float64), atol=1e-6)
    assert (2 * 2 * ls.sum()).item() == pytest.approx(2 * 2 * 4 * LOG2, rel=1e-6)
    assert torch.allclose(o.p["b0/s0/actnorm/shift"], torch.full((4,), -2.0, dtype=torch.float64), atol=1e-6)  # mean 1.0 after the -0.5 shift, std 0.5


def test_inv1x1_inverse_and_logdet():
    cfg = tiny_cfg(8, 8, 1)
    p = init_glow_params(cfg, seed=3, mode="perturbed")
    args = [torch.as_tensor(p["b1/s0/inv1x1/" + n], dtype=torch.float64) for n in ("P", "L", "U", "log_S", "sign_S")]
    Wm, Wi = inv1x1_weight(*args), inv1x1_weight_inverse(*args)
    assert torch.allclose(Wm @ Wi, torch.eye(Wm.shape[0], dtype=torch.float64), atol=1e-10)
    assert torch.slogdet(Wm)[1].item() == pytest.approx(args[3].sum().item(), abs=1e-10)


@pytest.mark.parametrize("L,H,W", [(2, 4, 4), (3, 8, 8), (4, 16, 16)])
@pytest.mark.parametrize("toy", [True, False])
def test_glow_roundtrip_and_fldj(L, H, W, toy):
    # GlowBijector_{2,3,4}blocks invertibility (unittest_flow_models.py:171-186) + fldj vs autograd Jacobian
    cfg = tiny_cfg(H, W, 1, L=L, K=2, F=8)
    p = init_glow_params(cfg, seed=5, mode="perturbed")
    o = GlowOracle(cfg, p, coupling_nn=toy_nn if toy else None)
    x = torch.rand(2, H, W, 1, dtype=torch.float64)
    z, fldj = o.forward(x)
    assert z.shape == (2,) + cfg.latent_shape
    xr = o.inverse(z)
    assert torch.allclose(xr, x, atol=1e-9)
    if H * W <= 64:
        J = torch.autograd.functional.jacobian(lambda a: o.forward(a[None])[0].reshape(-1), x[0])
        J = J.reshape(H * W, H * W)
        assert torch.slogdet(J)[1].item() == pytest.approx(fldj[0].item(), abs=1e-8)


def test_grad_log_prob_finite_difference():
    cfg = tiny_cfg(8, 8, 1, L=3, K=2, F=8)
    p = init_glow_params(cfg, seed=7, mode="perturbed")
    o = GlowOracle(cfg, p)
    x = torch.rand(1, 8, 8, 1, dtype=torch.float64)
    g, lp = o.grad_log_prob(x)
    eps = 1e-6
    for idx in [(0, 0, 0, 0), (0, 3, 5, 0), (0, 7, 7, 0)]:
        xp, xm = x.clone(), x.clone()
        xp[idx] += eps
        xm[idx] -= eps
        fd = (o.log_prob(xp) - o.log_prob(xm)).item() / (2 * eps)
        assert fd == pytest.approx(g[idx].item(), rel=1e-5, abs=1e-6)


def test_conv_same_matches_numpy_loops():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 5, 4, 2))
    k = rng.standard_normal((3, 3, 2, 3))
    b = rng.standard_normal(3)
    for dil in (1, 2):
        ref = np.zeros((1, 5, 4, 3))
        for h in range(5):
            for w in range(4):
                for i in range(3):
                    for j in range(3):
                        hh, ww = h + (i - 1) * dil, w + (j - 1) * dil
                        if 0 <= hh < 5 and 0 <= ww < 4:
                            ref[0, h, w] += x[0, hh, ww] @ k[i, j]
        ref += b
        got = conv2d_same(torch.as_tensor(x), torch.as_tensor(k), torch.as_tensor(b), dil).numpy()
        np.testing.assert_allclose(got, ref, atol=1e-12)


def test_faithful_init_first_actnorm_normalises():
    cfg = tiny_cfg(8, 8, 1, L=3, K=2, F=8)
    p = init_glow_params(cfg, seed=1, mode="faithful")
    o = GlowOracle(cfg, p)
    mb = np.random.default_rng(0).uniform(0, 1, (6, 8, 8, 1))
    o.init_actnorm(mb)
    s = squeeze(o.pre_forward(torch.as_tensor(mb)))
    a = s * torch.exp(o.p["b0/s0/actnorm/log_scale"]) + o.p["b0/s0/actnorm/shift"]
    assert torch.allclose(a.mean(dim=(0, 1, 2)), torch.zeros(4, dtype=torch.float64), atol=1e-6)
    assert torch.allclose(a.std(dim=(0, 1, 2), unbiased=False), torch.ones(4, dtype=torch.float64), atol=1e-5)
    # faithful conv3 == 0 -> couplings are the identity (quirk Q6): logdet has no data term
    x = torch.rand(2, 8, 8, 1, dtype=torch.float64)
    _, ld = o.forward(x)
    assert abs(ld[0].item() - ld[1].item()) < 1e-9


def test_log_prob_reference_graph_equals_single_pass():
    cfg = tiny_cfg(8, 8, 1, L=3, K=2, F=8)
    p = init_glow_params(cfg, seed=2, mode="perturbed")
    o = GlowOracle(cfg, p)
    x = torch.rand(2, 8, 8, 1, dtype=torch.float64)
    assert torch.allclose(o.log_prob(x), o.log_prob_reference_graph(x))
