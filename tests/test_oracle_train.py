"""CPU checks of the training-step oracle: autograd gradients vs central finite differences, Adamax known answer."""
import numpy as np

from audiosourcesep_b200 import GlowConfig
from audiosourcesep_b200.weights import init_glow_params
from oracle import train_oracle as to
from oracle.glow_oracle import GlowOracle

import torch


def test_oracle_gradients_match_finite_differences():
    cfg = GlowConfig(H=8, W=8, C=1, L=2, K=1, n_filters=8, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=3, mode="perturbed")
    x = np.random.default_rng(0).uniform(0, 1, (2, 8, 8, 1))
    loss, g = to.loss_and_grads(cfg, p, x, global_batch=4)

    def f(params):
        return float(-GlowOracle(cfg, params, dtype=torch.float64).log_prob(x).sum() / 4.0)

    assert abs(f(p) - loss) < 1e-9
    rng = np.random.default_rng(1)
    for name in ["b0/s0/actnorm/log_scale", "b1/s0/inv1x1/log_S", "b0/s0/inv1x1/L", "b0/s0/nn/conv2/kernel",
                 "b1/s0/nn/bn1/gamma", "b0/s0/nn/conv3/bias", "prior/log_scale", "b1/s0/nn/conv1/kernel"]:
        a = p[name].astype(np.float64)
        idx = tuple(rng.integers(0, s) for s in a.shape)
        if name.endswith("/L"):
            idx = (a.shape[0] - 1, 0)            # a strictly-lower (unmasked) entry
        h = 1e-5
        pp, pm = dict(p), dict(p)
        ap, am = a.copy(), a.copy()
        ap[idx] += h
        am[idx] -= h
        pp[name], pm[name] = ap, am
        fd = (f(pp) - f(pm)) / (2 * h)
        assert abs(fd - g[name][idx]) <= 1e-5 * max(1.0, abs(fd)), (name, fd, g[name][idx])


def test_adamax_known_answer():
    theta, m, u = np.array([1.0, -2.0], np.float32), np.zeros(2, np.float32), np.zeros(2, np.float32)
    g = np.array([0.5, -0.25], np.float32)
    theta, m, u = to.adamax_update(theta, g, m, u, t=1)
    # first step: m = 0.1 g, u = |g|, lr_t = 1e-3 / 0.1  ->  theta -= 1e-3 * sign(g) (up to eps)
    np.testing.assert_allclose(theta, [1.0 - 1e-3, -2.0 + 1e-3], rtol=0, atol=1e-6)
    np.testing.assert_allclose(m, 0.1 * g, atol=1e-8)
    np.testing.assert_allclose(u, np.abs(g), atol=0)
