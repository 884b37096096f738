"""Pins the BASIS / NCSN oracles: sigma schedule golden, parameter-count known answers,
primitive semantics vs independent numpy loops, Langevin algebra."""
import json
import os

import numpy as np
import pytest
import torch

from audiosourcesep_b200 import NCSNConfig
from audiosourcesep_b200.weights import count_ncsn_params, init_ncsn_params
from oracle import basis_oracle as bo
from oracle import ncsn_oracle as no

G = os.path.join(os.path.dirname(__file__), "golden")


def test_sigmas_match_reference_run_log():
    gold = json.load(open(os.path.join(G, "sigmas_v1.json")))
    s = bo.get_sigmas(gold["sigma1"], gold["sigmaL"], gold["num_classes"], gold["progression"])
    assert s.dtype == np.float32
    np.testing.assert_array_equal(s, np.asarray(gold["printed"], np.float64).astype(np.float32))
    s2 = bo.get_sigmas(gold["sigma1"], gold["sigmaL"], gold["num_classes"], "geometric")
    np.testing.assert_allclose(s2, s, rtol=1e-6)


def test_param_count_known_answers():
    gold = json.load(open(os.path.join(G, "param_counts.json")))
    assert count_ncsn_params(NCSNConfig(version="v1", ngf=192, num_classes=10)) == gold["ncsn_v1_ngf192_classes10"]
    assert count_ncsn_params(NCSNConfig(version="v2", ngf=128, num_classes=200)) == 29_695_233
    # the oracle's OWN walk of the reference constructors (no table shared with the product) gives the same answers
    from oracle.ncsn_oracle import count_params_by_walk
    assert count_params_by_walk("v1", 192, 10) == gold["ncsn_v1_ngf192_classes10"]
    assert count_params_by_walk("v2", 128, 200) == 29_695_233
    for ngf, nc in ((64, 10), (96, 10), (32, 3)):
        for v in ("v1", "v2"):
            assert count_params_by_walk(v, ngf, nc) == count_ncsn_params(NCSNConfig(version=v, ngf=ngf, num_classes=nc))


def test_mixing_db_is_power_mean_and_gradient():
    g, grad_g = bo.mixing_process("melspec", "dB")
    rng = np.random.default_rng(0)
    a, b = rng.uniform(0, 1, (2, 3, 4, 1)).astype(np.float32), rng.uniform(0, 1, (2, 3, 4, 1)).astype(np.float32)
    ref = 10 * np.log10((10 ** (a.astype(np.float64) / 10) + 10 ** (b.astype(np.float64) / 10)) / 2)
    np.testing.assert_allclose(g(a, b), ref, rtol=2e-6, atol=2e-6)
    ga, gb = grad_g(a, b)
    np.testing.assert_allclose(ga + gb, 1.0, rtol=1e-6)
    eps = 1e-3
    fd = (g(a + eps, b).astype(np.float64) - g(a - eps, b)) / (2 * eps)
    np.testing.assert_allclose(ga, fd, atol=2e-3)


def test_step_constants_and_update_algebra():
    sig = bo.get_sigmas(1.0, 0.01, 10, "logarithmic")
    eta, lam, ns = bo.step_constants(sig, 0)
    assert eta.dtype == np.float32 and lam.dtype == np.float32
    assert float(eta) == pytest.approx(2e-5 * 1e4, rel=1e-6)
    assert float(lam) == pytest.approx(1.0, rel=1e-7)
    eta9, lam9, _ = bo.step_constants(sig, 9)
    assert float(eta9) == pytest.approx(2e-5, rel=1e-6) and float(lam9) == pytest.approx(1e4, rel=1e-6)
    g, grad_g = bo.mixing_process("melspec", "dB")
    rng = np.random.default_rng(1)
    shp = (2, 4, 4, 1)
    x1, x2, mixed = (rng.uniform(0, 1, shp).astype(np.float32) for _ in range(3))
    z = np.zeros(shp, np.float32)
    # zero score, zero noise, mixture already consistent -> fixed point
    m = g(x1, x2)
    y1, y2 = bo.langevin_update(x1, x2, z, z, m, z, z, eta, lam, ns, g, grad_g)
    np.testing.assert_array_equal(y1, x1)
    np.testing.assert_array_equal(y2, x2)
    # both sources are updated from the OLD states (Q11)
    y1, y2 = bo.langevin_update(x1, x2, z, z, mixed, z, z, eta, lam, ns, g, grad_g)
    ga, gb = grad_g(x1, x2)
    np.testing.assert_allclose(y2, x2 + eta * (lam * gb * (mixed - m)), rtol=1e-6)


def test_post_processing_clip():
    x = np.array([-0.2, 0.0, 0.5, 1.0, 1.3], np.float32)
    np.testing.assert_allclose(bo.post_processing(x), [-100, -100, -40, 20, 20])


def test_pool_and_resize_semantics_vs_loops():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 6, 5, 2))
    xt = torch.as_tensor(x)
    avg = no.avg_pool5_same(xt).numpy()
    mx = no.max_pool5_same(xt).numpy()
    for h in range(6):
        for w in range(5):
            win = x[0, max(0, h - 2):h + 3, max(0, w - 2):w + 3]
            np.testing.assert_allclose(avg[0, h, w], win.mean(axis=(0, 1)), atol=1e-12)   # divisor = in-bounds taps
            np.testing.assert_allclose(mx[0, h, w], win.max(axis=(0, 1)), atol=1e-12)
    up = no.resize_bilinear(xt, (12, 10)).numpy()
    for h in range(12):
        for w in range(10):
            sh, sw = (h + 0.5) / 2 - 0.5, (w + 0.5) / 2 - 0.5     # half-pixel centres
            h0, w0 = int(np.floor(sh)), int(np.floor(sw))
            fh, fw = sh - h0, sw - w0
            c = lambda a, n: min(max(a, 0), n - 1)
            ref = ((1 - fh) * (1 - fw) * x[0, c(h0, 6), c(w0, 5)] + (1 - fh) * fw * x[0, c(h0, 6), c(w0 + 1, 5)]
                   + fh * (1 - fw) * x[0, c(h0 + 1, 6), c(w0, 5)] + fh * fw * x[0, c(h0 + 1, 6), c(w0 + 1, 5)])
            np.testing.assert_allclose(up[0, h, w], ref, atol=1e-12)


@pytest.mark.parametrize("version", ["v1", "v2"])
def test_ncsn_forward_shapes_and_conditioning(version):
    cfg = NCSNConfig(version=version, H=16, W=8, ngf=8, num_classes=4, sigma1=1.0, sigmaL=0.01)
    p = init_ncsn_params(cfg, seed=0, mode="perturbed")
    sig = bo.get_sigmas(cfg.sigma1, cfg.sigmaL, cfg.num_classes, "logarithmic")
    o = no.NCSNOracle(cfg, p, sigmas=sig)
    x = np.random.default_rng(0).uniform(0, 1, (3, 16, 8, 1))
    s0 = o.score(x, np.zeros(3, np.int64))
    s3 = o.score(x, np.full(3, 3, np.int64))
    assert s0.shape == (3, 16, 8, 1) and torch.isfinite(s0).all()
    assert not torch.allclose(s0, s3)
    # per-sample independence (instance norms only): a batch of one gives the same answer
    s_single = o.score(x[1:2], np.zeros(1, np.int64))
    assert torch.allclose(s_single[0], s0[1], atol=1e-10)
    if version == "v2":
        # output / sigma[idx] is the only conditioning (score_network_v2.py:275-276)
        assert torch.allclose(s0 * float(sig[0]), s3 * float(sig[3]), rtol=1e-9, atol=1e-12)


def test_synthetic_patch_statistics_match_real_fixture():
    from audiosourcesep_b200 import synthetic
    real = np.load(os.path.join(G, "real_patches.npz"))
    syn = synthetic.mel_patches_db(16, seed=0)[..., 0]
    assert syn.shape == (16, 96, 64) and syn.min() >= -100 and syn.max() <= 20
    rn = synthetic.normalise(real["gt1"])
    sn = synthetic.normalise(syn)
    assert abs(sn.mean() - rn.mean()) < 0.15 and 0.08 < sn.std() < 0.25
    f = syn - syn.mean()
    ac_f = (f[:, 1:, :] * f[:, :-1, :]).mean() / (f * f).mean()
    ac_t = (f[:, :, 1:] * f[:, :, :-1]).mean() / (f * f).mean()
    assert 0.6 < ac_f < 0.9 and 0.9 < ac_t < 0.99
