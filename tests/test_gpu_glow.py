"""GPU parity tests (B200): CUDA path through the C ABI vs the CPU oracle on the same seeded inputs.

Tolerances are the ones BASELINE.json's north_star states: log_prob within 1e-3 nats/dim,
inverse(forward(x)) <= 1e-4, plus tight fp32-level checks for the exact (CUDA-core) mode.
"""
import math

import numpy as np
import pytest
import torch

from audiosourcesep_b200 import GlowConfig, synthetic
from audiosourcesep_b200.weights import init_glow_params
from oracle.glow_oracle import GlowOracle, squeeze as o_squeeze, inv1x1_weight, inv1x1_weight_inverse

pytestmark = pytest.mark.gpu

LOG2 = math.log(2.0)


def _glow(cfg, params, precision):
    from audiosourcesep_b200.glow import Glow
    return Glow(cfg, params, precision=precision)


def _prec(name):
    from audiosourcesep_b200 import _lib
    return {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16, "bf16x2": _lib.PREC_BF16X2,
            "fp16x2": _lib.PREC_FP16X2, "fp16x3": _lib.PREC_FP16X3}[name]


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


# ----------------------------------------------------------------- single bijectors (unittest_flow_models.py cases)
def test_squeeze_matches_reference_order_and_roundtrips():
    from audiosourcesep_b200 import ops
    x = torch.arange(2 * 4 * 6 * 3, dtype=torch.float32).reshape(2, 4, 6, 3)
    y = ops.squeeze(x)
    assert torch.equal(y.cpu(), o_squeeze(x.double()).float())
    assert torch.equal(ops.squeeze(y, inverse=True).cpu(), x)         # exact equality, as the reference test demands


def test_actnorm_known_answer_and_roundtrip():
    from audiosourcesep_b200 import ops
    # ActNorm initialised on a minibatch of 2s and 1s has scale exactly 2 (unittest_flow_models.py:149-154)
    x = torch.randn(1, 2, 2, 1)
    ls = torch.full((1,), LOG2)
    sh = torch.full((1,), -3.0)
    y = ops.actnorm(x, ls, sh)
    np.testing.assert_allclose(_np(y), _np(x) * 2 - 3, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(_np(ops.actnorm(y, ls, sh, inverse=True)), _np(x), rtol=1e-6, atol=1e-6)


def test_coupling_known_answer_4log2():
    from audiosourcesep_b200 import ops
    # toy network log_s = log 2, t = 1 on a (2,2,2) event: fldj = 4 log 2, fldj == -ildj (:141-146, :44-46)
    x = torch.randn(1, 2, 2, 2)
    raw = math.atanh(LOG2)
    r = torch.cat([torch.full((1, 2, 2, 1), raw), torch.ones(1, 2, 2, 1)], dim=-1)
    y, fldj = ops.coupling(x, r)
    assert fldj.item() == pytest.approx(4 * LOG2, rel=1e-6)
    np.testing.assert_allclose(_np(y[..., 0]), 2 * _np(x[..., 0]) + 1, rtol=1e-6, atol=1e-6)
    assert torch.equal(y[..., 1].cpu(), x[..., 1])
    xr, ildj = ops.coupling(y, r, inverse=True)
    assert ildj.item() == pytest.approx(-fldj.item(), rel=1e-6)
    np.testing.assert_allclose(_np(xr), _np(x), atol=1e-6)


def test_inv1x1_roundtrip():
    from audiosourcesep_b200 import ops
    cfg = GlowConfig(H=8, W=8, C=1, L=2, K=1, n_filters=64, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=3, mode="perturbed")
    args = [torch.as_tensor(p["b1/s0/inv1x1/" + n], dtype=torch.float64) for n in ("P", "L", "U", "log_S", "sign_S")]
    Wm, Wi = inv1x1_weight(*args), inv1x1_weight_inverse(*args)
    x = torch.randn(2, 2, 2, 8)
    y = ops.inv1x1(x, Wm.float())
    np.testing.assert_allclose(_np(y), _np(x) @ Wm.numpy(), atol=1e-5)
    np.testing.assert_allclose(_np(ops.inv1x1(y, Wi.float())), _np(x), atol=1e-5)


# ----------------------------------------------------------------- whole model, exact (fp32) mode, small shapes
@pytest.mark.parametrize("L,H,W", [(2, 8, 8), (3, 16, 8), (4, 16, 16)])
def test_glow_fp32_matches_oracle(L, H, W):
    cfg = GlowConfig(H=H, W=W, C=1, L=L, K=3, n_filters=64, learntop=True, minval=-100.0, maxval=20.0)
    p = init_glow_params(cfg, seed=11, mode="perturbed")
    o = GlowOracle(cfg, p)
    m = _glow(cfg, p, _prec("fp32"))
    x = synthetic.mel_patches_db(3, seed=5, H=H, W=W)
    z_o, ld_o = o.forward(x)
    z, ld = m.forward_with_log_det(torch.as_tensor(x))
    np.testing.assert_allclose(_np(z), z_o.numpy(), atol=2e-4, rtol=2e-4)
    np.testing.assert_allclose(_np(ld), ld_o.numpy(), atol=2e-3, rtol=1e-5)
    lp = m.log_prob(torch.as_tensor(x))
    np.testing.assert_allclose(_np(lp), o.log_prob(x).numpy(), atol=5e-3, rtol=1e-5)
    # inverse(forward(x)) reconstruction <= 1e-4 in normalised units
    xr = m.inverse(z)
    assert np.max(np.abs(_np(xr) - x)) / 120.0 <= 1e-4
    np.testing.assert_allclose(_np(m.inverse(torch.as_tensor(z_o.numpy()))), x, atol=120 * 1e-4)
    # grad log p
    g_o, _ = o.grad_log_prob(x)
    g, lp2 = m.grad_log_prob(torch.as_tensor(x), return_log_prob=True)
    rel = np.linalg.norm(_np(g) - g_o.numpy()) / np.linalg.norm(g_o.numpy())
    assert rel < 1e-4, rel
    np.testing.assert_allclose(_np(lp2), _np(lp), atol=1e-3)
    # sample with injected latent draw
    eps = np.random.default_rng(0).standard_normal((2,) + cfg.latent_shape).astype(np.float32)
    xs = m.sample(2, eps=torch.as_tensor(eps))
    np.testing.assert_allclose(_np(xs), o.sample_from_latent(eps).numpy(), atol=120 * 2e-4)


def test_glow_no_learntop_and_empty_batch():
    cfg = GlowConfig(H=8, W=8, C=1, L=2, K=2, n_filters=64, learntop=False, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=1, mode="perturbed")
    o = GlowOracle(cfg, p)
    m = _glow(cfg, p, _prec("fp32"))
    x = np.random.default_rng(0).uniform(0, 1, (2, 8, 8, 1)).astype(np.float32)
    np.testing.assert_allclose(_np(m.log_prob(torch.as_tensor(x))), o.log_prob(x).numpy(), atol=2e-3)
    empty = torch.empty((0, 8, 8, 1))
    assert m.log_prob(empty).shape == (0,)


def test_init_actnorm_matches_oracle_including_quirk():
    cfg = GlowConfig(H=16, W=8, C=1, L=3, K=2, n_filters=64, minval=-100.0, maxval=20.0)
    p = init_glow_params(cfg, seed=4, mode="perturbed")
    mb = synthetic.mel_patches_db(6, seed=9, H=16, W=8)
    o = GlowOracle(cfg, p)
    new = o.init_actnorm(mb)
    m = _glow(cfg, p, _prec("fp32"))
    m.init_actnorm(torch.as_tensor(mb))
    for name, val in new.items():
        np.testing.assert_allclose(m.get_param(name), val, rtol=2e-4, atol=2e-4, err_msg=name)
    x = synthetic.mel_patches_db(2, seed=1, H=16, W=8)
    np.testing.assert_allclose(_np(m.log_prob(torch.as_tensor(x))), o.log_prob(x).numpy(), rtol=1e-4, atol=2e-2)


def test_shape_contract_errors():
    from audiosourcesep_b200._lib import AsepError
    cfg = GlowConfig(H=8, W=8, C=1, L=2, K=1, n_filters=64, minval=0.0, maxval=1.0)
    m = _glow(cfg, init_glow_params(cfg, seed=0), _prec("fp32"))
    with pytest.raises(AsepError):
        m.log_prob(torch.zeros(2, 8, 4, 1))
    with pytest.raises(AsepError):
        m.set_param("no/such/param", np.zeros(3, np.float32))
    with pytest.raises(AsepError):
        m.set_param("b0/s0/actnorm/shift", np.zeros(3, np.float32))


# ----------------------------------------------------------------- tcgen05 coupling network vs the fp32 kernels (on device)
@pytest.mark.parametrize("mode", ["bf16", "bf16x2", "fp16x2", "fp16x3"])
@pytest.mark.parametrize("block", [0, 1, 2])
def test_tc_coupling_nn_matches_fp32(block, mode):
    """tcgen05 kernel vs the CUDA-core fp32 kernels on the device, forward and data gradient.

    A ReLU network's gradient is piecewise constant in the pre-activations, so wherever bf16 rounding
    flips the sign of a pre-activation near zero the two gradients differ by that unit's whole
    contribution (measured: ~0.2% of units flip -> ~4.5% relative L2 difference).  The strict gradient
    check therefore uses biases that keep every ReLU active; the generic weights get a loose bound.
    """
    from audiosourcesep_b200 import ops
    cfg = GlowConfig(H=32, W=16, C=1, L=3, K=2, n_filters=512, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=21, mode="perturbed")
    for b in range(3):       # step 1 of every block: ReLUs always active
        p[f"b{b}/s1/nn/conv1/bias"] = np.full(512, 6.0, np.float32)
        p[f"b{b}/s1/nn/conv2/bias"] = np.full(512, 40.0, np.float32)
    m32 = _glow(cfg, p, _prec("fp32"))
    mtc = _glow(cfg, p, _prec(mode))
    Hb, Wb, Cb = cfg.level_shape(block)
    N = 5   # 5*16*8 = 640 pixels at block 0 -> 5 tiles; block 2: 5*4*2 = 40 pixels -> ragged single tile
    g = torch.Generator().manual_seed(block)
    state = torch.randn(N, Hb, Wb, Cb, generator=g) * 0.5
    gr = torch.randn(N, Hb, Wb, Cb, generator=g)
    # the split-precision modes differ from fp32 only by the 16-bit WEIGHT rounding (bf16 2^-9, fp16 2^-12)
    # (stage-1 / conv1 weights stay bf16 in the one- and two-product modes); fp16x3 carries 22-bit weights AND activations
    fwd_tol = {"bf16": 5e-3, "bf16x2": 4e-3, "fp16x2": 3e-3, "fp16x3": 1e-5}[mode]
    for step, bwd_tol in ((0, 8e-2), (1, 1e-2)):
        if mode == "fp16x3":
            bwd_tol = 2e-3 if step == 0 else 1e-4      # fp32-level masks; bf16-pair (2^-17) products in the gradient pass
        r32 = _np(m32.coupling_nn(block, step, state))
        rtc = _np(mtc.coupling_nn(block, step, state))
        scale = np.abs(r32).max()
        assert np.abs(rtc - r32).max() <= 2e-2 * scale, (np.abs(rtc - r32).max(), scale)
        rel_f = np.linalg.norm(rtc - r32) / np.linalg.norm(r32)
        assert rel_f < fwd_tol, (mode, rel_f)
        b32 = _np(m32.coupling_nn_backward(block, step, state, gr))
        btc = _np(mtc.coupling_nn_backward(block, step, state, gr))
        rel = np.linalg.norm(btc - b32) / np.linalg.norm(b32)
        print(f"[{mode} block {block} step {step}] forward rel {rel_f:.2e}, data-gradient rel {rel:.2e}")
        assert rel < bwd_tol, (step, rel)
        # deterministic: the same input gives bit-identical output
        assert np.array_equal(rtc, _np(mtc.coupling_nn(block, step, state)))


def test_coupling_nn_fp32_matches_oracle():
    cfg = GlowConfig(H=16, W=8, C=1, L=3, K=1, n_filters=64, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=2, mode="perturbed")
    o = GlowOracle(cfg, p)
    m = _glow(cfg, p, _prec("fp32"))
    for block in range(3):
        Hb, Wb, Cb = cfg.level_shape(block)
        state = torch.randn(2, Hb, Wb, Cb, generator=torch.Generator().manual_seed(block))
        xb = state[..., Cb // 2:].double().clone().requires_grad_(True)
        pre = f"b{block}/s0/"
        ls, t = o.nn(xb, pre)
        raw = torch.atanh(ls)
        r = _np(m.coupling_nn(block, 0, state))
        np.testing.assert_allclose(r[..., : Cb // 2], raw.detach().numpy(), atol=2e-4, rtol=2e-4)
        np.testing.assert_allclose(r[..., Cb // 2:], t.detach().numpy(), atol=2e-4, rtol=2e-4)
        gr = torch.randn(2, Hb, Wb, Cb, generator=torch.Generator().manual_seed(7))
        (torch.cat([raw, t], -1) * gr.double()).sum().backward()
        gx = _np(m.coupling_nn_backward(block, 0, state, gr))
        np.testing.assert_allclose(gx, xb.grad.numpy(), atol=2e-4, rtol=2e-3)


# ----------------------------------------------------------------- config-shape model (96x64, 512 filters)
@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16", "bf16x2", "fp16x2", "fp16x3"])
def test_glow_config_shape_vs_oracle(precision):
    cfg = GlowConfig(H=96, W=64, C=1, L=3, K=4, n_filters=512, minval=-100.0, maxval=20.0)
    p = init_glow_params(cfg, seed=2, mode="perturbed")
    o = GlowOracle(cfg, p)
    m = _glow(cfg, p, _prec(precision))
    x = synthetic.mel_patches_db(2, seed=0)
    D = cfg.dims
    lp_o = o.log_prob(x).numpy()
    lp = _np(m.log_prob(torch.as_tensor(x)))
    assert np.max(np.abs(lp - lp_o)) / D <= 1e-3, (lp, lp_o)          # nats/dim gate of the north star
    if precision == "fp32":
        assert np.max(np.abs(lp - lp_o)) <= 0.05
    z, _ = m.forward_with_log_det(torch.as_tensor(x))
    xr = _np(m.inverse(z))
    rt = np.max(np.abs(xr - x)) / 120.0
    print(f"[{precision}] round trip max-abs (normalised) = {rt:.3e}; |dlogp| max = {np.max(np.abs(lp - lp_o)):.4f} nats")
    # round-trip gate (<= 1e-4) holds in the exact mode.  With bf16 hidden activations the coupling
    # network is piecewise constant at the 2^-9 level, so inverse() -- which re-evaluates it on inputs
    # that differ from forward()'s by fp32 round-off -- reconstructs only to ~1e-2 (documented limit).
    # ASEP_PREC_FP16 (fp16 hidden activations, 2^-12) tightens it ~8x but still misses the gate.
    # The split-precision tensor-core modes (hidden activations as hi + lo pairs) meet the gate.
    assert rt <= {"fp32": 1e-4, "bf16": 3e-2, "fp16": 4e-3, "bf16x2": 1e-4, "fp16x2": 1e-4, "fp16x3": 1e-4}[precision], rt
    if precision == "fp16":
        assert np.max(np.abs(lp - lp_o)) / D <= 2e-4
    g_o, _ = o.grad_log_prob(x)
    g = _np(m.grad_log_prob(torch.as_tensor(x)))
    rel = np.linalg.norm(g - g_o.numpy()) / np.linalg.norm(g_o.numpy())
    print(f"[{precision}] grad_log_prob relative L2 error = {rel:.3e}")
    # ReLU sign flips make even fp32-vs-fp64 gradients differ at the 1e-4..1e-3 level
    # 16-bit WEIGHTS (bf16 in every tensor-core mode for conv1 and for the whole data-gradient pass) bound the score error
    # of the one- and two-product modes; the three-product mode (hi + lo weights as well) is the score-exact one
    assert rel < {"fp32": 2e-3, "bf16": 8e-2, "fp16": 8e-2, "bf16x2": 8e-2, "fp16x2": 8e-2, "fp16x3": 3e-3}[precision], rel


def test_glow_full_depth_bf16_vs_oracle():
    """The melspec_glow.yml model (L=3, K=40, 512 filters): log_prob gate at full depth, on synthetic
    patches and on real mel patches from the reference's shipped results.npz."""
    import os
    cfg = GlowConfig()
    p = init_glow_params(cfg, seed=2, mode="perturbed")
    x = synthetic.mel_patches_db(2, seed=3)
    o = GlowOracle(cfg, p, dtype=torch.float32)
    lp_o = o.log_prob(x).double().numpy()
    assert np.all(np.isfinite(lp_o))
    m = _glow(cfg, p, _prec("bf16"))
    lp = _np(m.log_prob(torch.as_tensor(x)))
    print("full depth: logp cuda", lp, "oracle", lp_o, "nats/dim err", np.max(np.abs(lp - lp_o)) / cfg.dims)
    assert np.max(np.abs(lp - lp_o)) / cfg.dims <= 1e-3, (lp, lp_o)
    real = np.load(os.path.join(os.path.dirname(__file__), "golden", "real_patches.npz"))["gt1"][:2, :, :, None]
    lp_r = _np(m.log_prob(torch.as_tensor(real)))
    lp_ro = o.log_prob(real).double().numpy()
    print("real patches: logp cuda", lp_r, "oracle", lp_ro)
    assert np.max(np.abs(lp_r - lp_ro)) / cfg.dims <= 1e-3
    m32 = _glow(cfg, p, _prec("fp32"))
    z, _ = m32.forward_with_log_det(torch.as_tensor(x))
    rt = np.max(np.abs(_np(m32.inverse(z)) - x)) / 120.0
    print("full depth fp32 round trip", rt)
    assert rt <= 1e-4, rt


@pytest.mark.parametrize("precision", ["bf16x2", "fp16x2", "fp16x3"])
def test_glow_full_depth_tensor_core_round_trip_gate(precision):
    """inverse(forward(x)) <= 1e-4 (normalised units) ON TENSOR CORES at the depth of every config (L=3, K=40, 512
    filters; flow_glow.py:187-196), at n_mixed = 30 patches (run_basis_sep.py:478) and on a ragged batch; log_prob
    against the oracle at the same depth; sample() against the fp32 CUDA-core path."""
    cfg = GlowConfig()
    p = init_glow_params(cfg, seed=2, mode="perturbed")
    m = _glow(cfg, p, _prec(precision))
    x = synthetic.mel_patches_db(30, seed=3)
    xt = torch.as_tensor(x)
    z, ld = m.forward_with_log_det(xt)
    xr = _np(m.inverse(z))
    rt = np.max(np.abs(xr - x)) / 120.0
    print(f"[{precision}] K=40 round trip max-abs (normalised) over 30 patches = {rt:.3e}")
    assert rt <= 1e-4, rt
    # the oracle at full depth on two of the patches
    o = GlowOracle(cfg, p, dtype=torch.float32)
    lp_o = o.log_prob(x[:2]).double().numpy()
    lp = _np(m.log_prob(xt[:2].contiguous()))
    print(f"[{precision}] K=40 log_prob err = {np.max(np.abs(lp - lp_o)) / cfg.dims:.3e} nats/dim")
    assert np.max(np.abs(lp - lp_o)) / cfg.dims <= 1e-3
    # latent -> data: tensor-core inverse vs the CUDA-core fp32 inverse on the same latent
    m32 = _glow(cfg, p, _prec("fp32"))
    x32 = _np(m32.inverse(z[:3].contiguous()))
    # (the two inverses use different weight roundings -- 16-bit vs fp32 -- and 120 inverse steps amplify that)
    assert np.max(np.abs(_np(m.inverse(z[:3].contiguous())) - x32)) / 120.0 <= (1e-3 if precision == "fp16x3" else 2e-2)
    # ragged single-sample batch gives the same bits as the batched call
    assert torch.equal(m.forward_with_log_det(xt[7:8].contiguous())[0], z[7:8])


# ----------------------------------------------------------------- size-independent properties at the bench size
@pytest.mark.parametrize("precision", ["bf16", "fp16", "bf16x2"])
def test_full_size_batch_is_sample_wise_and_order_invariant(precision):
    """At the benchmark's batch size (2048 patches, full-depth model) the oracle is too slow to compare against, so
    the check is a property the reference has by construction: log_prob and grad_log_prob are per-sample maps
    (flow_builder.py:127-141) -- the value of a patch may not depend on its position in the batch, on its neighbours
    in a 128-pixel tile, or on how many tiles every CTA walks.  Bit-exact: a row of an MMA does not see the others."""
    cfg = GlowConfig()
    p = init_glow_params(cfg, seed=2, mode="perturbed")
    m = _glow(cfg, p, _prec(precision))
    base = synthetic.mel_patches_db(64, seed=11)
    x = np.concatenate([np.roll(base, 3 * i, axis=2) for i in range(32)], axis=0)          # 2048 distinct patches
    xt = torch.as_tensor(x).cuda()
    lp = m.log_prob(xt)
    assert torch.isfinite(lp).all()
    perm = torch.randperm(x.shape[0], generator=torch.Generator().manual_seed(0)).cuda()
    lp_perm = m.log_prob(xt[perm].contiguous())
    assert torch.equal(lp_perm, lp[perm])
    for sl in (slice(0, 3), slice(1000, 1007), slice(2041, 2048)):                         # small batches: ragged tiles
        assert torch.equal(m.log_prob(xt[sl].contiguous()), lp[sl])
    g = m.grad_log_prob(xt[:256].contiguous())
    g_small = m.grad_log_prob(xt[100:105].contiguous())
    assert torch.equal(g_small, g[100:105])
    # duplicated patches (the cyclic shifts wrap after 64/3 steps) get identical values
    dup = m.log_prob(torch.cat([xt[:5], xt[:5]], 0))
    assert torch.equal(dup[:5], dup[5:])


def test_small_batch_graph_replay_equals_eager_launches():
    """At N <= 128 log_prob / grad_log_prob / inverse replay one captured CUDA graph per (direction, N) from the third
    call on (first: eager, second: capture + replay).  Replays must be bit-identical to eager launches, follow new
    inputs (staging copies), survive an interleaved call that re-carves the workspace, and count their launches."""
    from audiosourcesep_b200 import _lib
    from audiosourcesep_b200.glow import Glow
    cfg = GlowConfig(H=96, W=64, C=1, L=3, K=3, n_filters=512)
    params = init_glow_params(cfg, seed=2, mode="perturbed")
    m = Glow(cfg, params, precision=_lib.PREC_BF16)
    xa = torch.as_tensor(synthetic.mel_patches_db(5, seed=0)).cuda()
    xb = torch.as_tensor(synthetic.mel_patches_db(5, seed=1)).cuda()
    ref = Glow(cfg, params, precision=_lib.PREC_BF16)        # every call below is this handle's FIRST of its kind: eager
    want = {"lp_a": ref.log_prob(xa), "g_b": ref.grad_log_prob(xb)}
    want["z_b"] = ref.forward(xb)
    want["x_b"] = ref.inverse(want["z_b"])
    n0 = _lib.launch_count()
    first = m.log_prob(xb)                                   # eager
    eager_launches = _lib.launch_count() - n0
    m.log_prob(xb)                                           # capture + first replay
    n0 = _lib.launch_count()
    got = m.log_prob(xa)                                     # replay with a new input
    assert _lib.launch_count() - n0 == eager_launches
    assert torch.equal(got, want["lp_a"]) and not torch.equal(got, first)
    for _ in range(3):                                       # grad: the first call re-carves the workspace (save = true)
        g = m.grad_log_prob(xb)
    assert torch.equal(g, want["g_b"])
    for _ in range(3):                                       # log_prob again: its graph was dropped by the re-carve
        got = m.log_prob(xa)
    assert torch.equal(got, want["lp_a"])
    for _ in range(3):
        xr = m.inverse(want["z_b"])
    assert torch.equal(xr, want["x_b"])
    assert torch.equal(m.grad_log_prob(xb), want["g_b"])     # the gradient graph is still valid after the other directions
