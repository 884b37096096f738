"""GPU parity tests of the Glow training step (loss, every parameter gradient, Adamax) vs the oracle."""
import numpy as np
import pytest
import torch

from audiosourcesep_b200 import GlowConfig
from audiosourcesep_b200.weights import init_glow_params, is_trainable
from oracle import train_oracle as to

pytestmark = pytest.mark.gpu


def _model(cfg, p):
    from audiosourcesep_b200 import _lib
    from audiosourcesep_b200.glow import Glow
    m = Glow(cfg, p, precision=_lib.PREC_FP32)
    m.enable_training()
    return m


def _split(m, flat):
    flat = flat.detach().cpu().numpy()
    return {name: flat[off:off + int(np.prod(shape))].reshape(shape) for name, off, shape in m.trainable_layout()}


@pytest.mark.parametrize("noisy", [False, True])
def test_train_grads_match_oracle(noisy):
    cfg = GlowConfig(H=16, W=8, C=1, L=3, K=2, n_filters=64, minval=-100.0, maxval=20.0)
    p = init_glow_params(cfg, seed=4, mode="perturbed")
    rng = np.random.default_rng(0)
    x = rng.uniform(-90, 10, (3, 16, 8, 1)).astype(np.float32)
    noise = rng.standard_normal(x.shape).astype(np.float32) if noisy else None
    sigma = 0.6 if noisy else 0.0
    m = _model(cfg, p)
    assert m.num_trainable == sum(int(np.prod(s)) for n, s in m._shapes.items() if is_trainable(n))
    g, loss = m.train_grads(torch.as_tensor(x), global_batch=6, noise=None if noise is None else torch.as_tensor(noise),
                            sigma=sigma)
    loss_o, g_o = to.loss_and_grads(cfg, p, x, 6, noise=noise, sigma=sigma)
    assert abs(float(loss.item()) - loss_o) <= 1e-5 * abs(loss_o), (loss.item(), loss_o)
    got = _split(m, g)
    worst = 0.0
    for name, want in g_o.items():
        want = to.mask_structural(name, want)
        a = got[name].astype(np.float64)
        denom = max(np.linalg.norm(want), 1e-6 * np.sqrt(want.size))
        rel = np.linalg.norm(a - want) / denom
        worst = max(worst, rel)
        assert rel <= 2e-3, (name, rel, np.abs(a - want).max(), np.abs(want).max())
    print(f"[noisy={noisy}] worst relative gradient error over {len(g_o)} tensors = {worst:.3e}")


def test_data_parallel_shards_sum_to_the_full_batch_gradient():
    """Every rank passes the GLOBAL batch size, so summing the per-shard gradients (what the NCCL all-reduce does)
    gives the full-batch gradient (train_glow.py:31 compute_average_loss + MirroredStrategy SUM)."""
    cfg = GlowConfig(H=16, W=8, C=1, L=3, K=1, n_filters=64, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=5, mode="perturbed")
    x = torch.as_tensor(np.random.default_rng(1).uniform(0, 1, (4, 16, 8, 1)).astype(np.float32))
    m = _model(cfg, p)
    g_full, l_full = m.train_grads(x, global_batch=4)
    g_a, l_a = m.train_grads(x[:1], global_batch=4)
    g_b, l_b = m.train_grads(x[1:], global_batch=4)
    rel = float(torch.linalg.norm(g_a + g_b - g_full) / torch.linalg.norm(g_full))
    assert rel <= 1e-5, rel
    assert abs(float(l_a + l_b - l_full)) <= 1e-5 * abs(float(l_full))


def test_adamax_steps_match_oracle_and_refresh_constants():
    cfg = GlowConfig(H=16, W=8, C=1, L=3, K=1, n_filters=64, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=6, mode="perturbed")
    x = np.random.default_rng(2).uniform(0, 1, (2, 16, 8, 1)).astype(np.float32)
    m = _model(cfg, p)
    layout = m.trainable_layout()
    theta = m.get_flat().cpu().numpy()
    for name, off, shape in layout:                       # flat order = construction order of the trainables
        np.testing.assert_array_equal(theta[off:off + int(np.prod(shape))].reshape(shape), p[name])
    mm, uu = np.zeros_like(theta), np.zeros_like(theta)
    cur = dict(p)
    losses = []
    for t in (1, 2, 3):
        g, loss = m.train_grads(torch.as_tensor(x), global_batch=2)
        losses.append(float(loss.item()))
        m.adamax_step(g, lr=1e-3)
        # oracle: gradient at the oracle's own current parameters, Keras Adamax in float32
        _, g_o = to.loss_and_grads(cfg, cur, x, 2)
        flat_g = np.concatenate([to.mask_structural(n, g_o[n]).ravel() for n, _, _ in layout]).astype(np.float32)
        theta, mm, uu = to.adamax_update(theta, flat_g, mm, uu, t)
        for name, off, shape in layout:
            cur[name] = theta[off:off + int(np.prod(shape))].reshape(shape).copy()
        got = m.get_flat().cpu().numpy()
        # Adamax moves every coordinate by ~lr regardless of |g|, so coordinates whose tiny gradient differs in sign
        # between fp32 and fp64 move the other way: compare where the oracle gradient is not negligible
        big = np.abs(flat_g) > 1e-4 * np.abs(flat_g).max()
        assert np.abs(got - theta)[big].max() <= 2e-4, (t, np.abs(got - theta)[big].max())
    assert losses[2] < losses[0]                           # three steps on one batch reduce its loss
    # the refreshed device constants and a host re-prepare agree: log_prob before / after sync_host
    lp_dev = -2.0 * m.train_grads(torch.as_tensor(x), global_batch=2)[1].item()
    m.sync_host()
    lp_host = float(m.log_prob(torch.as_tensor(x)).sum().item())
    assert abs(lp_dev - lp_host) <= 1e-4 * abs(lp_host), (lp_dev, lp_host)


def test_train_loop_host_mirror_reduces_loss_on_synthetic_patches():
    """The train_glow mirror end to end on one GPU: data-dependent ActNorm init, 6 Adamax steps, loss goes down."""
    import argparse
    from audiosourcesep_b200 import train_glow as tg
    from audiosourcesep_b200.flow_models.flow_builder import build_glow
    data = tg.synthetic_dataset(24, 0, 32, 16)
    flow = build_glow(data[:8], [32, 16, 1], L=3, K=2, n_filters=64, learntop=True, data_type="melspec",
                      minval=-100.0, maxval=20.0, use_logit=False, seed=1)
    flow.enable_training()
    args = argparse.Namespace(batch_size=8, n_epochs=2, learning_rate=1e-3, optimizer="adamax", seed=0)
    hist = tg.train(flow, tg.setUp_optimizer(None, args), data, args, log=lambda *a: None)
    assert len(hist) == 6 and np.all(np.isfinite(hist)) and hist[-1] < hist[0]
    hist_noisy = tg.train(flow, tg.setUp_optimizer(None, args), data, args, sigma=0.5, log=lambda *a: None)
    assert len(hist_noisy) == 6 and np.all(np.isfinite(hist_noisy))


def test_train_grads_tensor_core_path_vs_oracle():
    """tcgen05 training path (bf16 forward / backward / weight-gradient GEMMs): every gradient tensor within the
    bf16-operand tolerance of the fp64 oracle; the loss within 1e-3 nats/dim."""
    from audiosourcesep_b200 import _lib
    from audiosourcesep_b200.glow import Glow
    cfg = GlowConfig(H=32, W=16, C=1, L=3, K=2, n_filters=512, minval=-100.0, maxval=20.0)
    p = init_glow_params(cfg, seed=4, mode="perturbed")
    rng = np.random.default_rng(0)
    x = rng.uniform(-90, 10, (5, 32, 16, 1)).astype(np.float32)
    m = Glow(cfg, p, precision=_lib.PREC_BF16)
    m.enable_training()
    g, loss = m.train_grads(torch.as_tensor(x), global_batch=5)
    loss_o, g_o = to.loss_and_grads(cfg, p, x, 5)
    assert abs(float(loss.item()) - loss_o) <= 1e-3 * cfg.dims, (loss.item(), loss_o)
    got = _split(m, g)
    worst, worst_name = 0.0, ""
    for name, want in g_o.items():
        want = to.mask_structural(name, want)
        a = got[name].astype(np.float64)
        denom = max(np.linalg.norm(want), 1e-6 * np.sqrt(want.size))
        rel = np.linalg.norm(a - want) / denom
        if rel > worst:
            worst, worst_name = rel, name
        assert rel <= 6e-2, (name, rel, np.abs(a - want).max(), np.abs(want).max())
    print(f"[tcgen05 training] worst relative gradient error = {worst:.3e} ({worst_name})")


def test_tensor_core_training_refreshes_tile_images_on_device():
    """After Adamax steps the bf16 tile images rebuilt on the device equal the ones the host would build."""
    from audiosourcesep_b200 import _lib
    from audiosourcesep_b200.glow import Glow
    cfg = GlowConfig(H=32, W=16, C=1, L=3, K=2, n_filters=512, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=6, mode="perturbed")
    x = torch.as_tensor(np.random.default_rng(2).uniform(0, 1, (4, 32, 16, 1)).astype(np.float32))
    m = Glow(cfg, p, precision=_lib.PREC_BF16)
    m.enable_training()
    losses = []
    for _ in range(3):
        g, loss = m.train_grads(x, global_batch=4)
        losses.append(float(loss.item()))
        m.adamax_step(g, lr=2e-5)                   # small steps: this random 600k-parameter model overshoots at 1e-3
    assert np.all(np.isfinite(losses)) and losses[2] < losses[0], losses
    lp_dev = -4.0 * m.train_grads(x, global_batch=4)[1].item()
    m.sync_host()                                   # host re-prepare from the trained parameters
    lp_host = float(m.log_prob(x).sum().item())
    assert abs(lp_dev - lp_host) <= 1e-4 * abs(lp_host), (lp_dev, lp_host)


def test_graph_replay_of_the_gradient_pass_equals_the_eager_pass(monkeypatch):
    """asep_glow_train_grads runs eagerly once per batch size, then captures the pass into a CUDA graph (private
    stream + staging buffers) and replays it: every call must return the gradients of ITS inputs."""
    from audiosourcesep_b200 import _lib
    from audiosourcesep_b200.glow import Glow
    cfg = GlowConfig(H=32, W=16, C=1, L=3, K=2, n_filters=512, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=8, mode="perturbed")
    rng = np.random.default_rng(3)
    xs = [torch.as_tensor(rng.uniform(0, 1, (4, 32, 16, 1)).astype(np.float32)).cuda() for _ in range(3)]
    noise = torch.as_tensor(rng.standard_normal((4, 32, 16, 1)).astype(np.float32)).cuda()

    monkeypatch.setenv("ASEP_NO_GRAPH", "1")
    ref = Glow(cfg, p, precision=_lib.PREC_BF16)
    ref.enable_training()
    want = [tuple(t.clone() for t in ref.train_grads(x, global_batch=4)) for x in xs]
    want_noisy = tuple(t.clone() for t in ref.train_grads(xs[0], global_batch=4, noise=noise, sigma=0.05))
    monkeypatch.delenv("ASEP_NO_GRAPH")

    m = Glow(cfg, p, precision=_lib.PREC_BF16)
    m.enable_training()
    m.train_grads(xs[0], global_batch=4)                       # eager call that sizes the scratch
    for x, (g_ref, l_ref) in zip(xs, want):                    # capture, then two replays on different inputs
        g, loss = m.train_grads(x, global_batch=4)
        torch.cuda.synchronize()
        assert float(loss.item()) == pytest.approx(float(l_ref.item()), rel=1e-6)
        # weight-gradient GEMMs reduce with fp32 atomics (split-K): equal up to summation order
        assert torch.allclose(g, g_ref, rtol=1e-4, atol=1e-6 * float(g_ref.abs().max())), float((g - g_ref).abs().max())
    g, loss = m.train_grads(xs[0], global_batch=4, noise=noise, sigma=0.05)   # different key -> new capture
    assert float(loss.item()) == pytest.approx(float(want_noisy[1].item()), rel=1e-6)
    assert torch.allclose(g, want_noisy[0], rtol=1e-4, atol=1e-6 * float(want_noisy[0].abs().max()))


def test_adam_step_matches_keras_update():
    """``optimizer: adam`` (train_utils.py:27-28): the library update equals the Keras formula on the same gradient."""
    cfg = GlowConfig(H=16, W=8, C=1, L=3, K=1, n_filters=64, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=6, mode="perturbed")
    x = torch.as_tensor(np.random.default_rng(2).uniform(0, 1, (2, 16, 8, 1)).astype(np.float32))
    m = _model(cfg, p)
    theta = m.get_flat().cpu().numpy()
    mm, vv = np.zeros_like(theta), np.zeros_like(theta)
    for t in (1, 2, 3):
        g, _ = m.train_grads(x, global_batch=2)
        m.apply_gradients(g, dict(kind="adam", lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7))
        theta, mm, vv = to.adam_update(theta, g.cpu().numpy(), mm, vv, t)
        np.testing.assert_allclose(m.get_flat().cpu().numpy(), theta, rtol=1e-5, atol=2e-7)


def test_training_without_learnable_prior():
    """learntop=False (flow_builder.py:143-144): no prior parameters in the flat vector; gradients match the oracle."""
    cfg = GlowConfig(H=16, W=8, C=1, L=3, K=1, n_filters=64, learntop=False, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=9, mode="perturbed")
    x = np.random.default_rng(5).uniform(0, 1, (3, 16, 8, 1)).astype(np.float32)
    m = _model(cfg, p)
    assert not any(n.startswith("prior/") for n, _, _ in m.trainable_layout())
    g, loss = m.train_grads(torch.as_tensor(x), global_batch=3)
    loss_o, g_o = to.loss_and_grads(cfg, p, x, 3)
    assert abs(float(loss.item()) - loss_o) <= 1e-5 * abs(loss_o)
    got = _split(m, g)
    for name, want in g_o.items():
        want = to.mask_structural(name, want)
        rel = np.linalg.norm(got[name] - want) / max(np.linalg.norm(want), 1e-6 * np.sqrt(want.size))
        assert rel <= 2e-3, (name, rel)


def test_log_prob_between_optimizer_steps_uses_the_live_log_det_constant():
    """The reference evaluates flow.log_prob on validation data every epoch (train_glow.py:150-170) without any
    'sync' call: after an optimizer step the ActNorm / 1x1 log-det constant must be the updated one."""
    from audiosourcesep_b200 import _lib
    from audiosourcesep_b200.glow import Glow
    cfg = GlowConfig(H=32, W=16, C=1, L=3, K=2, n_filters=512, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=6, mode="perturbed")
    x = torch.as_tensor(np.random.default_rng(2).uniform(0, 1, (4, 32, 16, 1)).astype(np.float32))
    for prec in (_lib.PREC_BF16, _lib.PREC_FP32):
        m = Glow(cfg, p, precision=prec)
        m.enable_training()
        lp0 = float(m.log_prob(x).sum().item())
        for _ in range(2):
            g, _ = m.train_grads(x, global_batch=4)
            m.adamax_step(g, lr=1e-3)
        lp_live = m.log_prob(x).cpu().numpy()                # no sync_host() in between
        fldj_live = m.forward_log_det_jacobian(x).cpu().numpy()
        # the training loss of the same batch uses the device-resident constants: -loss * global_batch = sum log_prob
        lp_train = -4.0 * float(m.train_grads(x, global_batch=4)[1].item())
        assert abs(float(lp_live.sum()) - lp_train) <= 2e-6 * abs(lp_train), (float(lp_live.sum()), lp_train)
        assert abs(lp_train - lp0) > 1e-3 * abs(lp0)         # the two steps did move the density
        m.sync_host()                                        # host re-derivation in double: agrees to fp32 round-off
        np.testing.assert_allclose(lp_live, m.log_prob(x).cpu().numpy(), rtol=2e-4)
        np.testing.assert_allclose(fldj_live, m.forward_log_det_jacobian(x).cpu().numpy(), rtol=2e-4)


def test_train_graph_survives_workspace_and_weight_reallocation():
    """train -> log_prob with a LARGER batch (re-carves the arena) -> sync_host (re-allocates the per-step constants
    and tile images) -> train again with the same graph key: the captured graph must not be replayed on the freed
    buffers (reference loop: per-epoch validation, sampling and checkpointing between train steps, train_glow.py:106-179)."""
    from audiosourcesep_b200 import _lib
    from audiosourcesep_b200.glow import Glow
    cfg = GlowConfig(H=32, W=16, C=1, L=3, K=2, n_filters=512, minval=0.0, maxval=1.0)
    p = init_glow_params(cfg, seed=8, mode="perturbed")
    rng = np.random.default_rng(3)
    x = torch.as_tensor(rng.uniform(0, 1, (4, 32, 16, 1)).astype(np.float32)).cuda()
    big = torch.as_tensor(rng.uniform(0, 1, (37, 32, 16, 1)).astype(np.float32)).cuda()
    m = Glow(cfg, p, precision=_lib.PREC_BF16)
    m.enable_training()
    g0, l0 = (t.clone() for t in m.train_grads(x, global_batch=4))      # eager
    g1, l1 = (t.clone() for t in m.train_grads(x, global_batch=4))      # capture + launch
    g2, l2 = (t.clone() for t in m.train_grads(x, global_batch=4))      # replay
    assert torch.allclose(g1, g0, rtol=1e-4, atol=1e-6 * float(g0.abs().max()))
    assert torch.allclose(g2, g0, rtol=1e-4, atol=1e-6 * float(g0.abs().max()))
    lp_big = m.log_prob(big)                                            # larger N: arena re-carved
    assert torch.isfinite(lp_big).all()
    g3, l3 = (t.clone() for t in m.train_grads(x, global_batch=4))
    assert torch.allclose(g3, g0, rtol=1e-4, atol=1e-6 * float(g0.abs().max())), float((g3 - g0).abs().max())
    m.train_grads(x, global_batch=4)                                    # re-captured
    m.sync_host()                                                       # constants and tile images re-allocated
    _ = m.sample(5)                                                     # inverse pass in between, another workspace shape
    # (sync_host re-derives the folded weights on the host in double, enable_training derived them on the device in
    #  fp32: a few bf16 tile-image entries differ in their last bit, hence the looser bound than for a pure replay)
    g_prev = None
    for _ in range(3):
        g4, l4 = m.train_grads(x, global_batch=4)
        assert float(l4.item()) == pytest.approx(float(l0.item()), rel=1e-5)
        rel = float(torch.linalg.norm(g4 - g0) / torch.linalg.norm(g0))
        assert rel <= 2e-3, rel
        if g_prev is not None:                                          # eager, capture and replay agree with each other
            assert torch.allclose(g4, g_prev, rtol=1e-4, atol=1e-6 * float(g0.abs().max()))
        g_prev = g4.clone()

