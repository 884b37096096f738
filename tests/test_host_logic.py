"""CPU tests of the host-side logic: config precedence, segment sharding and the final gather (gloo, world
size 2), the reference-module mirror, the results contract helpers."""
import argparse
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_module_paths_import():
    import audiosourcesep_b200.flow_models as fm
    from audiosourcesep_b200.flow_models import flow_builder, flow_glow, flow_tfp_bijectors
    for name in ("ActNorm", "Invertible1x1Conv", "AffineCouplingLayerSplit", "Squeeze", "SpecPreprocessing"):
        assert hasattr(flow_tfp_bijectors, name)
    for name in ("GlowStep", "GlowBlock", "GlowBijector_2blocks", "GlowBijector_3blocks", "GlowBijector_4blocks"):
        assert hasattr(flow_glow, name)
    assert callable(flow_builder.build_glow)
    from audiosourcesep_b200.ncsn import utils
    assert callable(utils.get_uncompiled_model) and callable(utils.get_uncompiled_model_v2) and callable(utils.get_sigmas)
    from audiosourcesep_b200 import run_basis_sep as rbs
    for name in ("compute_grad_logprob", "mixing_process", "basis_inner_loop", "basis_outer_loop", "main", "post_processing_fn"):
        assert callable(getattr(rbs, name))
    assert fm is not None
    # the remaining module paths the reference's scripts import
    from audiosourcesep_b200.flow_models import flow_tfk_layers
    assert flow_tfk_layers.ShiftAndLogScaleConvNet is flow_glow.ShiftAndLogScaleConvNet      # flow_tfk_layers.py:31
    from audiosourcesep_b200 import train_noisy_glow, train_utils
    for name in ("setUp_optimizer", "setUp_checkpoint", "get_config", "dict2namespace"):     # train_utils.py:23,62,114,123
        assert callable(getattr(train_utils, name))
    assert callable(train_noisy_glow.main) and train_noisy_glow.build_parser().parse_args([]).noisy is True
    # train_ncsn.py, datasets/data_loader.py, melspec_inversion_basis.py, bsseval_v4.py, oracle_systems.py
    from audiosourcesep_b200 import bsseval_v4, melspec_inversion_basis, oracle_systems, train_ncsn
    from audiosourcesep_b200.datasets import data_loader
    for mod, names in ((train_ncsn, ("train", "distributed_train_step", "get_noise_conditionned_data", "main")),
                       (data_loader, ("get_song_extract", "load_wav")),
                       (melspec_inversion_basis, ("stft_inversion_fn", "single_channel_wiener_filter", "main")),
                       (bsseval_v4, ("bss_eval", "bss_eval_sources", "bss_eval_sources_framewise", "bss_eval_images",
                                     "bss_eval_images_framewise", "validate", "Framing")),
                       (oracle_systems, ("IBM_melspec", "IRM_melspec"))):
        for name in names:
            assert callable(getattr(mod, name)), (mod.__name__, name)
    # bsseval_v4.Framing restates the reference's window arithmetic (bsseval_v4.py:377-418)
    assert [(w.start, w.stop) for w in bsseval_v4.Framing(8000, 6000, 20000)] == [(0, 8000), (6000, 14000), (12000, 20000)]
    assert [(w.start, w.stop) for w in bsseval_v4.Framing(np.inf, np.inf, 123)] == [(0, 123)]


def test_build_glow_argument_errors():
    from audiosourcesep_b200.flow_models.flow_builder import build_glow
    with pytest.raises(ValueError, match="L should be 2, 3 or 4"):          # flow_builder.py:77-78
        build_glow(None, [96, 64, 1], L=5)
    # the reference's SpecPreprocessing defaults to use_logit=True: no silent default here
    with pytest.raises(ValueError, match="use_logit"):
        build_glow(None, [96, 64, 1], L=3, data_type="melspec")
    with pytest.raises(NotImplementedError):
        build_glow(None, [96, 64, 1], L=3, data_type="melspec", use_logit=True)


def test_set_up_optimizer_offers_adam_and_adamax():
    from audiosourcesep_b200.train_utils import setUp_optimizer
    for kind in ("adam", "adamax"):                                           # train_utils.py:26-32
        opt = setUp_optimizer(None, argparse.Namespace(optimizer=kind, learning_rate=2e-3))
        assert opt == dict(kind=kind, lr=2e-3, beta1=0.9, beta2=0.999, eps=1e-7)
    with pytest.raises(ValueError, match="adam or adamax"):
        setUp_optimizer(None, argparse.Namespace(optimizer="sgd", learning_rate=1e-3))


def test_train_rejects_dataset_smaller_than_a_batch():
    from audiosourcesep_b200.train_glow import train
    args = argparse.Namespace(batch_size=8, n_epochs=1, learning_rate=1e-3, optimizer="adamax", seed=0)
    with pytest.raises(ValueError, match="fewer than one global batch"):
        train(object(), {}, np.zeros((4, 8, 8, 1), np.float32), args)


def test_checkpoint_manager_keeps_max_to_keep(tmp_path):
    from audiosourcesep_b200.train_utils import setUp_checkpoint

    class FakeModel:
        def __init__(self):
            self.variables = {"w": np.arange(3, dtype=np.float32)}
            self.restored = None

        def set_params(self, p):
            self.restored = p

        def prepare(self):
            pass

    m = FakeModel()
    ckpt, manager = setUp_checkpoint(None, m, None, max_to_keep=2, path=str(tmp_path / "tf_ckpts"))
    assert manager.latest_checkpoint is None and manager.restore() is None
    for i in range(4):
        m.variables = {"w": np.full(3, i, np.float32)}
        manager.save()
    files = sorted(os.listdir(tmp_path / "tf_ckpts"))
    assert files == ["ckpt-3.npz", "ckpt-4.npz"]
    assert manager.restore().endswith("ckpt-4.npz") and np.array_equal(m.restored["w"], np.full(3, 3, np.float32))


@pytest.mark.parametrize("yml,expect", [
    ("melspec_noisy_glow.yml", dict(K=40, L=3, n_filters=512, num_classes=10, T=100, version="v1", progression="logarithmic")),
    ("melspec_ncsnv1.yml", dict(n_filters=192, num_classes=10, T=100, version="v1", sigma1=1.0)),
    ("melspec_ncsnv2.yml", dict(n_filters=128, num_classes=200, T=8, version="v2", sigma1=30.0)),
])
def test_cli_config_precedence(yml, expect):
    """YAML over argparse defaults, CLI-only fields kept, keys missing from the YAML keep their defaults
    (the reference drops them: defect D2)."""
    from audiosourcesep_b200.run_basis_sep import build_parser, merge_config
    args = build_parser().parse_args(["r1", "r2", "--config", os.path.join(ROOT, "configs", yml), "--n_mixed", "7",
                                      "--model_type", "glow", "--output", "o"])
    args = merge_config(args)
    for k, v in expect.items():
        assert getattr(args, k) == v, (k, getattr(args, k), v)
    assert args.n_mixed == 7 and args.model_type == "glow" and args.output == "o" and args.l2_reg is None
    assert isinstance(args.num_classes, int)


def test_shard_range_covers_everything_once():
    from audiosourcesep_b200.run_basis_sep import shard_range
    for n in (0, 1, 7, 30, 31, 64):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                lo, hi = shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                got += list(range(lo, hi))
            assert got == list(range(n)), (n, world)
    assert shard_range(30, 8, 0) == (0, 4) and shard_range(30, 8, 7) == (28, 30)


def test_post_processing_matches_oracle():
    from audiosourcesep_b200.run_basis_sep import post_processing_fn
    from oracle import basis_oracle as bo
    x = np.random.default_rng(0).uniform(-0.3, 1.3, (3, 96, 64)).astype(np.float32)
    args = argparse.Namespace(minval=-100.0, maxval=20.0, use_logit=False, alpha=False)
    np.testing.assert_array_equal(post_processing_fn(args)(x), bo.post_processing(x))


def _gather_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from audiosourcesep_b200.run_basis_sep import gather_segments, shard_range
    full = np.arange(n * 6, dtype=np.float32).reshape(n, 3, 2)
    lo, hi = shard_range(n, world, rank)
    out = gather_segments(full[lo:hi], n)
    conv = gather_segments(np.stack([full[lo:hi], full[lo:hi] + 1]), n, axis=1)
    if rank == 0:
        q.put((np.array_equal(out, full), np.array_equal(conv, np.stack([full, full + 1]))))
    else:
        assert out is None and conv is None
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 30, 1])
def test_segment_gather_world_size_2_gloo(n):
    """The only cross-rank traffic of a separation run: rank 0 reassembles the per-rank segment blocks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok_final, ok_conv = q.get(timeout=10)
    assert ok_final and ok_conv


def test_langevin_constants_match_oracle():
    from audiosourcesep_b200.ncsn.utils import get_sigmas, langevin_step_constants
    from oracle import basis_oracle as bo
    for (s1, sL, n) in ((1.0, 0.01, 10), (30.0, 0.01, 200)):
        sig = get_sigmas(s1, sL, n, "logarithmic")
        np.testing.assert_array_equal(sig, bo.get_sigmas(s1, sL, n, "logarithmic"))
        for i in (0, n // 2, n - 1):
            assert langevin_step_constants(sig, i) == bo.step_constants(sig, i)


class _FakeFlow:
    """Stands in for the libasep Glow handle so the data-parallel plumbing of train_glow can run on CPU/gloo:
    gradient of 0.5*sum((theta - x_i)^2)/global_batch over the local shard."""

    def __init__(self):
        self.theta = torch.zeros(3)
        self.device = torch.device("cpu")

    def train_grads(self, batch, global_batch, noise=None, sigma=0.0):
        diff = self.theta[None, :] - batch.reshape(batch.shape[0], -1)[:, :3]
        return diff.sum(0) / global_batch, (0.5 * diff.pow(2).sum() / global_batch).reshape(1)

    def adamax_step(self, grads, lr, beta1, beta2, eps):
        self.theta = self.theta - lr * grads


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from audiosourcesep_b200.train_glow import distributed_train_step
    full = torch.arange(24, dtype=torch.float32).reshape(4, 6)
    flow = _FakeFlow()
    loss = distributed_train_step(flow, dict(lr=0.5, beta1=0.9, beta2=0.999, eps=1e-7), full[rank * 2:(rank + 1) * 2], 4)
    q.put((rank, flow.theta.numpy().copy(), float(loss)))
    dist.destroy_process_group()


def test_data_parallel_train_step_world_size_2_gloo():
    """Shards + SUM all-reduce reproduce the single-process step on the global batch (train_glow.py:31,42-54)."""
    from audiosourcesep_b200.train_glow import distributed_train_step
    full = torch.arange(24, dtype=torch.float32).reshape(4, 6)
    ref = _FakeFlow()
    loss_ref = float(distributed_train_step(ref, dict(lr=0.5, beta1=0.9, beta2=0.999, eps=1e-7), full, 4))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, theta, loss in out:
        np.testing.assert_allclose(theta, ref.theta.numpy(), rtol=1e-6)
        assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
