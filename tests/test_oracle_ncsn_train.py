"""CPU tests of the NCSN train-step oracle (oracle/train_ncsn_oracle.py) and of the host logic of train_ncsn."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist

from audiosourcesep_b200 import NCSNConfig
from audiosourcesep_b200.ncsn.utils import get_sigmas
from audiosourcesep_b200.weights import init_ncsn_params
from oracle import train_ncsn_oracle as to
from oracle.ncsn_oracle import NCSNOracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("version", ["v1", "v2"])
def test_dsm_oracle_gradients_match_finite_differences(version):
    cfg = NCSNConfig(version=version, H=16, W=16, ngf=64, num_classes=5, sigma1=1.0, sigmaL=0.01)
    p = init_ncsn_params(cfg, seed=5, mode="perturbed")
    sig = get_sigmas(cfg.sigma1, cfg.sigmaL, cfg.num_classes, "logarithmic")
    rng = np.random.default_rng(0)
    x = rng.random((2, 16, 16, 1)).astype(np.float32)
    z = rng.standard_normal((2, 16, 16, 1)).astype(np.float32)
    idx = np.array([1, 4])
    loss, g = to.dsm_loss_and_grads(cfg, p, sig, x, z, idx, 4)
    assert np.isfinite(loss) and loss > 0
    names = ["Res3_1/conv1/kernel", "refine2/CRP/conv_1/kernel", "normalizer/in_gamma", "begin_conv/kernel", "end_conv/bias",
             "Res1_1/norm1/" + ("embed" if version == "v1" else "alpha")]
    for name in names:
        i = tuple(int(v) // 2 for v in p[name].shape)
        if name.endswith("embed"):
            i = (1, i[1])                                   # a row that a sample of the batch actually selects
        eps = 1e-4
        vals = []
        for sgn in (+1, -1):
            pp = {k: v.astype(np.float64).copy() for k, v in p.items()}
            pp[name][i] += sgn * eps
            vals.append(float(to.dsm_loss(NCSNOracle(cfg, pp, sigmas=sig), sig, x, z, idx, 4)))
        fd = (vals[0] - vals[1]) / (2 * eps)
        assert abs(fd - g[name][i]) <= 1e-5 * max(1.0, abs(fd)), (name, fd, g[name][i])
    if version == "v1":                                     # Embedding rows no sample selects receive no gradient
        assert np.all(g["Res1_1/norm1/embed"][[0, 2, 3]] == 0.0)


def test_dsm_loss_known_answer():
    """A score network that outputs exactly the target gives zero loss; scaling the residual by a doubles it 4x."""
    class _Fake:
        dtype = torch.float64

        def __init__(self, a):
            self.a = a

        def score(self, xt, idx):
            return self.a * torch.ones_like(xt)

    sig = np.array([2.0, 0.5])
    x = np.zeros((2, 4, 4, 1))
    z = np.zeros((2, 4, 4, 1))
    # z = 0: target = 0, loss = 1/2 * sum a^2 * sigma^2 / global_batch = 1/2 * 16 a^2 (4 + 0.25) / 2
    for a in (1.0, 2.0):
        loss = float(to.dsm_loss(_Fake(a), sig, x, z, np.array([0, 1]), 2))
        assert abs(loss - 0.5 * 16 * a * a * 4.25 / 2) < 1e-12


def test_noise_level_sampling_reproduces_the_channel_count_quirk():
    """train_ncsn.py:34: local_batch_size = X.shape[-1] -> ONE level per replica batch for 1-channel patches."""
    from audiosourcesep_b200.train_ncsn import get_noise_conditionned_data, setUp_optimizer, shard
    gen = torch.Generator(device="cpu").manual_seed(0)
    X = torch.zeros((6, 8, 8, 1))
    idx, z = get_noise_conditionned_data(X, 10, gen)
    assert idx.shape == (6,) and idx.dtype == torch.int32 and len(set(idx.tolist())) == 1
    assert z.shape == X.shape and abs(float(z.std()) - 1.0) < 0.1
    idx2, _ = get_noise_conditionned_data(X, 1000, gen, per_sample=True)
    assert len(set(idx2.tolist())) > 1
    with pytest.raises(ValueError):
        get_noise_conditionned_data(torch.zeros((6, 8, 8, 2)), 10, gen)
    perm = np.arange(16)
    assert sorted(np.concatenate([shard(perm, 1, 8, r, 2) for r in range(2)]).tolist()) == list(range(8, 16))
    class A: optimizer = "sgd"; learning_rate = 1e-3
    with pytest.raises(ValueError):
        setUp_optimizer(None, A)


class _FakeScoreModel:
    """Stands in for the libasep handle: DSM-shaped quadratic loss whose gradient is linear in the local shard."""

    def __init__(self):
        self.theta = torch.zeros(3)

    def train_grads(self, batch, z, idx, global_batch):
        diff = self.theta[None, :] - (batch + z).reshape(batch.shape[0], -1)[:, :3]
        return diff.sum(0) / global_batch, (0.5 * diff.pow(2).sum() / global_batch).reshape(1)

    def apply_gradients(self, grads, optimizer):
        self.theta = self.theta - optimizer["lr"] * grads


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from audiosourcesep_b200.train_ncsn import distributed_train_step
    full = torch.arange(24, dtype=torch.float32).reshape(4, 6)
    z = torch.ones_like(full)
    m = _FakeScoreModel()
    sl = slice(rank * 2, (rank + 1) * 2)
    loss = distributed_train_step(m, dict(kind="adam", lr=0.5), full[sl], 4, torch.zeros(2, dtype=torch.int32), z[sl])
    q.put((rank, m.theta.numpy().copy(), float(loss)))
    dist.destroy_process_group()


def test_ncsn_data_parallel_step_world_size_2_gloo():
    """Shards + SUM all-reduce reproduce the single-process step on the global batch (train_ncsn.py:26-29,57-62)."""
    from audiosourcesep_b200.train_ncsn import distributed_train_step
    full = torch.arange(24, dtype=torch.float32).reshape(4, 6)
    ref = _FakeScoreModel()
    loss_ref = float(distributed_train_step(ref, dict(kind="adam", lr=0.5), full, 4, torch.zeros(4, dtype=torch.int32), torch.ones_like(full)))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, theta, loss in out:
        np.testing.assert_allclose(theta, ref.theta.numpy(), rtol=1e-6)
        assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
