"""CPU ORACLE (test infrastructure, NOT a product path) -- NCSN denoising-score-matching training step.

Restates, on top of ``NCSNOracle`` with torch autograd in float64:
  train_ncsn.py:33-44   get_noise_conditionned_data: used_sigma = sigmas[sigma_idx]; noise = N(0,1) * used_sigma;
                        perturbed_X = X + noise; target = -noise / used_sigma^2; sample_weight = used_sigma^2
  train_ncsn.py:26-29   compute_train_loss: 1/2 * sum_{h,w,c} (scores - target)^2 * sample_weight, averaged with
                        tf.nn.compute_average_loss(global_batch_size) = sum / global batch
  train_ncsn.py:46-54   tape.gradient w.r.t. model.trainable_variables, optimizer.apply_gradients (Adam: train_oracle.py)
Quirk kept by the HOST code, not here: local_batch_size = X.shape[-1] (:34) is the channel count, so with 1-channel
patches ONE noise level is drawn per replica batch; the oracle takes whatever ``sigma_idx`` it is given.
No reference golden vector exists for these gradients ("parity unpinned"); the restatement is cross-checked by finite
differences in tests/test_oracle_ncsn_train.py.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from .ncsn_oracle import NCSNOracle


def dsm_loss(oracle: NCSNOracle, sigmas, x, z, sigma_idx, global_batch: int) -> torch.Tensor:
    """``z``: standard-normal draws (the reference's noise is z * used_sigma)."""
    dt = oracle.dtype
    idx = torch.as_tensor(np.asarray(sigma_idx), dtype=torch.long).reshape(-1)
    x = torch.as_tensor(np.asarray(x), dtype=dt)
    z = torch.as_tensor(np.asarray(z), dtype=dt)
    if idx.numel() == 1:
        idx = idx.repeat(x.shape[0])
    used = torch.as_tensor(np.asarray(sigmas), dtype=dt)[idx].reshape(-1, 1, 1, 1)
    noise = z * used
    scores = oracle.score(x + noise, idx)
    target = -noise / used ** 2
    per_example = 0.5 * ((scores - target) ** 2).sum(dim=(1, 2, 3)) * used.reshape(-1) ** 2
    return per_example.sum() / float(global_batch)


def dsm_loss_and_grads(cfg, params: Dict[str, np.ndarray], sigmas, x, z, sigma_idx,
                       global_batch: int) -> Tuple[float, Dict[str, np.ndarray]]:
    o = NCSNOracle(cfg, params, sigmas=sigmas, dtype=torch.float64)
    names = list(params)
    for n in names:
        o.p[n].requires_grad_(True)
    loss = dsm_loss(o, sigmas, x, z, sigma_idx, global_batch)
    grads = torch.autograd.grad(loss, [o.p[n] for n in names], allow_unused=True)
    out = {}
    for n, g in zip(names, grads):
        out[n] = np.zeros(params[n].shape, np.float64) if g is None else g.detach().numpy()
    return float(loss.detach()), out
