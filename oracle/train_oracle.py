"""CPU ORACLE (test infrastructure, NOT a product path) -- Glow training step.

Restates, on top of ``GlowOracle`` with torch autograd in float64:
  train_glow.py:29-31            compute_train_loss: sum_i -log_prob(x_i) / global_batch_size
  train_glow.py:38-43            tape.gradient w.r.t. flow.trainable_variables, optimizer.apply_gradients
  train_noisy_glow.py:30-33      X <- X + sigma * N(0,1) (raw data units) before log_prob
  train_utils.py:29-30           Keras Adamax (tf.keras 2.2: m = b1 m + (1-b1) g; u = max(b2 u, |g|);
                                 theta -= lr / (1 - b1^t) * m / (u + eps), eps = 1e-7)
No reference golden vector exists for gradients ("parity unpinned"); the restatement is cross-checked by finite
differences in tests/test_oracle_train.py.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from audiosourcesep_b200.weights import is_trainable
from .glow_oracle import GlowOracle


def loss_and_grads(cfg, params: Dict[str, np.ndarray], x: np.ndarray, global_batch: int,
                   noise: Optional[np.ndarray] = None, sigma: float = 0.0) -> Tuple[float, Dict[str, np.ndarray]]:
    o = GlowOracle(cfg, params, dtype=torch.float64)
    names = [n for n in params if is_trainable(n)]
    for n in names:
        o.p[n].requires_grad_(True)
    xin = torch.as_tensor(np.asarray(x), dtype=torch.float64)
    if noise is not None:
        xin = xin + float(sigma) * torch.as_tensor(np.asarray(noise), dtype=torch.float64)
    loss = -o.log_prob(xin).sum() / float(global_batch)
    grads = torch.autograd.grad(loss, [o.p[n] for n in names], allow_unused=True)
    out = {}
    for n, g in zip(names, grads):
        out[n] = np.zeros(params[n].shape, np.float64) if g is None else g.detach().numpy()
    return float(loss.detach()), out


def mask_structural(name: str, g: np.ndarray) -> np.ndarray:
    """The reference multiplies L by the strictly-lower mask and U by the strictly-upper mask
    (flow_tfp_bijectors.py:300-303), so the masked entries receive zero gradient."""
    if name.endswith("inv1x1/L"):
        return np.tril(g, -1)
    if name.endswith("inv1x1/U"):
        return np.triu(g, 1)
    return g


def adamax_update(theta, g, m, u, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """One Keras Adamax step in float32; returns (theta, m, u)."""
    theta, g, m, u = (np.asarray(a, np.float32) for a in (theta, g, m, u))
    m = (np.float32(b1) * m + np.float32(1.0 - b1) * g).astype(np.float32)
    u = np.maximum(np.float32(b2) * u, np.abs(g)).astype(np.float32)
    lr_t = np.float32(lr / (1.0 - b1 ** t))
    theta = (theta - lr_t * m / (u + np.float32(eps))).astype(np.float32)
    return theta, m, u


def adam_update(theta, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """One Keras Adam step (OptimizerV2, non-amsgrad; train_utils.py:27-28) in float32; returns (theta, m, v)."""
    theta, g, m, v = (np.asarray(a, np.float32) for a in (theta, g, m, v))
    m = (np.float32(b1) * m + np.float32(1.0 - b1) * g).astype(np.float32)
    v = (np.float32(b2) * v + np.float32(1.0 - b2) * g * g).astype(np.float32)
    lr_t = np.float32(lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t))
    theta = (theta - lr_t * m / (np.sqrt(v) + np.float32(eps))).astype(np.float32)
    return theta, m, v

