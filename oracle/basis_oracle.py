"""CPU ORACLE (test infrastructure, NOT a product path) -- BASIS annealed Langevin loop.

Restates, in numpy float32 with the reference's evaluation order:
  ncsn/utils.py:7-14            get_sigmas
  run_basis_sep.py:106-149      mixing_process (dB / power / image g and grad_g)
  run_basis_sep.py:152-181      basis_inner_loop update (eta, lambda, noise, old-state update)
  run_basis_sep.py:217-260      basis_outer_loop (sigma loop + snapshots)
  run_basis_sep.py:82-96        post_processing_fn (melspec, use_logit False/True)
  ncsn/utils.py:17-38           anneal_langevin_dynamics (K=1, lambda=0 special case)
Pins: the sigma schedule printed in the shipped run log
(basis_sep_results/beethoven_sonata_1_sep_1min/out.log:44-116) -- tests/golden/sigmas_v1.json.
Langevin states themselves have no reference golden vector: "parity unpinned".
Noise is INJECTED (arrays of standard normals) so that CUDA and oracle see the same draws.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

LN10 = np.float32(np.log(10.0))


def get_sigmas(sigma1, sigmaL, num_classes, progression="geometric"):
    """ncsn/utils.py:7-14."""
    if progression == "geometric":
        sigmas = np.exp(np.linspace(np.log(sigma1), np.log(sigmaL), num=num_classes))
    elif progression == "logarithmic":
        sigmas = np.logspace(np.log(sigma1) / np.log(10), np.log(sigmaL) / np.log(10), num=num_classes)
    else:
        raise ValueError("progression should be geometric or logarithmic")
    return sigmas.astype(np.float32)


def mixing_process(data_type: str = "melspec", scale: str = "dB"):
    """run_basis_sep.py:106-149 -> (g, grad_g) on float32 numpy arrays."""
    if data_type == "image":
        def g(*sources):
            return np.mean(np.stack(sources, 0), axis=0, dtype=np.float32)

        def grad_g(*sources):
            K = len(sources)
            return [np.ones_like(s, dtype=np.float32) / np.float32(K) for s in sources]
    elif scale == "power":
        def g(*sources):
            s = np.stack(sources, 0).astype(np.float32)
            return np.mean(np.sqrt(s), axis=0, dtype=np.float32) ** 2

        def grad_g(*sources):
            s = np.stack(sources, 0).astype(np.float32)
            gs = 1.0 / (np.sqrt(s) + np.float32(1e-8))
            gs = gs * np.mean(np.sqrt(s), axis=0, dtype=np.float32, keepdims=True) ** 2
            return [gs[i] for i in range(len(sources))]
    else:
        def g(*sources):
            K = len(sources)
            s = np.stack(sources, 0).astype(np.float32) * LN10 / np.float32(10.0)
            m = s.max(axis=0)
            lse = m + np.log(np.sum(np.exp(s - m), axis=0, dtype=np.float32))
            return ((np.float32(10.0) / LN10) * (lse - np.float32(np.log(float(K))))).astype(np.float32)

        def grad_g(*sources):
            s = np.stack(sources, 0).astype(np.float32) * LN10 / np.float32(10.0)
            e = np.exp(s - s.max(axis=0, keepdims=True))
            sm = e / e.sum(axis=0, keepdims=True)
            return [sm[i].astype(np.float32) for i in range(len(sources))]
    return g, grad_g


def step_constants(sigmas: np.ndarray, sigma_idx: int, delta: float = 2e-5) -> Tuple[np.float32, np.float32, np.float32]:
    """eta, lambda, sqrt(2 eta) exactly as run_basis_sep.py:158-164 casts them.

    ``sigmas`` are float32 numpy scalars; ``delta * (sigma / sigmaL) ** 2`` is evaluated in
    numpy float32/Python-float mixed arithmetic (float32 scalar ops promote with the Python
    float ``delta`` to float64 under NumPy 1.x value-based casting as shipped with TF 2.2),
    then cast by ``tf.constant(..., float32)``.
    """
    sigma = np.float32(sigmas[sigma_idx])
    ratio = np.float32(sigma / np.float32(sigmas[-1]))
    ratio_sq = np.float32(ratio * ratio)                       # float32 ** 2 stays float32
    eta = np.float32(np.float64(delta) * np.float64(ratio_sq))  # Python float * np.float32 -> float64
    lam = np.float32(1.0 / np.float64(np.float32(sigma * sigma)))
    noise_scale = np.float32(np.sqrt(np.float32(np.float32(2.0) * eta)))
    return eta, lam, noise_scale


def langevin_update(x1, x2, s1, s2, mixed, n1, n2, eta, lam, noise_scale, g, grad_g):
    """One update of run_basis_sep.py:163-181; n1/n2 are standard-normal draws."""
    eps1 = noise_scale * n1
    eps2 = noise_scale * n2
    mixing = g(x1, x2)
    gm1, gm2 = grad_g(x1, x2)
    x1n = x1 + eta * (s1 + lam * gm1 * (mixed - mixing)) + eps1
    x2n = x2 + eta * (s2 + lam * gm2 * (mixed - mixing)) + eps2
    return x1n.astype(np.float32), x2n.astype(np.float32)


def basis_run(mixed, x1, x2, score1: Callable, score2: Callable, sigmas, T: int,
              noise: Callable[[int, int], Tuple[np.ndarray, np.ndarray]],
              delta: float = 2e-5, data_type="melspec", scale="dB",
              per_step: Optional[List] = None):
    """basis_outer_loop + basis_inner_loop.  ``score_k(x, sigma_idx) -> grad log p`` (float32);
    ``noise(sigma_idx, t) -> (n1, n2)`` standard normal arrays.  Returns x1, x2, x_arr."""
    g, grad_g = mixing_process(data_type, scale)
    x_arr = {"x1": [x1.copy()], "x2": [x2.copy()]}
    for i in range(len(sigmas)):
        eta, lam, ns = step_constants(sigmas, i, delta)
        for t in range(T):
            n1, n2 = noise(i, t)
            s1 = score1(x1, i)
            s2 = score2(x2, i)
            x1, x2 = langevin_update(x1, x2, s1, s2, mixed, n1, n2, eta, lam, ns, g, grad_g)
            if per_step is not None:
                per_step.append((x1.copy(), x2.copy()))
        x_arr["x1"].append(x1.copy())
        x_arr["x2"].append(x2.copy())
    return x1, x2, x_arr


def post_processing(x, minval=-100.0, maxval=20.0, use_logit=False, alpha=1e-10):
    """run_basis_sep.py:82-96 (melspec, dB)."""
    x = np.asarray(x, dtype=np.float32)
    if use_logit:
        x = 1.0 / (1.0 + np.exp(-x))
        x = (x - alpha) / (1.0 - 2.0 * alpha)
    x = x * (maxval - minval) + minval
    return np.clip(x, minval, maxval)


def anneal_langevin_dynamics(x, score: Callable, sigmas, n_steps_each, step_lr, noise):
    """ncsn/utils.py:17-38 with injected noise."""
    for i in range(len(sigmas)):
        ratio = np.float32(np.float32(sigmas[i]) / np.float32(sigmas[-1]))
        step = np.float32(np.float64(step_lr) * np.float64(np.float32(ratio * ratio)))
        for s in range(n_steps_each):
            n = noise(i, s)
            x = (x + step * score(x, i) + n * np.float32(np.sqrt(step * np.float32(2.0)))).astype(np.float32)
    return x


def sdr_db(ref: np.ndarray, est: np.ndarray) -> float:
    """Plain 10 log10(||s||^2 / ||s - s_hat||^2) on flattened patches."""
    ref = np.asarray(ref, np.float64).ravel()
    est = np.asarray(est, np.float64).ravel()
    return float(10.0 * np.log10(np.sum(ref ** 2) / np.sum((ref - est) ** 2)))
