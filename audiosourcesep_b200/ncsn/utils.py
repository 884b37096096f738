"""Host-side mirror of the reference's ``ncsn/utils.py`` (noise schedule and score-model builders)."""
from __future__ import annotations

import numpy as np


def get_sigmas(sigma1, sigmaL, num_classes, progression="geometric"):
    """Geometric noise schedule sigma_1 ... sigma_L as float32 (reference: ncsn/utils.py:7-14)."""
    if progression == "geometric":
        sigmas = np.exp(np.linspace(np.log(sigma1), np.log(sigmaL), num=num_classes))
    elif progression == "logarithmic":
        sigmas = np.logspace(np.log(sigma1) / np.log(10), np.log(sigmaL) / np.log(10), num=num_classes)
    else:
        raise ValueError("progression should be geometric or logarithmic")
    return sigmas.astype(np.float32)


def langevin_step_constants(sigmas, sigma_idx, delta=2e-5):
    """eta, lambda and sqrt(2 eta) with the reference's float32/float64 casts
    (reference: run_basis_sep.py:158-164; delta is hard-wired to 2e-5 at :239)."""
    sigma = np.float32(sigmas[sigma_idx])
    ratio = np.float32(sigma / np.float32(sigmas[-1]))
    eta = np.float32(np.float64(delta) * np.float64(np.float32(ratio * ratio)))
    lam = np.float32(1.0 / np.float64(np.float32(sigma * sigma)))
    noise_scale = np.float32(np.sqrt(np.float32(np.float32(2.0) * eta)))
    return eta, lam, noise_scale
