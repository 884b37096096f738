"""Ideal-mask oracle systems on mel spectrograms -- mirror of the reference's ``oracle_systems.py:264-350``
(``IBM_melspec`` / ``IRM_melspec``); the masking runs in libasep.so (csrc/bsseval.cu: k_ideal_mask).
The STFT-domain IRM / IBM / MWF functions of the reference (oracle_systems.py:19-262, written for musdb tracks) are not
on the separation path and are not rebuilt."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _mask(mixture, sources, binary: bool, theta: float, device=None):
    dev = torch.device("cuda", _lib.init(device))
    as_np = not torch.is_tensor(sources)
    m = torch.as_tensor(np.asarray(mixture, dtype=np.float32) if as_np else mixture, dtype=torch.float32).to(dev).contiguous()
    s = torch.as_tensor(np.asarray(sources, dtype=np.float32) if as_np else sources, dtype=torch.float32).to(dev).contiguous()
    if tuple(s.shape[1:]) != tuple(m.shape):
        raise ValueError(f"sources {tuple(s.shape)} must be [nsrc, *mixture.shape] with mixture {tuple(m.shape)}")
    out = torch.empty_like(s)
    dm, ds, do = _lib.dl(m), _lib.dl(s), _lib.dl(out)
    _lib.check(_lib.load().asep_ideal_mask(dm.ptr, ds.ptr, do.ptr, int(binary), float(theta), _lib.stream_ptr()))
    return out.cpu().numpy() if as_np else out


def IBM_melspec(mixture, sources, theta=0.5):
    """Ideal Binary Mask (oracle_systems.py:264-311): the mixture bin goes to a source when source / (eps + mixture) >= theta."""
    return _mask(mixture, sources, True, theta)


def IRM_melspec(mixture, sources, alpha=2):
    """Ideal Ratio Mask (oracle_systems.py:314-350): mixture * source / (sum of sources + eps).  ``alpha`` is accepted and
    ignored, as in the reference (its body never uses it)."""
    return _mask(mixture, sources, False, 0.0)
