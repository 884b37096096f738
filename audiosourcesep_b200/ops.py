"""Functional access to the stateless kernels of libasep.so (single bijectors, Langevin update)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def _prep(t: torch.Tensor) -> torch.Tensor:
    dev = torch.device("cuda", _lib.init())
    if t.dtype != torch.float32:
        t = t.float()
    return t.to(dev).contiguous()


def actnorm(x, log_scale, shift, inverse: bool = False) -> torch.Tensor:
    """ActNorm._forward / _inverse (reference: flow_tfp_bijectors.py:242-247)."""
    x, log_scale, shift = _prep(x), _prep(log_scale), _prep(shift)
    y = torch.empty_like(x)
    a, b, c, d = _lib.dl(x), _lib.dl(log_scale), _lib.dl(shift), _lib.dl(y)
    _lib.check(_lib.load().asep_actnorm(a.ptr, b.ptr, c.ptr, d.ptr, int(inverse), _lib.stream_ptr()))
    return y


def inv1x1(x, w) -> torch.Tensor:
    """y[..., o] = sum_i x[..., i] w[i, o]  (reference: flow_tfp_bijectors.py:304-305, :316)."""
    x, w = _prep(x), _prep(w)
    y = torch.empty_like(x)
    a, b, c = _lib.dl(x), _lib.dl(w), _lib.dl(y)
    _lib.check(_lib.load().asep_inv1x1(a.ptr, b.ptr, c.ptr, _lib.stream_ptr()))
    return y


def coupling(x, r, inverse: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """AffineCouplingLayerSplit given r = [raw_log_s | t]; returns (y, log-det of this direction)."""
    x, r = _prep(x), _prep(r)
    y = torch.empty_like(x)
    ld = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
    a, b, c, d = _lib.dl(x), _lib.dl(r), _lib.dl(y), _lib.dl(ld)
    _lib.check(_lib.load().asep_coupling(a.ptr, b.ptr, c.ptr, d.ptr, int(inverse), _lib.stream_ptr()))
    return y, ld


def squeeze(x, inverse: bool = False) -> torch.Tensor:
    """Squeeze._forward / _inverse (reference: flow_tfp_bijectors.py:170-180)."""
    x = _prep(x)
    N, H, W, C = x.shape
    if not inverse:
        if H % 2 or W % 2:
            raise ValueError("Squeeze needs even H and W")
        y = torch.empty((N, H // 2, W // 2, 4 * C), dtype=torch.float32, device=x.device)
    else:
        if C % 4:
            raise ValueError("inverse Squeeze needs C divisible by 4")
        y = torch.empty((N, H * 2, W * 2, C // 4), dtype=torch.float32, device=x.device)
    a, b = _lib.dl(x), _lib.dl(y)
    _lib.check(_lib.load().asep_squeeze(a.ptr, b.ptr, int(inverse), _lib.stream_ptr()))
    return y


def mixing_db(x1, x2):
    """g(x1,x2) and grad_g (reference: run_basis_sep.py:131-147)."""
    x1, x2 = _prep(x1), _prep(x2)
    g, w1, w2 = torch.empty_like(x1), torch.empty_like(x1), torch.empty_like(x1)
    t = [_lib.dl(v) for v in (x1, x2, g, w1, w2)]
    _lib.check(_lib.load().asep_mixing_db(*(v.ptr for v in t), _lib.stream_ptr()))
    return g, w1, w2


def langevin_step(x1, x2, s1, s2, mixed, eta: float, lam: float, noise_scale: float,
                  n1: Optional[torch.Tensor] = None, n2: Optional[torch.Tensor] = None, seed: int = 0,
                  step: int = 0, elem_offset: int = 0, nan_count: Optional[torch.Tensor] = None) -> None:
    """In-place fused update of both sources (reference: run_basis_sep.py:163-181)."""
    for t in (x1, x2):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("x1/x2 must be contiguous float32 CUDA tensors (updated in place)")
    s1, s2, mixed = _prep(s1), _prep(s2), _prep(mixed)
    n1 = None if n1 is None else _prep(n1)
    n2 = None if n2 is None else _prep(n2)
    t = [_lib.dl(v) for v in (x1, x2, s1, s2, mixed, n1, n2)]
    dn = _lib.dl(nan_count)
    _lib.check(_lib.load().asep_langevin_step(*(v.ptr for v in t), float(eta), float(lam), float(noise_scale),
                                              int(seed), int(step), int(elem_offset), dn.ptr, _lib.stream_ptr()))


def philox_normal(shape, seed: int, step: int, stream_id: int, elem_offset: int = 0) -> torch.Tensor:
    dev = torch.device("cuda", _lib.init())
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    d = _lib.dl(out)
    _lib.check(_lib.load().asep_philox_normal(d.ptr, int(seed), int(step), int(stream_id), int(elem_offset),
                                              _lib.stream_ptr()))
    return out


def basis_glow_inner(m1, m2, mixed, x1, x2, T: int, eta: float, lam: float, noise_scale: float,
                     noise1=None, noise2=None, seed: int = 0, step0: int = 0, elem_offset: int = 0,
                     per_step: Optional[torch.Tensor] = None, nan_count: Optional[torch.Tensor] = None) -> None:
    """T Langevin steps at one noise level with two Glow priors, entirely inside the library."""
    mixed = _prep(mixed)
    noise1 = None if noise1 is None else _prep(noise1)
    noise2 = None if noise2 is None else _prep(noise2)
    t = [_lib.dl(v) for v in (mixed, x1, x2)]
    dn1, dn2, dps, dnan = _lib.dl(noise1), _lib.dl(noise2), _lib.dl(per_step), _lib.dl(nan_count)
    _lib.check(_lib.load().asep_basis_glow_inner(m1.handle, m2.handle, t[0].ptr, t[1].ptr, t[2].ptr, int(T),
                                                 float(eta), float(lam), float(noise_scale), dn1.ptr, dn2.ptr,
                                                 int(seed), int(step0), int(elem_offset), dps.ptr, dnan.ptr,
                                                 _lib.stream_ptr()))


def basis_ncsn_inner(m1, m2, mixed, x1, x2, sigma_idx: int, T: int, eta: float, lam: float, noise_scale: float,
                     noise1=None, noise2=None, seed: int = 0, step0: int = 0, elem_offset: int = 0,
                     per_step: Optional[torch.Tensor] = None, nan_count: Optional[torch.Tensor] = None) -> None:
    """T Langevin steps at noise level ``sigma_idx`` with two NCSN score networks, inside the library."""
    mixed = _prep(mixed)
    noise1 = None if noise1 is None else _prep(noise1)
    noise2 = None if noise2 is None else _prep(noise2)
    t = [_lib.dl(v) for v in (mixed, x1, x2)]
    dn1, dn2, dps, dnan = _lib.dl(noise1), _lib.dl(noise2), _lib.dl(per_step), _lib.dl(nan_count)
    _lib.check(_lib.load().asep_basis_ncsn_inner(m1.handle, m2.handle, t[0].ptr, t[1].ptr, t[2].ptr, int(sigma_idx),
                                                 int(T), float(eta), float(lam), float(noise_scale), dn1.ptr, dn2.ptr,
                                                 int(seed), int(step0), int(elem_offset), dps.ptr, dnan.ptr,
                                                 _lib.stream_ptr()))


def basis_run(m1, m2, mixed, x1, x2, T: int, eta, lam, noise_scale, seed: int = 0, elem_offset: int = 0,
              snapshots: Optional[torch.Tensor] = None, nan_count: Optional[torch.Tensor] = None) -> None:
    """The whole sigma x T loop inside the library (reference: run_basis_sep.py:217-260): ``eta / lam / noise_scale`` are
    length-L sequences of the per-level float32 constants.  ``m1 / m2``: two score networks, two Glow priors, or two
    length-L lists of Glow priors (one fine-tuned pair per noise level)."""
    import ctypes
    L = len(eta)
    arr = [(ctypes.c_float * L)(*[float(v) for v in seq]) for seq in (eta, lam, noise_scale)]
    mixed = _prep(mixed)
    t = [_lib.dl(v) for v in (mixed, x1, x2)]
    dsn, dnan = _lib.dl(snapshots), _lib.dl(nan_count)
    first = m1[0] if isinstance(m1, (list, tuple)) else m1
    if hasattr(first, "cfg") and hasattr(first.cfg, "K"):                 # Glow priors
        l1 = list(m1) if isinstance(m1, (list, tuple)) else [m1]
        l2 = list(m2) if isinstance(m2, (list, tuple)) else [m2]
        h1 = (ctypes.c_void_p * len(l1))(*[m.handle.value for m in l1])
        h2 = (ctypes.c_void_p * len(l2))(*[m.handle.value for m in l2])
        _lib.check(_lib.load().asep_basis_glow_run(h1, h2, len(l1), t[0].ptr, t[1].ptr, t[2].ptr, L, int(T), arr[0], arr[1],
                                                   arr[2], int(seed), int(elem_offset), dsn.ptr, dnan.ptr, _lib.stream_ptr()))
    else:
        _lib.check(_lib.load().asep_basis_ncsn_run(m1.handle, m2.handle, t[0].ptr, t[1].ptr, t[2].ptr, L, int(T), arr[0], arr[1],
                                                   arr[2], int(seed), int(elem_offset), dsn.ptr, dnan.ptr, _lib.stream_ptr()))
