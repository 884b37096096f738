"""Python handle of one Glow prior living in libasep.so.

This is the object behind ``flow_models.flow_builder.build_glow``: it offers the TFP
distribution surface the reference's scripts use -- ``log_prob``, ``sample``,
``trainable_variables`` / ``variables`` (reference: flow_models/flow_builder.py:127-144,
train_glow.py:30,39,74) -- and the bijector surface ``forward`` / ``inverse`` /
``forward_log_det_jacobian`` / ``inverse_log_det_jacobian``
(reference: flow_models/flow_glow.py:176-225).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterable, Optional

import numpy as np
import torch

from . import _lib
from .config import GlowConfig
from .weights import glow_param_shapes, is_trainable


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    if t.device != device:
        t = t.to(device)
    return t.contiguous()


class Glow:
    def __init__(self, cfg: GlowConfig, params: Optional[Dict[str, np.ndarray]] = None,
                 precision: int = _lib.PREC_BF16, device: Optional[int] = None):
        self.cfg = cfg
        self.device_index = _lib.init(device)
        self.device = torch.device("cuda", self.device_index)
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        c = _lib.GlowCfg(cfg.H, cfg.W, cfg.C, cfg.L, cfg.K, cfg.n_filters, int(cfg.learntop), cfg.minval, cfg.maxval)
        _lib.check(self._lib.asep_glow_create(ctypes.byref(c), ctypes.byref(self._h)))
        self.precision = precision
        self._shapes = glow_param_shapes(cfg)
        if params is not None:
            self.set_params(params)
            self.prepare(precision)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.asep_glow_destroy(h)
            except Exception:
                pass
            self._h = ctypes.c_void_p()

    # ---- parameters
    def set_param(self, name: str, value) -> None:
        t = torch.as_tensor(np.asarray(value) if not torch.is_tensor(value) else value)
        t = t.detach().to(torch.float32).contiguous()
        d = _lib.dl(t)
        _lib.check(self._lib.asep_glow_set_param(self._h, name.encode(), d.ptr))

    def set_params(self, params: Dict[str, np.ndarray]) -> None:
        for name, value in params.items():
            self.set_param(name, value)

    def get_param(self, name: str) -> np.ndarray:
        out = torch.empty(self._shapes[name], dtype=torch.float32)
        d = _lib.dl(out)
        _lib.check(self._lib.asep_glow_get_param(self._h, name.encode(), d.ptr))
        return out.numpy()

    def prepare(self, precision: Optional[int] = None) -> None:
        if precision is not None:
            self.precision = precision
        _lib.check(self._lib.asep_glow_prepare(self._h, int(self.precision)))

    @property
    def variables(self) -> Dict[str, np.ndarray]:
        return {n: self.get_param(n) for n in self._shapes}

    @property
    def trainable_variables(self) -> Dict[str, np.ndarray]:
        return {n: self.get_param(n) for n in self._shapes if is_trainable(n)}

    def init_actnorm(self, minibatch: torch.Tensor) -> None:
        x = _f32c(minibatch, self.device)
        d = _lib.dl(x)
        _lib.check(self._lib.asep_glow_init_actnorm(self._h, d.ptr, _lib.stream_ptr()))

    # ---- bijector surface (Chain([glow, SpecPreprocessing]))
    def forward_with_log_det(self, x: torch.Tensor):
        x = _f32c(x, self.device)
        N = x.shape[0]
        z = torch.empty((N,) + self.cfg.latent_shape, dtype=torch.float32, device=self.device)
        ld = torch.empty((N,), dtype=torch.float32, device=self.device)
        dx, dz, dl_ = _lib.dl(x), _lib.dl(z), _lib.dl(ld)
        _lib.check(self._lib.asep_glow_forward(self._h, dx.ptr, dz.ptr, dl_.ptr, _lib.stream_ptr()))
        return z, ld

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward_with_log_det(x)[0]

    def forward_log_det_jacobian(self, x: torch.Tensor, event_ndims: int = 3) -> torch.Tensor:
        return self.forward_with_log_det(x)[1]

    def inverse(self, z: torch.Tensor) -> torch.Tensor:
        z = _f32c(z, self.device)
        N = z.shape[0]
        x = torch.empty((N, self.cfg.H, self.cfg.W, self.cfg.C), dtype=torch.float32, device=self.device)
        dz, dx = _lib.dl(z), _lib.dl(x)
        _lib.check(self._lib.asep_glow_inverse(self._h, dz.ptr, dx.ptr, _lib.stream_ptr()))
        return x

    def inverse_log_det_jacobian(self, z: torch.Tensor, event_ndims: int = 3) -> torch.Tensor:
        return -self.forward_log_det_jacobian(self.inverse(z))

    # ---- distribution surface
    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        x = _f32c(x, self.device)
        lp = torch.empty((x.shape[0],), dtype=torch.float32, device=self.device)
        dx, dlp = _lib.dl(x), _lib.dl(lp)
        _lib.check(self._lib.asep_glow_log_prob(self._h, dx.ptr, dlp.ptr, _lib.stream_ptr()))
        return lp

    def grad_log_prob(self, x: torch.Tensor, return_log_prob: bool = False):
        """compute_grad_logprob (reference: run_basis_sep.py:73-79)."""
        x = _f32c(x, self.device)
        g = torch.empty_like(x)
        lp = torch.empty((x.shape[0],), dtype=torch.float32, device=self.device) if return_log_prob else None
        dx, dg, dlp = _lib.dl(x), _lib.dl(g), _lib.dl(lp)
        _lib.check(self._lib.asep_glow_grad_log_prob(self._h, dx.ptr, dg.ptr, dlp.ptr, _lib.stream_ptr()))
        return (g, lp) if return_log_prob else g

    def sample(self, n: int, eps: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None):
        if eps is None:
            eps = torch.randn((n,) + self.cfg.latent_shape, dtype=torch.float32, device=self.device, generator=generator)
        eps = _f32c(eps, self.device)
        x = torch.empty((eps.shape[0], self.cfg.H, self.cfg.W, self.cfg.C), dtype=torch.float32, device=self.device)
        de, dx = _lib.dl(eps), _lib.dl(x)
        _lib.check(self._lib.asep_glow_sample(self._h, de.ptr, dx.ptr, _lib.stream_ptr()))
        return x

    # ---- training (train_glow.py:29-44; Keras Adamax, train_utils.py:29-30)
    def enable_training(self, precision: Optional[int] = None) -> int:
        """Move the trainables into one flat device vector; returns its length.  ``precision``: PREC_FP32 (CUDA-core
        exact mode) or PREC_BF16 / PREC_FP16 (tcgen05 forward / backward / weight-gradient GEMMs); default: keep the
        current one, except that the split-precision inference modes (no weight-gradient path) fall back to PREC_BF16."""
        if precision is None and self.precision in (_lib.PREC_BF16X2, _lib.PREC_FP16X2, _lib.PREC_FP16X3):
            precision = _lib.PREC_BF16
        if precision is not None and precision != self.precision:
            self.prepare(precision)
        _lib.check(self._lib.asep_glow_enable_training(self._h))
        n = ctypes.c_int64()
        _lib.check(self._lib.asep_glow_num_trainable(self._h, ctypes.byref(n)))
        self.num_trainable = int(n.value)
        return self.num_trainable

    def trainable_layout(self):
        """[(name, offset, shape)] of the flat trainable / gradient vector."""
        out, off = [], 0
        for name, shape in self._shapes.items():
            if is_trainable(name):
                n = int(np.prod(shape))
                out.append((name, off, tuple(shape)))
                off += n
        return out

    def train_grads(self, x: torch.Tensor, global_batch: int, noise: Optional[torch.Tensor] = None, sigma: float = 0.0):
        """(grads [num_trainable], loss [1]) of loss = sum_i -log_prob(x_i + sigma*noise_i) / global_batch."""
        x = _f32c(x, self.device)
        noise = None if noise is None else _f32c(noise, self.device)
        grads = torch.empty((self.num_trainable,), dtype=torch.float32, device=self.device)
        loss = torch.empty((1,), dtype=torch.float32, device=self.device)
        dx, dn, dg, dls = _lib.dl(x), _lib.dl(noise), _lib.dl(grads), _lib.dl(loss)
        _lib.check(self._lib.asep_glow_train_grads(self._h, dx.ptr, dn.ptr, float(sigma), int(global_batch), dg.ptr,
                                                   dls.ptr, _lib.stream_ptr()))
        return grads, loss

    def adamax_step(self, grads: torch.Tensor, lr: float = 1e-3, beta1: float = 0.9, beta2: float = 0.999,
                    eps: float = 1e-7) -> None:
        grads = _f32c(grads, self.device)
        dg = _lib.dl(grads)
        _lib.check(self._lib.asep_glow_adamax_step(self._h, dg.ptr, float(lr), float(beta1), float(beta2), float(eps),
                                                   _lib.stream_ptr()))

    def adam_step(self, grads: torch.Tensor, lr: float = 1e-3, beta1: float = 0.9, beta2: float = 0.999,
                  eps: float = 1e-7) -> None:
        """Keras Adam (``optimizer: adam``, reference: train_utils.py:27-28)."""
        grads = _f32c(grads, self.device)
        dg = _lib.dl(grads)
        _lib.check(self._lib.asep_glow_adam_step(self._h, dg.ptr, float(lr), float(beta1), float(beta2), float(eps),
                                                 _lib.stream_ptr()))

    def apply_gradients(self, grads: torch.Tensor, optimizer: dict) -> None:
        """``optimizer.apply_gradients`` (reference: train_glow.py:42-43) with the dictionary ``setUp_optimizer`` returns."""
        opt = dict(optimizer)
        kind = opt.pop("kind", "adamax")
        if kind == "adamax":
            self.adamax_step(grads, **opt)
        elif kind == "adam":
            self.adam_step(grads, **opt)
        else:
            raise ValueError("optimizer argument should be adam or adamax")

    def get_flat(self) -> torch.Tensor:
        t = torch.empty((self.num_trainable,), dtype=torch.float32, device=self.device)
        d = _lib.dl(t)
        _lib.check(self._lib.asep_glow_get_flat(self._h, d.ptr, _lib.stream_ptr()))
        return t

    def set_flat(self, theta: torch.Tensor) -> None:
        theta = _f32c(theta, self.device)
        d = _lib.dl(theta)
        _lib.check(self._lib.asep_glow_set_flat(self._h, d.ptr, _lib.stream_ptr()))

    def sync_host(self) -> None:
        _lib.check(self._lib.asep_glow_sync_host(self._h))

    # ---- coupling network of one step (test / profiling seam)
    def coupling_nn(self, block: int, step: int, state: torch.Tensor) -> torch.Tensor:
        state = _f32c(state, self.device)
        r = torch.empty_like(state)
        ds, dr = _lib.dl(state), _lib.dl(r)
        _lib.check(self._lib.asep_glow_coupling_nn(self._h, block, step, ds.ptr, dr.ptr, _lib.stream_ptr()))
        return r

    def coupling_nn_backward(self, block: int, step: int, state: torch.Tensor, gr: torch.Tensor) -> torch.Tensor:
        state, gr = _f32c(state, self.device), _f32c(gr, self.device)
        out = torch.empty(state.shape[:-1] + (state.shape[-1] // 2,), dtype=torch.float32, device=self.device)
        ds, dg, do = _lib.dl(state), _lib.dl(gr), _lib.dl(out)
        _lib.check(self._lib.asep_glow_coupling_nn_backward(self._h, block, step, ds.ptr, dg.ptr, do.ptr,
                                                            _lib.stream_ptr()))
        return out

    @property
    def handle(self) -> ctypes.c_void_p:
        return self._h
