"""Python handle of one NCSN score network living in libasep.so.

The object behind ``ncsn.utils.get_uncompiled_model[_v2]``: callable like the reference's Keras model,
``model([perturbed_X, sigma_idx], training=True) -> score`` or with a dict keyed ``'perturbed_X'`` /
``'sigma_idx'`` (reference: ncsn/utils.py:41-64, train_ncsn.py:43,52, run_basis_sep.py:167-170).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch

from .. import _lib
from ..config import NCSNConfig
from ..weights import ncsn_param_shapes


class ScoreModel:
    def __init__(self, cfg: NCSNConfig, params: Dict[str, np.ndarray], sigmas=None, device: Optional[int] = None,
                 name: str = "ScoreNetwork", precision: int = _lib.PREC_BF16):
        """``precision``: ``_lib.PREC_BF16`` (one bf16 tcgen05 product per convolution, the throughput mode) or
        ``_lib.PREC_BF16X3`` (split-bf16 operands, three products: matches the fp32 reference to ~1e-5)."""
        self.cfg = cfg
        self.precision = int(precision)
        self.name = name
        self.device_index = _lib.init(device)
        self.device = torch.device("cuda", self.device_index)
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        version = 1 if cfg.version == "v1" else 2
        c = _lib.NcsnCfg(version, cfg.H, cfg.W, cfg.C, cfg.ngf, cfg.num_classes)
        _lib.check(self._lib.asep_ncsn_create(ctypes.byref(c), ctypes.byref(self._h)))
        self._shapes = ncsn_param_shapes(cfg)
        self._params = {}
        if sigmas is not None:
            self.set_sigmas(sigmas)
        self.set_params(params)
        _lib.check(self._lib.asep_ncsn_set_precision(self._h, self.precision))
        self.prepare()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.asep_ncsn_destroy(h)
            except Exception:
                pass
            self._h = ctypes.c_void_p()

    def set_sigmas(self, sigmas) -> None:
        t = torch.as_tensor(np.asarray(sigmas, dtype=np.float32)).contiguous()
        d = _lib.dl(t)
        _lib.check(self._lib.asep_ncsn_set_sigmas(self._h, d.ptr))

    def set_params(self, params: Dict[str, np.ndarray]) -> None:
        missing = [n for n in self._shapes if n not in params]
        if missing:
            raise ValueError(f"score network parameters missing: {missing[:5]} ... ({len(missing)} in total)")
        for name, shape in self._shapes.items():
            a = np.ascontiguousarray(params[name], dtype=np.float32)
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"{name}: shape {a.shape}, expected {shape}")
            t = torch.as_tensor(a)
            d = _lib.dl(t)
            _lib.check(self._lib.asep_ncsn_set_param(self._h, name.encode(), d.ptr))
            self._params[name] = a

    def prepare(self) -> None:
        _lib.check(self._lib.asep_ncsn_prepare(self._h))

    @property
    def variables(self) -> Dict[str, np.ndarray]:
        return dict(self._params)

    trainable_variables = variables

    def count_params(self) -> int:
        return int(sum(v.size for v in self._params.values()))

    def score(self, x: torch.Tensor, sigma_idx) -> torch.Tensor:
        x = torch.as_tensor(x)
        if x.dtype != torch.float32:
            x = x.float()
        x = x.to(self.device).contiguous()
        idx = torch.as_tensor(sigma_idx)
        if idx.ndim == 0:
            idx = idx.repeat(x.shape[0])
        idx = idx.to(device=self.device, dtype=torch.int32).contiguous()
        self._check_idx(idx, x.shape[0])
        out = torch.empty_like(x)
        dx, di, do = _lib.dl(x), _lib.dl(idx), _lib.dl(out)
        _lib.check(self._lib.asep_ncsn_forward(self._h, dx.ptr, di.ptr, do.ptr, _lib.stream_ptr()))
        return out

    def _check_idx(self, idx: torch.Tensor, n: int) -> None:
        """The kernels gather Embedding rows / noise levels by sigma_idx: an out-of-range index must not reach them."""
        if idx.numel() != n:
            raise ValueError(f"sigma_idx holds {idx.numel()} entries for a batch of {n}")
        if n and (int(idx.min()) < 0 or int(idx.max()) >= int(self.cfg.num_classes)):
            raise ValueError(f"sigma_idx out of range [0, {self.cfg.num_classes})")

    def __call__(self, inputs, training: bool = True) -> torch.Tensor:
        if isinstance(inputs, dict):
            return self.score(inputs["perturbed_X"], inputs["sigma_idx"])
        return self.score(inputs[0], inputs[1])

    @property
    def handle(self) -> ctypes.c_void_p:
        return self._h

    # ---- training surface (reference: train_ncsn.py:26-57; the arithmetic runs in libasep.so)
    def enable_training(self) -> None:
        """Moves every parameter into one flat fp32 device vector and builds the data-gradient weight images."""
        _lib.check(self._lib.asep_ncsn_enable_training(self._h))
        n = ctypes.c_int64()
        _lib.check(self._lib.asep_ncsn_num_trainable(self._h, ctypes.byref(n)))
        self.num_trainable = int(n.value)
        self._spans = {}
        for name in self._shapes:
            off, cnt = ctypes.c_int64(), ctypes.c_int64()
            _lib.check(self._lib.asep_ncsn_param_span(self._h, name.encode(), ctypes.byref(off), ctypes.byref(cnt)))
            self._spans[name] = (int(off.value), int(cnt.value))

    def train_grads(self, x: torch.Tensor, noise: torch.Tensor, sigma_idx, global_batch: int):
        """Gradient of the denoising-score-matching loss of this rank's samples w.r.t. the flat parameter vector, and
        the loss (both pre-divided by ``global_batch``, so the SUM over data-parallel ranks is the reference's
        ``compute_average_loss``).  ``noise``: standard-normal draws; perturbed_X = x + sigma[idx] * noise."""
        x = torch.as_tensor(x, dtype=torch.float32).to(self.device).contiguous()
        noise = torch.as_tensor(noise, dtype=torch.float32).to(self.device).contiguous()
        idx = torch.as_tensor(sigma_idx)
        if idx.ndim == 0 or idx.numel() == 1:        # train_ncsn.py:34-35: one level for the whole replica batch when C == 1
            idx = idx.reshape(-1)[:1].repeat(x.shape[0])
        idx = idx.to(device=self.device, dtype=torch.int32).contiguous()
        self._check_idx(idx, x.shape[0])
        grads = torch.empty((self.num_trainable,), dtype=torch.float32, device=self.device)
        loss = torch.empty((1,), dtype=torch.float32, device=self.device)
        d = [_lib.dl(t) for t in (x, noise, idx, grads, loss)]
        _lib.check(self._lib.asep_ncsn_train_grads(self._h, d[0].ptr, d[1].ptr, d[2].ptr, int(global_batch), d[3].ptr, d[4].ptr,
                                                   _lib.stream_ptr()))
        return grads, loss

    def apply_gradients(self, grads: torch.Tensor, optimizer: dict) -> None:
        """Keras Adam (train_utils.py:27-28) on the flat vector; the tile images are rebuilt on the device."""
        if optimizer.get("kind", "adam") != "adam":
            raise ValueError("the score networks train with Adam (train_ncsn.py:405-406 default)")
        d = _lib.dl(grads)
        _lib.check(self._lib.asep_ncsn_adam_step(self._h, d.ptr, float(optimizer["lr"]), float(optimizer["beta1"]),
                                                 float(optimizer["beta2"]), float(optimizer["eps"]), _lib.stream_ptr()))

    def get_flat(self) -> torch.Tensor:
        theta = torch.empty((self.num_trainable,), dtype=torch.float32, device=self.device)
        d = _lib.dl(theta)
        _lib.check(self._lib.asep_ncsn_get_flat(self._h, d.ptr, _lib.stream_ptr()))
        return theta

    def set_flat(self, theta: torch.Tensor) -> None:
        theta = torch.as_tensor(theta, dtype=torch.float32).to(self.device).contiguous()
        d = _lib.dl(theta)
        _lib.check(self._lib.asep_ncsn_set_flat(self._h, d.ptr, _lib.stream_ptr()))

    def unflatten(self, flat: torch.Tensor) -> Dict[str, np.ndarray]:
        """Splits a flat vector (parameters or gradients) into the named tensors of ``ncsn_param_shapes``."""
        host = flat.detach().cpu().numpy()
        return {name: host[off:off + cnt].reshape(self._shapes[name]).copy() for name, (off, cnt) in self._spans.items()}

    def sync_host(self) -> None:
        """``variables`` <- the trained values (checkpoint writers read them)."""
        self._params = self.unflatten(self.get_flat())
