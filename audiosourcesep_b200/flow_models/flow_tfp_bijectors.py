"""Host-side mirror of the reference's elementary bijectors (flow_models/flow_tfp_bijectors.py).

Same class names, constructor arguments and the TFP ``Bijector`` protocol
(``forward`` / ``inverse`` / ``forward_log_det_jacobian`` / ``inverse_log_det_jacobian`` with
``event_ndims=3``); the arithmetic runs in libasep.so on the B200.  Tensors are float32 NHWC
``torch.Tensor`` on the CUDA device.
"""
from __future__ import annotations

import math
from typing import Callable, Sequence

import numpy as np
import scipy.linalg
import torch

from .. import ops


class Bijector:
    """Minimal stand-in for ``tfp.bijectors.Bijector`` (only what the reference's code paths use)."""

    def __init__(self, forward_min_event_ndims=3, name="bijector"):
        self.forward_min_event_ndims = forward_min_event_ndims
        self.name = name

    def forward(self, x):
        return self._forward(x)

    def inverse(self, y):
        return self._inverse(y)

    def forward_log_det_jacobian(self, x, event_ndims=3):
        return self._forward_log_det_jacobian(x)

    def inverse_log_det_jacobian(self, y, event_ndims=3):
        return -self._forward_log_det_jacobian(self._inverse(y))


class Chain(Bijector):
    """``tfb.Chain``: the list is applied RIGHT-TO-LEFT in ``forward``."""

    def __init__(self, bijectors: Sequence[Bijector], name="chain"):
        super().__init__(name=name)
        self.bijectors = list(bijectors)

    def _forward(self, x):
        for b in reversed(self.bijectors):
            x = b.forward(x)
        return x

    def _inverse(self, y):
        for b in self.bijectors:
            y = b.inverse(y)
        return y

    def _forward_log_det_jacobian(self, x):
        total = None
        for b in reversed(self.bijectors):
            ld = b.forward_log_det_jacobian(x, event_ndims=3)
            total = ld if total is None else total + ld
            x = b.forward(x)
        return total


class Invert(Bijector):
    def __init__(self, bijector: Bijector, name="invert"):
        super().__init__(name=name)
        self.bijector = bijector

    def _forward(self, x):
        return self.bijector.inverse(x)

    def _inverse(self, y):
        return self.bijector.forward(y)

    def _forward_log_det_jacobian(self, x):
        return self.bijector.inverse_log_det_jacobian(x, event_ndims=3)

    def inverse_log_det_jacobian(self, y, event_ndims=3):
        return self.bijector.forward_log_det_jacobian(y, event_ndims=3)


class AffineCouplingLayerSplit(Bijector):
    """reference: flow_tfp_bijectors.py:124-153.  ``shift_and_log_scale_layer(event_shape, **kwargs)``
    must return a callable ``xb -> (log_s, t)`` -- the plug-in seam the reference's tests use."""

    def __init__(self, event_shape, shift_and_log_scale_layer: Callable, name="AffineCouplingLayer", **kwargs):
        super().__init__(name=name)
        self.H, self.W, self.C = event_shape
        assert self.C % 2 == 0
        self.shift_and_log_scale_fn = shift_and_log_scale_layer([self.H, self.W, self.C // 2], **kwargs)

    def _raw(self, xb):
        log_s, t = self.shift_and_log_scale_fn(xb)
        # the kernel applies tanh itself; callables return the activated log-scale like the reference
        raw = torch.atanh(torch.clamp(torch.as_tensor(log_s, dtype=torch.float32), -1 + 1e-7, 1 - 1e-7))
        return torch.cat([raw.to(xb.device), torch.as_tensor(t, dtype=torch.float32).to(xb.device)], dim=-1)

    def _forward(self, x):
        x = torch.as_tensor(x, dtype=torch.float32).cuda()
        return ops.coupling(x, self._raw(x[..., self.C // 2:]))[0]

    def _inverse(self, y):
        y = torch.as_tensor(y, dtype=torch.float32).cuda()
        return ops.coupling(y, self._raw(y[..., self.C // 2:]), inverse=True)[0]

    def _forward_log_det_jacobian(self, x):
        x = torch.as_tensor(x, dtype=torch.float32).cuda()
        return ops.coupling(x, self._raw(x[..., self.C // 2:]))[1]


class Squeeze(Bijector):
    """reference: flow_tfp_bijectors.py:156-199."""

    def __init__(self, event_shape_in, name="Squeeze"):
        super().__init__(name=name)
        H, W, C = event_shape_in
        self.H, self.W, self.C = H, W, C
        assert H % 2 == 0
        assert W % 2 == 0
        self.event_shape_out = (H // 2, W // 2, 4 * C)

    def _forward(self, x):
        x = torch.as_tensor(x, dtype=torch.float32)
        return ops.squeeze(x.reshape(-1, self.H, self.W, self.C))

    def _inverse(self, y):
        y = torch.as_tensor(y, dtype=torch.float32)
        return ops.squeeze(y.reshape(-1, self.H // 2, self.W // 2, 4 * self.C), inverse=True)

    def _forward_log_det_jacobian(self, x):
        return torch.zeros(x.shape[0], device="cuda")


class ActNorm(Bijector):
    """reference: flow_tfp_bijectors.py:202-253 (data-dependent init from ``minibatch``)."""

    def __init__(self, event_shape, minibatch, normalize="channel", name="ActNorm"):
        super().__init__(name=name)
        self.H, self.W, self.C = event_shape
        minibatch = torch.as_tensor(minibatch, dtype=torch.float32)
        _, mh, mw, mc = minibatch.shape
        assert self.H == mh
        assert self.W == mw
        assert self.C == mc
        if normalize != "channel":
            raise NotImplementedError("only normalize='channel' is on the separation hot path")
        mean = minibatch.double().mean(dim=(0, 1, 2))
        std = minibatch.double().std(dim=(0, 1, 2), unbiased=False) + 1e-8
        self.log_scale = torch.log(1.0 / std).float().cuda()
        self.shift = (-mean / std).float().cuda()

    def _forward(self, x):
        return ops.actnorm(x, self.log_scale, self.shift)

    def _inverse(self, y):
        return ops.actnorm(y, self.log_scale, self.shift, inverse=True)

    def _forward_log_det_jacobian(self, x):
        log_det = self.H * self.W * self.log_scale.sum()
        return log_det.repeat(x.shape[0])


class Invertible1x1Conv(Bijector):
    """reference: flow_tfp_bijectors.py:256-322 (QR -> LU parameterisation)."""

    def __init__(self, event_shape, name="inv1x1conv", seed=None):
        super().__init__(name=name)
        self.height, self.width, self.C = event_shape
        rng = np.random.default_rng(seed)
        np_w = np.linalg.qr(rng.standard_normal((self.C, self.C)))[0]
        np_p, np_l, np_u = scipy.linalg.lu(np_w)
        np_s = np.diag(np_u)
        self.P = np_p
        self.P_inv = np.linalg.inv(np_p)
        self.Sign_s = np.sign(np_s)
        self.L = np_l
        self.Log_s = np.log(np.abs(np_s))
        self.U = np.triu(np_u, k=1)
        self.l_mask = np.tril(np.ones((self.C, self.C)), -1)

    def _w(self):
        L = self.L * self.l_mask + np.eye(self.C)
        u = self.U * self.l_mask.T + np.diag(self.Sign_s * np.exp(self.Log_s))
        return L, u

    def _forward(self, x):
        L, u = self._w()
        return ops.inv1x1(x, torch.as_tensor(self.P @ (L @ u), dtype=torch.float32))

    def _inverse(self, y):
        L, u = self._w()
        w_inv = np.linalg.inv(u) @ (np.linalg.inv(L) @ self.P_inv)
        return ops.inv1x1(y, torch.as_tensor(w_inv, dtype=torch.float32))

    def _forward_log_det_jacobian(self, x):
        log_det = self.height * self.width * float(np.sum(self.Log_s))
        return torch.full((x.shape[0],), log_det, device="cuda")


class SpecPreprocessing(Bijector):
    """reference: flow_tfp_bijectors.py:364-396 (the configs use ``use_logit=False``)."""

    def __init__(self, minval, maxval, alpha=1e-10, use_logit=True, name="SpecPreprocessing", **kwargs):
        super().__init__(name=name)
        self.maxval, self.minval, self.alpha, self.use_logit = maxval, minval, alpha, use_logit

    def _forward(self, x):
        x = torch.as_tensor(x, dtype=torch.float32).cuda()
        x = (x - self.minval) / (self.maxval - self.minval)
        if self.use_logit:
            x = (1.0 - 2.0 * self.alpha) * x + self.alpha
            return torch.log(x) - torch.log(1.0 - x)
        return x - 0.5

    def _inverse(self, y):
        y = torch.as_tensor(y, dtype=torch.float32).cuda()
        if self.use_logit:
            y = (torch.sigmoid(y) - self.alpha) / (1.0 - 2.0 * self.alpha)
        else:
            y = y + 0.5
        return y * (self.maxval - self.minval) + self.minval

    def _forward_log_det_jacobian(self, x):
        x = torch.as_tensor(x, dtype=torch.float32).cuda()
        n = x.shape[0]
        d = x[0].numel()
        log_det = torch.full((n,), d * math.log(1.0 / (self.maxval - self.minval)), device=x.device)
        if self.use_logit:
            xs = (1.0 - 2.0 * self.alpha) * (x - self.minval) / (self.maxval - self.minval) + self.alpha
            log_det = log_det + (-torch.log(xs) - torch.log(1.0 - xs) + math.log(1.0 - 2.0 * self.alpha)).reshape(n, -1).sum(1)
        return log_det
