"""Host-side mirror of the reference's ``flow_models/flow_glow.py``.

Two ways to build the multi-scale Glow bijector, behind the reference's class names:

* **fused path** (what the scripts use): ``GlowBijector_{2,3,4}blocks(K, event_shape,
  ShiftAndLogScaleConvNet, n_filters, minibatch)`` creates ONE ``asep_glow_t`` handle in libasep.so
  (flow_glow.py:80-329 restated in CUDA: every step is one tcgen05 coupling-network launch plus one
  fused element-wise launch) with the reference's random init and its data-dependent ActNorm init.
* **composable path** (what the reference's unit tests use, unittest_flow_models.py:76-83): when
  ``shift_and_log_scale_layer`` is any other factory ``(event_shape, **kw) -> callable(xb)->(log_s,t)``
  the bijector is assembled from the single-bijector kernels (`ActNorm`, `Invertible1x1Conv`,
  `AffineCouplingLayerSplit`, `Squeeze`) exactly as the reference chains them.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from ..config import GlowConfig
from ..glow import Glow
from ..weights import init_glow_params
from .flow_tfp_bijectors import (ActNorm, AffineCouplingLayerSplit, Bijector, Chain, Invertible1x1Conv, Squeeze)


class ShiftAndLogScaleConvNet:
    """Marker for the coupling network of flow_tfk_layers.py:31-84.  Passing this class as
    ``shift_and_log_scale_layer`` selects the fused CUDA path; it is never called in Python."""

    def __init__(self, event_shape, n_filters=512, **kwargs):
        raise TypeError("ShiftAndLogScaleConvNet runs inside libasep.so; build it through GlowBijector_*blocks")


class GlowStep(Bijector):
    """reference: flow_glow.py:9-31 -- Chain([coupling, inv1x1, actnorm])."""

    def __init__(self, event_shape, shift_and_log_scale_layer, minibatch, name="glowStep", **kwargs):
        super().__init__(name=name)
        self.actnorm = ActNorm(event_shape, minibatch, name="ActNorm")
        self.inv1x1conv = Invertible1x1Conv(event_shape, name="inv1x1conv")
        self.coupling_layer = AffineCouplingLayerSplit(event_shape, shift_and_log_scale_layer, name="couplingLayer",
                                                       **kwargs)
        self.bijector = Chain([self.coupling_layer, self.inv1x1conv, self.actnorm])

    def _forward(self, x):
        return self.bijector.forward(x)

    def _inverse(self, y):
        return self.bijector.inverse(y)

    def _forward_log_det_jacobian(self, x):
        return self.bijector.forward_log_det_jacobian(x, event_ndims=3)


class GlowBlock(Bijector):
    """reference: flow_glow.py:34-77 -- Chain(glow_steps + [squeeze]) (steps run K-1..0 after the squeeze)."""

    def __init__(self, K, event_shape, shift_and_log_scale_layer, minibatch, name="glowBlock", **kwargs):
        super().__init__(name=name)
        self.squeeze = Squeeze(event_shape)
        self.event_shape_out = self.squeeze.event_shape_out
        minibatch_updated = self.squeeze.forward(minibatch)
        self.glow_steps = []
        for k in range(K):
            step = GlowStep(self.event_shape_out, shift_and_log_scale_layer, minibatch_updated,
                            name="glowStep_" + str(k), **kwargs)
            minibatch_updated = step.forward(minibatch_updated)
            self.glow_steps.append(step)
        self.chain = self.glow_steps + [self.squeeze]
        self.bijector = Chain(self.chain)

    def _forward(self, x):
        return self.bijector.forward(x)

    def _inverse(self, y):
        return self.bijector.inverse(y)

    def _forward_log_det_jacobian(self, x):
        return self.bijector.forward_log_det_jacobian(x, event_ndims=3)


class _GlowBijectorBase(Bijector):
    L = 0

    def __init__(self, K, event_shape, shift_and_log_scale_layer, n_filters, minibatch, name=None, seed=0, **kwargs):
        super().__init__(name=name or f"GlowBijector_{self.L}blocks")
        self.H, self.W, self.C = event_shape
        self.K = K
        self.fused = shift_and_log_scale_layer is ShiftAndLogScaleConvNet
        if self.fused:
            # identity SpecPreprocessing (x' = x - (-0.5) - 0.5, zero log-det) so the handle is the bare bijector
            self.cfg = GlowConfig(H=self.H, W=self.W, C=self.C, L=self.L, K=K, n_filters=n_filters, learntop=False,
                                  minval=-0.5, maxval=0.5)
            params = init_glow_params(self.cfg, seed=seed, mode="faithful")
            precision = _lib.PREC_BF16 if n_filters == 512 else _lib.PREC_FP32
            self.model = Glow(self.cfg, params, precision=precision)
            self.model.init_actnorm(torch.as_tensor(minibatch, dtype=torch.float32))
        else:
            self._build_composable(K, event_shape, shift_and_log_scale_layer, n_filters, minibatch, **kwargs)

    # ---- composable path: literal transcription of the constructors incl. the raw-minibatch quirk (Q7)
    def _build_composable(self, K, event_shape, layer, n_filters, minibatch, **kwargs):
        minibatch = torch.as_tensor(minibatch, dtype=torch.float32).cuda()
        self.blocks = []
        shape = list(event_shape)
        carried = minibatch
        for b in range(self.L):
            # flow_glow.py:162-165,171-174: the 3/4-block classes hand the RAW minibatch to blocks 2..;
            # Squeeze's -1 reshape re-tiles it.  The 2-block class passes the carried batch (:97-100).
            src = carried if (b == 0 or self.L == 2) else minibatch
            blk = GlowBlock(K, shape, layer, src.reshape(-1, *shape), name=f"glowBlock{b + 1}", n_filters=n_filters,
                            **kwargs)
            self.blocks.append(blk)
            H, W, C = blk.event_shape_out
            if b + 1 < self.L:
                out = blk.forward(carried)
                carried = out[..., C // 2:]
                shape = [H, W, C // 2]

    def _latent_dims(self):
        s = 1 << self.L
        return self.H // s, self.W // s

    def _forward(self, x):
        if self.fused:
            return self.model.forward(x)
        Hl, Wl = self._latent_dims()
        zs, h = [], torch.as_tensor(x, dtype=torch.float32).cuda()
        for b, blk in enumerate(self.blocks):
            o = blk.forward(h)
            if b + 1 < self.L:
                C = o.shape[-1]
                zs.append(o[..., : C // 2].reshape(o.shape[0], Hl, Wl, -1))      # plain reshape, flow_glow.py:179
                h = o[..., C // 2:].contiguous()
            else:
                zs.append(o)
        return torch.cat(zs, dim=-1)

    def _inverse(self, y):
        if self.fused:
            return self.model.inverse(y)
        y = torch.as_tensor(y, dtype=torch.float32).cuda()
        parts, rest = [], y
        for _ in range(self.L - 1):
            C = rest.shape[-1]
            parts.append(rest[..., : C // 2])
            rest = rest[..., C // 2:]
        h = self.blocks[-1].inverse(rest.contiguous())
        for b in reversed(range(self.L - 1)):
            Hb, Wb, Cb = self.blocks[b].event_shape_out
            zb = parts[b].reshape(y.shape[0], Hb, Wb, Cb // 2)
            h = self.blocks[b].inverse(torch.cat([zb, h], dim=-1).contiguous())
        return h

    def _forward_log_det_jacobian(self, x):
        if self.fused:
            return self.model.forward_log_det_jacobian(x)
        total, h = None, torch.as_tensor(x, dtype=torch.float32).cuda()
        for b, blk in enumerate(self.blocks):
            ld = blk.forward_log_det_jacobian(h, event_ndims=3)
            total = ld if total is None else total + ld
            if b + 1 < self.L:
                o = blk.forward(h)
                h = o[..., o.shape[-1] // 2:].contiguous()
        return total


class GlowBijector_2blocks(_GlowBijectorBase):
    """reference: flow_glow.py:80-142."""
    L = 2


class GlowBijector_3blocks(_GlowBijectorBase):
    """reference: flow_glow.py:145-225."""
    L = 3


class GlowBijector_4blocks(_GlowBijectorBase):
    """reference: flow_glow.py:228-329."""
    L = 4
