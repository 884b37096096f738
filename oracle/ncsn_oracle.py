"""CPU ORACLE (test infrastructure, NOT a product path) -- NCSN v1 / v2 score networks.

torch-CPU restatement (float64 or float32) of
  ncsn/score_network.py:7-296      CondCRPBlock, CondRCUBlock, CondMSFBlock, CondRefineBlock,
                                   ConditionalResidualBlock, ConditionalInstanceNorm2dPlus,
                                   CondRefineNetDilated
  ncsn/score_network_v2.py:6-283   CRPBlock, RCUBlock, MSFBlock, RefineBlock, ResidualBlock,
                                   InstanceNorm2dPlus, RefineNetDilated
  ncsn/utils.py:41-64              model([perturbed_X, sigma_idx]) call contract
Third-party semantics restated from their documented behaviour (TensorFlow 2.2.0 / Keras,
tensorflow-addons 0.10.0, neither installable here): Conv2D padding='same' incl. dilation,
AveragePooling2D 'same' (divisor = in-bounds taps), MaxPooling2D 'same' (-inf padding),
AveragePooling2D(2) valid stride 2, tf.image.resize bilinear half-pixel, ELU(alpha=1),
tfa.InstanceNormalization(epsilon=1e-3, population variance, affine).
Pins: parameter-count known answer 67,464,769 (trained_ncsn/.../out.log:35) via the shared
parameter inventory; an independent numpy re-implementation of the primitives
(tests/test_oracle_ncsn.py).  Network outputs have no reference golden vector:
"parity unpinned".
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from .glow_oracle import conv2d_same

IN_EPS = 1e-3      # tfa InstanceNormalization default epsilon
PLUS_EPS = 1e-5    # score_network.py:205


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def avg_pool5_same(x):
    return _nhwc(F.avg_pool2d(_nchw(x), 5, stride=1, padding=2, count_include_pad=False))


def max_pool5_same(x):
    return _nhwc(F.max_pool2d(_nchw(x), 5, stride=1, padding=2))


def avg_pool2(x):
    return _nhwc(F.avg_pool2d(_nchw(x), 2))


def resize_bilinear(x, shape):
    if tuple(x.shape[1:3]) == tuple(shape):
        return x
    return _nhwc(F.interpolate(_nchw(x), size=tuple(shape), mode="bilinear", align_corners=False))


class NCSNOracle:
    def __init__(self, cfg, params: Dict[str, np.ndarray], sigmas=None, dtype=torch.float64,
                 bf16_operands: bool = False):
        """``bf16_operands=True`` rounds every tensor-core convolution operand (activations and kernels) to
        bfloat16 before an exact fp32/fp64 accumulation -- the arithmetic contract of the CUDA path
        (conv_tc.cu); used to separate "the kernel is wrong" from "bf16 operands are less precise"."""
        self.cfg = cfg
        self.dtype = dtype
        self.bf16_operands = bf16_operands
        self.p = {k: torch.as_tensor(np.asarray(v), dtype=dtype) for k, v in params.items()}
        self.v1 = cfg.version == "v1"
        self.sigmas = None if sigmas is None else torch.as_tensor(np.asarray(sigmas), dtype=dtype)

    # ---- layers
    def conv(self, x, name, dilation=1):
        k = self.p[name + "/kernel"]
        if self.bf16_operands and name != "begin_conv":
            x = x.to(torch.bfloat16).to(self.dtype)
            if name != "end_conv":                     # the 1-channel end convolution keeps fp32 weights
                k = k.to(torch.bfloat16).to(self.dtype)
        return conv2d_same(x, k, self.p.get(name + "/bias"), dilation)

    def norm(self, x, y, name):
        """score_network.py:203-221 / score_network_v2.py:188-199."""
        p = self.p
        C = x.shape[-1]
        means = x.mean(dim=(1, 2), keepdim=True)
        m = means.mean(dim=-1, keepdim=True)
        v = means.var(dim=-1, unbiased=False, keepdim=True)
        means = (means - m) / torch.sqrt(v + PLUS_EPS)
        mu = x.mean(dim=(1, 2), keepdim=True)
        var = x.var(dim=(1, 2), unbiased=False, keepdim=True)
        h = (x - mu) * torch.rsqrt(var + IN_EPS) * p[name + "/in_gamma"] + p[name + "/in_beta"]
        if self.v1:
            e = p[name + "/embed"][y]                      # [N, 3C]
            gamma, alpha, beta = e[:, :C], e[:, C:2 * C], e[:, 2 * C:]
            gamma, alpha, beta = (t.reshape(-1, 1, 1, C) for t in (gamma, alpha, beta))
        else:
            gamma, alpha, beta = p[name + "/gamma"], p[name + "/alpha"], p[name + "/beta"]
        return gamma * h + means * alpha + beta

    def res_block(self, x, y, spec):
        n = spec["name"]
        o = self.norm(x, y, n + "/norm1")
        o = F.elu(o)
        o = self.conv(o, n + "/conv1", spec["conv1"][4])
        o = self.norm(o, y, n + "/norm2")
        o = F.elu(o)
        o = self.conv(o, n + "/conv2", spec["conv2"][4])
        if spec["pool"]:
            o = avg_pool2(o)
        if spec["shortcut"] is None:
            sc = x
        else:
            sc = self.conv(x, n + "/shortcut", spec["shortcut"][4])
            if spec["pool"]:
                sc = avg_pool2(sc)
        return sc + o

    def rcu(self, x, y, prefix, n_blocks, n_stages):
        """score_network.py:47-54: norm -> conv, NO activation (quirk Q8); v2: conv only."""
        for i in range(n_blocks):
            residual = x
            for j in range(n_stages):
                if self.v1:
                    x = self.norm(x, y, f"{prefix}/norm_{i + 1}_{j + 1}")
                x = self.conv(x, f"{prefix}/conv_{i + 1}_{j + 1}")
            x = x + residual
        return x

    def crp(self, x, y, prefix):
        x = F.elu(x)
        path = x
        for i in range(2):
            if self.v1:
                path = self.norm(path, y, f"{prefix}/norm_{i + 1}")
                path = avg_pool5_same(path)
            else:
                path = max_pool5_same(path)
            path = self.conv(path, f"{prefix}/conv_{i + 1}")
            x = x + path
        return x

    def msf(self, xs, y, prefix, shape):
        sums = None
        for i, xi in enumerate(xs):
            h = self.norm(xi, y, f"{prefix}/norm_{i + 1}") if self.v1 else xi
            h = self.conv(h, f"{prefix}/conv_{i + 1}")
            h = resize_bilinear(h, shape)
            sums = h if sums is None else sums + h
        return sums

    def refine(self, xs, y, r, shape):
        n = r["name"]
        hs = [self.rcu(xi, y, f"{n}/RCU_{i + 1}", 2, 2) for i, xi in enumerate(xs)]
        h = self.msf(hs, y, f"{n}/MSF", shape) if len(xs) > 1 else hs[0]
        h = self.crp(h, y, f"{n}/CRP")
        return self.rcu(h, y, f"{n}/RCU_output", 3 if r["end"] else 1, 2)

    # ---- model([x, sigma_idx], training=True)
    def score(self, x, sigma_idx):
        from audiosourcesep_b200.weights import ncsn_layout
        x = torch.as_tensor(np.asarray(x), dtype=self.dtype) if not torch.is_tensor(x) else x.to(self.dtype)
        y = torch.as_tensor(np.asarray(sigma_idx), dtype=torch.long)
        if y.ndim == 0:
            y = y.repeat(x.shape[0])
        res, refine = ncsn_layout(self.cfg.ngf)
        if self.v1:
            x = 2.0 * x - 1.0                               # score_network.py:277-278
        out = self.conv(x, "begin_conv")
        layers = []
        h = out
        for i, spec in enumerate(res):
            h = self.res_block(h, y, spec)
            if i % 2 == 1:
                layers.append(h)
        l1, l2, l3, l4 = layers
        ref1 = self.refine([l4], y, refine[0], l4.shape[1:3])
        ref2 = self.refine([l3, ref1], y, refine[1], l3.shape[1:3])
        ref3 = self.refine([l2, ref2], y, refine[2], l2.shape[1:3])
        o = self.refine([l1, ref3], y, refine[3], l1.shape[1:3])
        o = self.norm(o, y, "normalizer")
        o = F.elu(o)
        o = self.conv(o, "end_conv")
        if not self.v1:
            o = o / self.sigmas[y].reshape(-1, 1, 1, 1)     # score_network_v2.py:275-276
        return o
