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
Pins: parameter-count known answer 67,464,769 (trained_ncsn/.../out.log:35) -- reproduced both by the product's
parameter inventory and by this module's own walk (``count_params_by_walk``, which shares no table with the product); an independent numpy re-implementation of the primitives
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


def count_params_by_walk(version: str, ngf: int, num_classes: int, C: int = 1) -> int:
    """Trainable parameters counted by walking the reference constructors (score_network.py:181-272,
    score_network_v2.py:174-250) -- independent of audiosourcesep_b200/weights.py.  v1, ngf=192, 10 classes must give
    the 67,464,769 printed by the reference's own training log (trained_ncsn/.../out.log:35)."""
    v1 = version == "v1"
    conv = lambda k, ci, co, bias: k * k * ci * co + (co if bias else 0)
    # ConditionalInstanceNorm2dPlus: Embedding(num_classes, 3C) + tfa InstanceNormalization gamma/beta;
    # InstanceNorm2dPlus: alpha, gamma, beta vectors + InstanceNormalization gamma/beta
    norm = (lambda c: num_classes * 3 * c + 2 * c) if v1 else (lambda c: 3 * c + 2 * c)
    inner_norm = norm if v1 else (lambda c: 0)                  # v2 RCU / MSF / CRP have no norm layers
    n = conv(3, C, ngf, True) + norm(ngf) + conv(3, ngf, C, True)           # begin_conv, normalizer, end_conv

    def res(cin, cout, resample=None, dilation=None):
        t = norm(cin)
        if resample == "down":
            t += conv(3, cin, cin, dilation is not None) + norm(cin) + conv(3, cin, cout, True)
            t += conv(3 if dilation is not None else 1, cin, cout, True)
        else:
            t += conv(3, cin, cout, dilation is not None) + norm(cout) + conv(3, cout, cout, dilation is not None)
            # Keras builds lazily: a shortcut that call() never uses (input_dim == output_dim) creates no variables
            if cin != cout:
                t += conv(3, cin, cin if dilation is not None else cout, dilation is not None)
        return t

    n += 2 * res(ngf, ngf) + res(ngf, 2 * ngf, "down") + res(2 * ngf, 2 * ngf)
    n += res(2 * ngf, 2 * ngf, "down", 2) + res(2 * ngf, 2 * ngf, None, 2)
    n += res(2 * ngf, 2 * ngf, "down", 4) + res(2 * ngf, 2 * ngf, None, 4)

    def rcu(c, blocks):
        return blocks * 2 * (inner_norm(c) + conv(3, c, c, False))

    def refine(in_planes, feat, start=False, end=False):
        t = sum(rcu(c, 2) for c in in_planes)
        if not start:
            t += sum(inner_norm(c) + conv(3, c, feat, True) for c in in_planes)
        t += 2 * (inner_norm(feat) + conv(3, feat, feat, False))
        return t + rcu(feat, 3 if end else 1)

    n += refine([2 * ngf], 2 * ngf, start=True) + refine([2 * ngf, 2 * ngf], 2 * ngf)
    n += refine([2 * ngf, 2 * ngf], ngf) + refine([ngf, ngf], ngf, end=True)
    return n


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
    def conv(self, x, name, dilation=1, bias=None, ksize=3):
        """``bias`` / ``ksize``: what the reference constructor builds for this layer (use_bias, kernel_size); checked
        against the parameter dictionary so that a wrong bias flag or kernel size in the product's inventory
        (audiosourcesep_b200/weights.py) cannot pass as common mode."""
        k = self.p[name + "/kernel"]
        assert k.shape[0] == k.shape[1] == ksize, (name, tuple(k.shape), ksize)
        if bias is not None:
            assert ((name + "/bias") in self.p) == bias, f"{name}: use_bias should be {bias}"
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

    def res_block(self, x, y, name, cin, cout, resample=None, dilation=None):
        """(Conditional)ResidualBlock, branch by branch as the constructor and call() of
        score_network.py:121-178 / score_network_v2.py:110-171 write it.  The oracle OWNS this description:
        it does not read the product's layout table (audiosourcesep_b200/weights.py)."""
        d = dilation if dilation is not None else 1
        pool = resample == "down" and dilation is None          # conv2 / shortcut are Sequential([Conv2D, AveragePooling2D(2)])
        o = self.norm(x, y, name + "/norm1")
        o = F.elu(o)
        # use_bias: Keras default True everywhere except the undilated convs written with use_bias=False
        # (score_network.py:139 conv1 of the pooled 'down' block; :156-159 all three convs of the plain block)
        dil = dilation is not None
        o = self.conv(o, name + "/conv1", d, bias=dil)           # down: Conv2D(input_dim); else Conv2D(output_dim)
        assert o.shape[-1] == (cin if resample == "down" else cout)
        o = self.norm(o, y, name + "/norm2")
        o = F.elu(o)
        o = self.conv(o, name + "/conv2", d, bias=dil or resample == "down")
        assert o.shape[-1] == cout
        if pool:
            o = avg_pool2(o)
        if cout == cin and resample is None:                     # score_network.py:173-174
            sc = x
        else:
            # down + dilation: 3x3 dilated; down without dilation: 1x1 then AveragePooling2D(2); otherwise 3x3
            sc = self.conv(x, name + "/shortcut", d, bias=dil or resample == "down", ksize=1 if pool else 3)
            if pool:
                sc = avg_pool2(sc)
        return sc + o

    def rcu(self, x, y, prefix, n_blocks, n_stages):
        """score_network.py:47-54: norm -> conv, NO activation (quirk Q8); v2: conv only."""
        for i in range(n_blocks):
            residual = x
            for j in range(n_stages):
                if self.v1:
                    x = self.norm(x, y, f"{prefix}/norm_{i + 1}_{j + 1}")
                x = self.conv(x, f"{prefix}/conv_{i + 1}_{j + 1}", bias=False)           # score_network.py:38
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
            path = self.conv(path, f"{prefix}/conv_{i + 1}", bias=False)                    # score_network.py:13
            x = x + path
        return x

    def msf(self, xs, y, prefix, shape):
        sums = None
        for i, xi in enumerate(xs):
            h = self.norm(xi, y, f"{prefix}/norm_{i + 1}") if self.v1 else xi
            h = self.conv(h, f"{prefix}/conv_{i + 1}", bias=True)                           # score_network.py:66
            h = resize_bilinear(h, shape)
            sums = h if sums is None else sums + h
        return sums

    def refine(self, xs, y, name, shape, end=False):
        """(Cond)RefineBlock (score_network.py:82-118): adapters RCU(2, 2) per input, MSF when there are several inputs,
        CRP(2 stages), output RCU(3 if end else 1, 2)."""
        hs = [self.rcu(xi, y, f"{name}/RCU_{i + 1}", 2, 2) for i, xi in enumerate(xs)]
        h = self.msf(hs, y, f"{name}/MSF", shape) if len(xs) > 1 else hs[0]
        h = self.crp(h, y, f"{name}/CRP")
        return self.rcu(h, y, f"{name}/RCU_output", 3 if end else 1, 2)

    # ---- model([x, sigma_idx], training=True)
    def score(self, x, sigma_idx):
        x = torch.as_tensor(np.asarray(x), dtype=self.dtype) if not torch.is_tensor(x) else x.to(self.dtype)
        y = torch.as_tensor(np.asarray(sigma_idx), dtype=torch.long)
        if y.ndim == 0:
            y = y.repeat(x.shape[0])
        if self.v1:
            x = 2.0 * x - 1.0                               # score_network.py:277-278
        out = self.conv(x, "begin_conv", bias=True)
        ngf = self.cfg.ngf
        # score_network.py:238-260 (v2: score_network_v2.py:216-238): res1 at full resolution, res2 halves it through
        # AveragePooling2D, res3 / res4 are 'down' blocks that only dilate (2, 4) -- no further down-sampling
        l1 = self.res_block(out, y, "Res1_1", ngf, ngf)
        l1 = self.res_block(l1, y, "Res1_2", ngf, ngf)
        l2 = self.res_block(l1, y, "Res2_1", ngf, 2 * ngf, resample="down")
        l2 = self.res_block(l2, y, "Res2_2", 2 * ngf, 2 * ngf)
        l3 = self.res_block(l2, y, "Res3_1", 2 * ngf, 2 * ngf, resample="down", dilation=2)
        l3 = self.res_block(l3, y, "Res3_2", 2 * ngf, 2 * ngf, dilation=2)
        l4 = self.res_block(l3, y, "Res4_1", 2 * ngf, 2 * ngf, resample="down", dilation=4)
        l4 = self.res_block(l4, y, "Res4_2", 2 * ngf, 2 * ngf, dilation=4)
        # score_network.py:262-265, :285-288
        ref1 = self.refine([l4], y, "refine1", l4.shape[1:3])
        ref2 = self.refine([l3, ref1], y, "refine2", l3.shape[1:3])
        ref3 = self.refine([l2, ref2], y, "refine3", l2.shape[1:3])
        o = self.refine([l1, ref3], y, "refine4", l1.shape[1:3], end=True)
        o = self.norm(o, y, "normalizer")
        o = F.elu(o)
        o = self.conv(o, "end_conv", bias=True)
        if not self.v1:
            o = o / self.sigmas[y].reshape(-1, 1, 1, 1)     # score_network_v2.py:275-276
        return o
