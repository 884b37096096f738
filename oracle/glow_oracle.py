"""CPU ORACLE (test infrastructure, NOT a product path) -- Glow prior.

A restatement in torch-CPU (float64 "truth" or float32 "as-TF") of the reference's Glow
bijector chain, written from the cited lines; it never runs on the GPU and the product
(`audiosourcesep_b200`) never imports it.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this module.

Parity pin status: the closed-form cases of the reference's own unit tests
(unittest_flow_models.py:122-186: coupling 4*log2, ActNorm 4*log2, fldj == -ildj, round
trips) are reproduced in tests/test_oracle_glow.py.  Full-model ``log_prob`` /
``grad_log_prob`` have NO reference golden vector (TensorFlow 2.2 / TFP 0.9 cannot be
installed here): for those this oracle is "parity unpinned" (see DESIGN.md).

Reference files restated (all under /root/reference):
  flow_models/flow_tfp_bijectors.py:124-153  AffineCouplingLayerSplit
  flow_models/flow_tfp_bijectors.py:156-199  Squeeze
  flow_models/flow_tfp_bijectors.py:202-253  ActNorm (+ data-dependent init :222-240)
  flow_models/flow_tfp_bijectors.py:256-322  Invertible1x1Conv
  flow_models/flow_tfp_bijectors.py:364-396  SpecPreprocessing
  flow_models/flow_tfk_layers.py:31-84       ShiftAndLogScaleConvNet
  flow_models/flow_glow.py:9-329             GlowStep / GlowBlock / GlowBijector_{2,3,4}blocks
  flow_models/flow_builder.py:60-146         build_glow (prior, Invert(Chain))
  run_basis_sep.py:73-79                     compute_grad_logprob
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # Keras BatchNormalization default epsilon


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a), dtype=dtype)


# --------------------------------------------------------------------------- primitives
def squeeze(x: torch.Tensor) -> torch.Tensor:
    """flow_tfp_bijectors.py:170-174: out channel = c*4 + dh*2 + dw."""
    N, H, W, C = x.shape
    x = x.reshape(N, H // 2, 2, W // 2, 2, C)
    x = x.permute(0, 1, 3, 5, 2, 4)
    return x.reshape(N, H // 2, W // 2, C * 4)


def unsqueeze(y: torch.Tensor) -> torch.Tensor:
    """flow_tfp_bijectors.py:176-180."""
    N, H, W, C4 = y.shape
    C = C4 // 4
    y = y.reshape(N, H, W, C, 2, 2)
    y = y.permute(0, 1, 4, 2, 5, 3)
    return y.reshape(N, H * 2, W * 2, C)


def conv2d_same(x: torch.Tensor, k_hwio: torch.Tensor, bias: Optional[torch.Tensor], dilation: int = 1) -> torch.Tensor:
    """TF ``padding='same'`` stride-1 cross-correlation on NHWC with HWIO filters."""
    kh = k_hwio.shape[0]
    pad = dilation * (kh // 2)
    y = F.conv2d(x.permute(0, 3, 1, 2).contiguous(), k_hwio.permute(3, 2, 0, 1).contiguous(), bias, padding=pad, dilation=dilation)
    return y.permute(0, 2, 3, 1)


def inv1x1_weight(P, L, U, log_S, sign_S) -> torch.Tensor:
    """flow_tfp_bijectors.py:300-303: W = P (L*tril(-1)+I) (U*triu(+1)+diag(sign*exp(log_s)))."""
    C = P.shape[0]
    l_mask = torch.tril(torch.ones(C, C, dtype=P.dtype), -1)
    Lm = L * l_mask + torch.eye(C, dtype=P.dtype)
    Um = U * l_mask.t() + torch.diag(sign_S * torch.exp(log_S))
    return P @ (Lm @ Um)


def inv1x1_weight_inverse(P, L, U, log_S, sign_S) -> torch.Tensor:
    """flow_tfp_bijectors.py:309-315: W^-1 = U^-1 L^-1 P^-1 (P^-1 is the stored P_inv)."""
    C = P.shape[0]
    l_mask = torch.tril(torch.ones(C, C, dtype=P.dtype), -1)
    Lm = L * l_mask + torch.eye(C, dtype=P.dtype)
    Um = U * l_mask.t() + torch.diag(sign_S * torch.exp(log_S))
    return torch.linalg.inv(Um) @ (torch.linalg.inv(Lm) @ torch.linalg.inv(P))


# --------------------------------------------------------------------------- the model
class GlowOracle:
    """log_prob / forward / inverse / fldj / grad_log_prob of one Glow prior.

    ``coupling_nn`` optionally replaces ShiftAndLogScaleConvNet by a callable
    ``f(xb) -> (log_s, t)`` -- the seam the reference's tests use to inject the toy
    constant network (unittest_flow_models.py:76-83).
    """

    def __init__(self, cfg, params: Dict[str, np.ndarray], dtype=torch.float64,
                 coupling_nn: Optional[Callable] = None):
        self.cfg = cfg
        self.dtype = dtype
        self.p = {k: _t(v, dtype) for k, v in params.items()}
        self.coupling_nn = coupling_nn

    # ---- coupling network: flow_tfk_layers.py:73-84, BatchNorm in inference mode
    def nn(self, xb: torch.Tensor, pre: str) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.coupling_nn is not None:
            return self.coupling_nn(xb)
        p = self.p

        def bn(x, name):
            g = p[pre + f"nn/{name}/gamma"] / torch.sqrt(p[pre + f"nn/{name}/moving_variance"] + BN_EPS)
            return (x - p[pre + f"nn/{name}/moving_mean"]) * g + p[pre + f"nn/{name}/beta"]

        h = torch.relu(conv2d_same(xb, p[pre + "nn/conv1/kernel"], p[pre + "nn/conv1/bias"]))
        h = bn(h, "bn1")
        h = torch.relu(h @ p[pre + "nn/conv2/kernel"] + p[pre + "nn/conv2/bias"])
        h = bn(h, "bn2")
        r = conv2d_same(h, p[pre + "nn/conv3/kernel"], p[pre + "nn/conv3/bias"])
        C = r.shape[-1]
        return torch.tanh(r[..., : C // 2]), r[..., C // 2:]

    # ---- one GlowStep: flow_glow.py:21-22 (actnorm -> 1x1 -> coupling)
    def step_forward(self, x: torch.Tensor, pre: str) -> Tuple[torch.Tensor, torch.Tensor]:
        p = self.p
        N, H, W, C = x.shape
        a = x * torch.exp(p[pre + "actnorm/log_scale"]) + p[pre + "actnorm/shift"]       # :243
        Wm = inv1x1_weight(*(p[pre + "inv1x1/" + n] for n in ("P", "L", "U", "log_S", "sign_S")))
        u = a @ Wm                                                                         # :304-305
        ua, ub = u[..., : C // 2], u[..., C // 2:]                                         # :135
        log_s, t = self.nn(ub, pre)
        ya = torch.exp(log_s) * ua + t                                                     # :137-138
        y = torch.cat([ya, ub], dim=-1)
        logdet = (H * W * (p[pre + "actnorm/log_scale"].sum() + p[pre + "inv1x1/log_S"].sum())
                  + log_s.reshape(N, -1).sum(dim=1))                                       # :250-253,:319-322,:150-153
        return y, logdet

    def step_inverse(self, y: torch.Tensor, pre: str) -> torch.Tensor:
        p = self.p
        C = y.shape[-1]
        ya, yb = y[..., : C // 2], y[..., C // 2:]
        log_s, t = self.nn(yb, pre)
        ua = (ya - t) / torch.exp(log_s)                                                   # :146
        u = torch.cat([ua, yb], dim=-1)
        Wi = inv1x1_weight_inverse(*(p[pre + "inv1x1/" + n] for n in ("P", "L", "U", "log_S", "sign_S")))
        a = u @ Wi                                                                         # :316
        return (a - p[pre + "actnorm/shift"]) / torch.exp(p[pre + "actnorm/log_scale"])   # :246-247

    # ---- GlowBlock: flow_glow.py:51-52 -- squeeze, then steps K-1 ... 0
    def block_forward(self, x, b):
        x = squeeze(x)
        logdet = torch.zeros(x.shape[0], dtype=self.dtype)
        for k in reversed(range(self.cfg.K)):
            x, ld = self.step_forward(x, f"b{b}/s{k}/")
            logdet = logdet + ld
        return x, logdet

    def block_inverse(self, y, b):
        for k in range(self.cfg.K):
            y = self.step_inverse(y, f"b{b}/s{k}/")
        return unsqueeze(y)

    # ---- GlowBijector_{2,3,4}blocks: flow_glow.py:102-108,176-185,273-286
    def bijector_forward(self, x):
        cfg = self.cfg
        N = x.shape[0]
        Hl, Wl, _ = cfg.latent_shape
        zs = []
        logdet = torch.zeros(N, dtype=self.dtype)
        h = x
        for b in range(cfg.L):
            o, ld = self.block_forward(h, b)
            logdet = logdet + ld
            if b < cfg.L - 1:
                C = o.shape[-1]
                z, h = o[..., : C // 2], o[..., C // 2:]
                zs.append(z.reshape(N, Hl, Wl, -1))          # plain row-major reshape (Q2)
            else:
                zs.append(o)
        return torch.cat(zs, dim=-1), logdet

    def bijector_inverse(self, z):
        cfg = self.cfg
        N = z.shape[0]
        # concat(z1, concat(z2, ... zL)): z_i has half of what remains
        parts = []
        rest = z
        for b in range(cfg.L - 1):
            C = rest.shape[-1]
            parts.append(rest[..., : C // 2])
            rest = rest[..., C // 2:]
        parts.append(rest)
        h = self.block_inverse(parts[-1], cfg.L - 1)
        for b in reversed(range(cfg.L - 1)):
            Hb, Wb, Cb = cfg.level_shape(b)
            zb = parts[b].reshape(N, Hb, Wb, Cb // 2)
            h = self.block_inverse(torch.cat([zb, h], dim=-1), b)
        return h

    # ---- SpecPreprocessing (use_logit=False): flow_tfp_bijectors.py:372-396
    def pre_forward(self, x):
        return (x - self.cfg.minval) / (self.cfg.maxval - self.cfg.minval) - 0.5

    def pre_inverse(self, y):
        return (y + 0.5) * (self.cfg.maxval - self.cfg.minval) + self.cfg.minval

    def pre_fldj(self):
        return self.cfg.dims * math.log(1.0 / (self.cfg.maxval - self.cfg.minval))

    # ---- distribution surface: flow_builder.py:127-144
    def forward(self, x):
        """Chain([glow, preprocessing]).forward -> (z, fldj)."""
        x = _t(x, self.dtype) if not torch.is_tensor(x) else x
        z, ld = self.bijector_forward(self.pre_forward(x))
        return z, ld + self.pre_fldj()

    def inverse(self, z):
        z = _t(z, self.dtype) if not torch.is_tensor(z) else z
        return self.pre_inverse(self.bijector_inverse(z))

    def prior_log_prob(self, z):
        N = z.shape[0]
        if self.cfg.learntop:
            loc, ls = self.p["prior/loc"], self.p["prior/log_scale"]
        else:
            loc = torch.zeros(self.cfg.latent_shape, dtype=self.dtype)
            ls = torch.zeros(self.cfg.latent_shape, dtype=self.dtype)
        q = (z - loc) / torch.exp(ls)
        lp = -0.5 * q * q - ls - 0.5 * math.log(2.0 * math.pi)
        return lp.reshape(N, -1).sum(dim=1)

    def log_prob(self, x):
        z, fldj = self.forward(x)
        return self.prior_log_prob(z) + fldj

    def grad_log_prob(self, x):
        """run_basis_sep.py:73-79 (autograd instead of GradientTape)."""
        x = _t(x, self.dtype).clone().requires_grad_(True)
        lp = self.log_prob(x)
        (g,) = torch.autograd.grad(lp.sum(), x)
        return g.detach(), lp.detach()

    def sample_from_latent(self, eps):
        """sample(n) with the standard-normal draw ``eps`` injected: z = loc + exp(ls)*eps."""
        eps = _t(eps, self.dtype)
        if self.cfg.learntop:
            z = self.p["prior/loc"] + torch.exp(self.p["prior/log_scale"]) * eps
        else:
            z = eps
        return self.inverse(z)

    # ---- the reference's graph re-evaluates the NN in _forward_log_det_jacobian and re-runs
    # block forwards (flow_glow.py:198-209); used only to time the CPU baseline faithfully.
    def log_prob_reference_graph(self, x):
        x = _t(x, self.dtype) if not torch.is_tensor(x) else x
        z, _ = self.forward(x)              # Chain.forward
        _, fldj = self.forward(x)           # Chain.forward_log_det_jacobian re-walks the chain
        return self.prior_log_prob(z) + fldj

    # ---- ActNorm data-dependent init incl. the raw-minibatch quirk (Q7)
    def init_actnorm(self, minibatch_raw):
        """flow_builder.py:96 + flow_glow.py:44-49,156-174 + flow_tfp_bijectors.py:222-240.

        Walks the minibatch through steps in CONSTRUCTION order 0..K-1; for L >= 3 blocks
        2.. receive the raw (preprocessed) minibatch re-tiled by Squeeze's -1 reshape.
        Returns the dict of new actnorm parameters (float32 numpy) and updates self.p.
        """
        cfg = self.cfg
        mb0 = self.pre_forward(_t(minibatch_raw, self.dtype))
        new = {}

        def squeeze_any(x, H, W, C):
            x = x.reshape(-1, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 5, 2, 4)
            return x.reshape(-1, H // 2, W // 2, C * 4)

        carried = mb0
        for b in range(cfg.L):
            Hb, Wb, Cb = cfg.level_shape(b)
            Hin, Win, Cin = Hb * 2, Wb * 2, Cb // 4
            quirk = cfg.L >= 3 and b >= 1
            src = mb0 if quirk else carried
            mb = squeeze_any(src, Hin, Win, Cin)
            for k in range(cfg.K):
                pre = f"b{b}/s{k}/"
                mean = mb.mean(dim=(0, 1, 2))
                std = mb.std(dim=(0, 1, 2), unbiased=False) + 1e-8
                self.p[pre + "actnorm/log_scale"] = torch.log(1.0 / std)
                self.p[pre + "actnorm/shift"] = -mean / std
                new[pre + "actnorm/log_scale"] = self.p[pre + "actnorm/log_scale"].numpy().astype(np.float32)
                new[pre + "actnorm/shift"] = self.p[pre + "actnorm/shift"].numpy().astype(np.float32)
                mb, _ = self.step_forward(mb, pre)
            if b < cfg.L - 1:
                # the constructors re-run block.forward (steps K-1..0) on the carried batch
                o, _ = self.block_forward(carried, b)
                carried = o[..., o.shape[-1] // 2:]
        return new
