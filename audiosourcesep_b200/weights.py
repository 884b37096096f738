"""Seeded parameter generators for the Glow prior.

The reference ships no checkpoints (SURVEY.md ground facts), so every parity and
throughput run uses random-init weights.  Two generators:

* ``faithful``  -- what the reference constructors produce: QR->LU 1x1 convolution
  (reference: flow_models/flow_tfp_bijectors.py:271-294), Glorot-uniform conv1/conv2
  with zero bias and a ZERO conv3 (reference: flow_models/flow_tfk_layers.py:56-70),
  BatchNorm gamma=1/beta=0/moving stats 0/1, prior loc=0/log-scale=0
  (reference: flow_models/flow_builder.py:132-139).  ActNorm is left at identity
  here; its data-dependent init is a model operation (``Glow.init_actnorm``).
* ``perturbed`` -- same structure but every tensor non-trivial (non-zero conv3,
  jittered BN statistics, prior, ActNorm) so that parity tests are not vacuous
  ("two traps").

Parameter names (flat dict, float32 numpy):
  b{b}/s{k}/actnorm/{log_scale,shift}                       [C]
  b{b}/s{k}/inv1x1/{P,L,U} [C,C]   b{b}/s{k}/inv1x1/{log_S,sign_S} [C]
  b{b}/s{k}/nn/conv1/{kernel [3,3,C/2,F], bias [F]}
  b{b}/s{k}/nn/bn1/{gamma,beta,moving_mean,moving_variance} [F]
  b{b}/s{k}/nn/conv2/{kernel [F,F] (in,out), bias [F]}
  b{b}/s{k}/nn/bn2/{...}
  b{b}/s{k}/nn/conv3/{kernel [3,3,F,C], bias [C]}
  prior/{loc,log_scale}                                     [H/2^L, W/2^L, C*4^L]
``k`` is the CONSTRUCTION index of the step (``glowStep_k``); the forward pass runs
k = K-1 ... 0 (reference: flow_models/flow_glow.py:51-52).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import scipy.linalg

from .config import GlowConfig

STEP_PARAM_SUFFIXES = (
    "actnorm/log_scale", "actnorm/shift",
    "inv1x1/P", "inv1x1/L", "inv1x1/U", "inv1x1/log_S", "inv1x1/sign_S",
    "nn/conv1/kernel", "nn/conv1/bias",
    "nn/bn1/gamma", "nn/bn1/beta", "nn/bn1/moving_mean", "nn/bn1/moving_variance",
    "nn/conv2/kernel", "nn/conv2/bias",
    "nn/bn2/gamma", "nn/bn2/beta", "nn/bn2/moving_mean", "nn/bn2/moving_variance",
    "nn/conv3/kernel", "nn/conv3/bias",
)
# Variables the reference marks trainable=False (flow_tfp_bijectors.py:281-287) plus
# the BatchNorm moving statistics (never updated).
FROZEN_SUFFIXES = ("inv1x1/P", "inv1x1/sign_S",
                   "nn/bn1/moving_mean", "nn/bn1/moving_variance",
                   "nn/bn2/moving_mean", "nn/bn2/moving_variance")


def step_param_shapes(C: int, F: int) -> Dict[str, Tuple[int, ...]]:
    return {
        "actnorm/log_scale": (C,), "actnorm/shift": (C,),
        "inv1x1/P": (C, C), "inv1x1/L": (C, C), "inv1x1/U": (C, C),
        "inv1x1/log_S": (C,), "inv1x1/sign_S": (C,),
        "nn/conv1/kernel": (3, 3, C // 2, F), "nn/conv1/bias": (F,),
        "nn/bn1/gamma": (F,), "nn/bn1/beta": (F,),
        "nn/bn1/moving_mean": (F,), "nn/bn1/moving_variance": (F,),
        "nn/conv2/kernel": (F, F), "nn/conv2/bias": (F,),
        "nn/bn2/gamma": (F,), "nn/bn2/beta": (F,),
        "nn/bn2/moving_mean": (F,), "nn/bn2/moving_variance": (F,),
        "nn/conv3/kernel": (3, 3, F, C), "nn/conv3/bias": (C,),
    }


def glow_param_shapes(cfg: GlowConfig) -> Dict[str, Tuple[int, ...]]:
    shapes: Dict[str, Tuple[int, ...]] = {}
    for b in range(cfg.L):
        _, _, C = cfg.level_shape(b)
        for k in range(cfg.K):
            for suf, shp in step_param_shapes(C, cfg.n_filters).items():
                shapes[f"b{b}/s{k}/{suf}"] = shp
    if cfg.learntop:
        shapes["prior/loc"] = cfg.latent_shape
        shapes["prior/log_scale"] = cfg.latent_shape
    return shapes


def glow_param_names(cfg: GlowConfig) -> List[str]:
    return list(glow_param_shapes(cfg).keys())


def is_trainable(name: str) -> bool:
    return not any(name.endswith(s) for s in FROZEN_SUFFIXES)


def count_trainable(cfg: GlowConfig) -> int:
    """Known answer for the melspec config: 39,598,720 + 2*6,144."""
    return int(sum(int(np.prod(s)) for n, s in glow_param_shapes(cfg).items() if is_trainable(n)))


def _glorot_uniform(rng, shape, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape)


def _inv1x1_init(rng, C):
    """QR of randn -> LU (reference: flow_tfp_bijectors.py:271-278)."""
    w = np.linalg.qr(rng.standard_normal((C, C)))[0]
    p, l, u = scipy.linalg.lu(w)
    s = np.diag(u)
    return p, l, np.triu(u, k=1), np.log(np.abs(s)), np.sign(s)


def init_glow_params(cfg: GlowConfig, seed: int = 2, mode: str = "perturbed", coupling_gain: float = None,
                     actnorm_jitter: float = None) -> Dict[str, np.ndarray]:
    """Generate a full parameter set.  ``mode`` in {'faithful', 'perturbed'}.

    ``coupling_gain`` scales conv3 (strength of the couplings) and ``actnorm_jitter`` the ActNorm
    parameters of the perturbed mode.  A ReLU coupling network is positively homogeneous, so strong
    un-normalised couplings make a 120-step flow grow geometrically; the defaults are therefore
    0.5 / 0.1 for shallow test models (K <= 8) and 0.1 / 0.03 for deep ones (checked with the
    oracle: the melspec model stays O(1), log_prob ~ -3.5e4 nats).
    """
    if mode not in ("faithful", "perturbed"):
        raise ValueError("mode must be 'faithful' or 'perturbed'")
    rng = np.random.Generator(np.random.PCG64(seed))
    F = cfg.n_filters
    pert = mode == "perturbed"
    if coupling_gain is None:
        coupling_gain = 0.5 if cfg.K <= 8 else 0.1
    if actnorm_jitter is None:
        actnorm_jitter = 0.1 if cfg.K <= 8 else 0.03
    out: Dict[str, np.ndarray] = {}
    for b in range(cfg.L):
        _, _, C = cfg.level_shape(b)
        Ch = C // 2
        for k in range(cfg.K):
            pre = f"b{b}/s{k}/"
            p, l, u, log_s, sign_s = _inv1x1_init(rng, C)
            if pert:
                l = l + np.tril(rng.normal(0, 0.02, (C, C)), -1)
                u = u + np.triu(rng.normal(0, 0.02, (C, C)), 1)
                log_s = log_s + rng.normal(0, 0.05, C)
                out[pre + "actnorm/log_scale"] = rng.normal(0, actnorm_jitter, C)
                out[pre + "actnorm/shift"] = rng.normal(0, actnorm_jitter, C)
                out[pre + "nn/conv1/kernel"] = rng.normal(0, np.sqrt(2.0 / (9 * Ch)), (3, 3, Ch, F))
                out[pre + "nn/conv1/bias"] = rng.normal(0, 0.1, F)
                out[pre + "nn/conv2/kernel"] = rng.normal(0, np.sqrt(2.0 / F), (F, F))
                out[pre + "nn/conv2/bias"] = rng.normal(0, 0.1, F)
                out[pre + "nn/conv3/kernel"] = rng.normal(0, coupling_gain / np.sqrt(9 * F), (3, 3, F, C))
                out[pre + "nn/conv3/bias"] = rng.normal(0, 0.1 * coupling_gain, C)
                for bn in ("bn1", "bn2"):
                    out[pre + f"nn/{bn}/gamma"] = rng.normal(1.0, 0.1, F)
                    out[pre + f"nn/{bn}/beta"] = rng.normal(0.0, 0.1, F)
                    out[pre + f"nn/{bn}/moving_mean"] = rng.normal(0.0, 0.1, F)
                    out[pre + f"nn/{bn}/moving_variance"] = rng.uniform(0.5, 1.5, F)
            else:
                out[pre + "actnorm/log_scale"] = np.zeros(C)
                out[pre + "actnorm/shift"] = np.zeros(C)
                out[pre + "nn/conv1/kernel"] = _glorot_uniform(rng, (3, 3, Ch, F), 9 * Ch, 9 * F)
                out[pre + "nn/conv1/bias"] = np.zeros(F)
                out[pre + "nn/conv2/kernel"] = _glorot_uniform(rng, (F, F), F, F)
                out[pre + "nn/conv2/bias"] = np.zeros(F)
                out[pre + "nn/conv3/kernel"] = np.zeros((3, 3, F, C))
                out[pre + "nn/conv3/bias"] = np.zeros(C)
                for bn in ("bn1", "bn2"):
                    out[pre + f"nn/{bn}/gamma"] = np.ones(F)
                    out[pre + f"nn/{bn}/beta"] = np.zeros(F)
                    out[pre + f"nn/{bn}/moving_mean"] = np.zeros(F)
                    out[pre + f"nn/{bn}/moving_variance"] = np.ones(F)
            out[pre + "inv1x1/P"] = p
            out[pre + "inv1x1/L"] = l
            out[pre + "inv1x1/U"] = u
            out[pre + "inv1x1/log_S"] = log_s
            out[pre + "inv1x1/sign_S"] = sign_s
    if cfg.learntop:
        if pert:
            out["prior/loc"] = rng.normal(0, 0.1, cfg.latent_shape)
            out["prior/log_scale"] = rng.normal(0, 0.1, cfg.latent_shape)
        else:
            out["prior/loc"] = np.zeros(cfg.latent_shape)
            out["prior/log_scale"] = np.zeros(cfg.latent_shape)
    shapes = glow_param_shapes(cfg)
    res = {}
    for name, shp in shapes.items():
        a = np.ascontiguousarray(out[name], dtype=np.float32)
        assert a.shape == tuple(shp), (name, a.shape, shp)
        res[name] = a
    return res


# =============================================================================== NCSN
# Parameter inventory of the score networks, walked from the reference constructors
# (ncsn/score_network.py:224-272, ncsn/score_network_v2.py:202-251).  Known answers:
# v1 ngf=192 / 10 classes -> 67,464,769 trainables (reference log
# trained_ncsn/ncsn_piano_192_32_dB_custom_loop/out.log:35); v2 ngf=128 -> 29,695,233.
def _norm_shapes(prefix, C, version, num_classes):
    if version == "v1":
        # Embedding(num_classes, 3C) rows = [gamma | alpha | beta]  (score_network.py:190-195)
        return {prefix + "/embed": (num_classes, 3 * C),
                prefix + "/in_gamma": (C,), prefix + "/in_beta": (C,)}
    return {prefix + "/alpha": (C,), prefix + "/gamma": (C,), prefix + "/beta": (C,),
            prefix + "/in_gamma": (C,), prefix + "/in_beta": (C,)}


def _conv_shapes(prefix, k, cin, cout, bias):
    d = {prefix + "/kernel": (k, k, cin, cout)}
    if bias:
        d[prefix + "/bias"] = (cout,)
    return d


def ncsn_res_block_spec(name, cin, cout, resample, dilation):
    """Static description of one (Conditional)ResidualBlock
    (score_network.py:121-178 / score_network_v2.py:110-171)."""
    down = resample == "down"
    dil = dilation or 1
    if down and dilation is None:
        spec = dict(conv1=(3, cin, cin, False, 1), norm2=cin, conv2=(3, cin, cout, True, 1), pool=True,
                    shortcut=(1, cin, cout, True, 1))
    elif down:
        spec = dict(conv1=(3, cin, cin, True, dil), norm2=cin, conv2=(3, cin, cout, True, dil), pool=False,
                    shortcut=(3, cin, cout, True, dil))
    elif dilation is not None:
        spec = dict(conv1=(3, cin, cout, True, dil), norm2=cout, conv2=(3, cout, cout, True, dil), pool=False,
                    shortcut=None if cin == cout else (3, cin, cin, True, dil))
    else:
        spec = dict(conv1=(3, cin, cout, False, 1), norm2=cout, conv2=(3, cout, cout, False, 1), pool=False,
                    shortcut=None if cin == cout else (3, cin, cout, False, 1))
    spec.update(name=name, cin=cin, cout=cout)
    return spec


def ncsn_layout(ngf):
    """Residual stages and refine blocks in call order (score_network.py:238-272)."""
    res = [
        ncsn_res_block_spec("Res1_1", ngf, ngf, None, None), ncsn_res_block_spec("Res1_2", ngf, ngf, None, None),
        ncsn_res_block_spec("Res2_1", ngf, 2 * ngf, "down", None), ncsn_res_block_spec("Res2_2", 2 * ngf, 2 * ngf, None, None),
        ncsn_res_block_spec("Res3_1", 2 * ngf, 2 * ngf, "down", 2), ncsn_res_block_spec("Res3_2", 2 * ngf, 2 * ngf, None, 2),
        ncsn_res_block_spec("Res4_1", 2 * ngf, 2 * ngf, "down", 4), ncsn_res_block_spec("Res4_2", 2 * ngf, 2 * ngf, None, 4),
    ]
    refine = [
        dict(name="refine1", in_planes=[2 * ngf], features=2 * ngf, start=True, end=False),
        dict(name="refine2", in_planes=[2 * ngf, 2 * ngf], features=2 * ngf, start=False, end=False),
        dict(name="refine3", in_planes=[2 * ngf, 2 * ngf], features=ngf, start=False, end=False),
        dict(name="refine4", in_planes=[ngf, ngf], features=ngf, start=False, end=True),
    ]
    return res, refine


def ncsn_param_shapes(cfg) -> Dict[str, Tuple[int, ...]]:
    v, ngf, nc = cfg.version, cfg.ngf, cfg.num_classes
    norm_in_rcu_msf_crp = v == "v1"      # v2 RCU/MSF/CRP have no norm layers
    sh: Dict[str, Tuple[int, ...]] = {}
    sh.update(_conv_shapes("begin_conv", 3, cfg.C, ngf, True))
    res, refine = ncsn_layout(ngf)
    for s in res:
        n = s["name"]
        sh.update(_norm_shapes(n + "/norm1", s["cin"], v, nc))
        k, ci, co, b, _ = s["conv1"]
        sh.update(_conv_shapes(n + "/conv1", k, ci, co, b))
        sh.update(_norm_shapes(n + "/norm2", s["norm2"], v, nc))
        k, ci, co, b, _ = s["conv2"]
        sh.update(_conv_shapes(n + "/conv2", k, ci, co, b))
        if s["shortcut"] is not None:
            k, ci, co, b, _ = s["shortcut"]
            sh.update(_conv_shapes(n + "/shortcut", k, ci, co, b))

    def rcu(prefix, feat, n_blocks, n_stages):
        for i in range(n_blocks):
            for j in range(n_stages):
                if norm_in_rcu_msf_crp:
                    sh.update(_norm_shapes(f"{prefix}/norm_{i + 1}_{j + 1}", feat, v, nc))
                sh.update(_conv_shapes(f"{prefix}/conv_{i + 1}_{j + 1}", 3, feat, feat, False))

    for r in refine:
        n = r["name"]
        for i, cin in enumerate(r["in_planes"]):
            rcu(f"{n}/RCU_{i + 1}", cin, 2, 2)
        if not r["start"]:
            for i, cin in enumerate(r["in_planes"]):
                if norm_in_rcu_msf_crp:
                    sh.update(_norm_shapes(f"{n}/MSF/norm_{i + 1}", cin, v, nc))
                sh.update(_conv_shapes(f"{n}/MSF/conv_{i + 1}", 3, cin, r["features"], True))
        for i in range(2):
            if norm_in_rcu_msf_crp:
                sh.update(_norm_shapes(f"{n}/CRP/norm_{i + 1}", r["features"], v, nc))
            sh.update(_conv_shapes(f"{n}/CRP/conv_{i + 1}", 3, r["features"], r["features"], False))
        rcu(f"{n}/RCU_output", r["features"], 3 if r["end"] else 1, 2)
    sh.update(_norm_shapes("normalizer", ngf, v, nc))
    sh.update(_conv_shapes("end_conv", 3, ngf, cfg.C, True))
    return sh


def count_ncsn_params(cfg) -> int:
    return int(sum(int(np.prod(s)) for s in ncsn_param_shapes(cfg).values()))


def init_ncsn_params(cfg, seed: int = 2, mode: str = "perturbed") -> Dict[str, np.ndarray]:
    """``faithful``: Glorot-uniform kernels, zero biases, norm gains N(0, 0.02) (quirk Q9,
    score_network.py:187-188), instance-norm gamma=1/beta=0.  ``perturbed``: gains N(1, 0.1)
    etc. so that no branch is numerically muted."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pert = mode == "perturbed"
    out = {}
    for name, shp in ncsn_param_shapes(cfg).items():
        leaf = name.rsplit("/", 1)[1]
        if leaf == "kernel":
            k, _, ci, co = shp
            if pert:
                a = rng.normal(0, np.sqrt(1.0 / (k * k * ci)), shp)
            else:
                a = _glorot_uniform(rng, shp, k * k * ci, k * k * co)
        elif leaf == "bias":
            a = rng.normal(0, 0.05, shp) if pert else np.zeros(shp)
        elif leaf == "embed":
            C = shp[1] // 3
            if pert:
                a = np.concatenate([rng.normal(1, 0.1, (shp[0], C)), rng.normal(0, 0.1, (shp[0], C)),
                                    rng.normal(0, 0.1, (shp[0], C))], axis=-1)
            else:
                a = np.concatenate([rng.normal(0, 0.02, (shp[0], C)), rng.normal(0, 0.02, (shp[0], C)),
                                    np.zeros((shp[0], C))], axis=-1)
        elif leaf == "gamma":
            a = rng.normal(1, 0.1, shp) if pert else rng.normal(0, 0.02, shp)
        elif leaf == "alpha":
            a = rng.normal(0, 0.1, shp) if pert else rng.normal(0, 0.02, shp)
        elif leaf == "beta":
            a = rng.normal(0, 0.1, shp) if pert else np.zeros(shp)
        elif leaf == "in_gamma":
            a = rng.normal(1, 0.1, shp) if pert else np.ones(shp)
        elif leaf == "in_beta":
            a = rng.normal(0, 0.1, shp) if pert else np.zeros(shp)
        else:
            raise KeyError(name)
        out[name] = np.ascontiguousarray(a, dtype=np.float32)
    return out
