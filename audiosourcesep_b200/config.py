"""Model configurations of the separation hot path.

Mirrors the hyper-parameters the reference reads from ``configs/melspec_*.yml``
(reference: configs/melspec_glow.yml:1-18, configs/melspec_noisy_glow.yml:1-24,
configs/melspec_ncsnv1.yml, configs/melspec_ncsnv2.yml) and the derived shapes of
``flow_builder.build_glow`` (reference: flow_models/flow_builder.py:60-76).
"""
from __future__ import annotations

import argparse
import dataclasses
from typing import List, Tuple

import yaml


@dataclasses.dataclass(frozen=True)
class GlowConfig:
    """Shape contract of one Glow prior (reference: flow_models/flow_glow.py:80-330)."""

    H: int = 96
    W: int = 64
    C: int = 1
    L: int = 3
    K: int = 40
    n_filters: int = 512
    learntop: bool = True
    minval: float = -100.0
    maxval: float = 20.0

    def __post_init__(self):
        if self.L not in (2, 3, 4):
            raise ValueError("L should be 2, 3 or 4")  # flow_builder.py:77-78
        if self.H % (1 << self.L) or self.W % (1 << self.L):
            raise ValueError("H and W must be divisible by 2**L (Squeeze asserts, flow_tfp_bijectors.py:165-166)")

    def level_shape(self, b: int) -> Tuple[int, int, int]:
        """(H, W, C) of the state inside block ``b`` (after its squeeze)."""
        c = self.C
        for _ in range(b):
            c = (c * 4) // 2
        return self.H >> (b + 1), self.W >> (b + 1), c * 4

    def level_shapes(self) -> List[Tuple[int, int, int]]:
        return [self.level_shape(b) for b in range(self.L)]

    @property
    def latent_shape(self) -> Tuple[int, int, int]:
        """``base_distr_shape`` (reference: flow_builder.py:65-76)."""
        s = 1 << self.L
        return self.H // s, self.W // s, self.C * s * s

    @property
    def dims(self) -> int:
        return self.H * self.W * self.C


@dataclasses.dataclass(frozen=True)
class NCSNConfig:
    """Score-network shape contract (reference: ncsn/utils.py:41-64)."""

    version: str = "v1"
    H: int = 96
    W: int = 64
    C: int = 1
    ngf: int = 192
    num_classes: int = 10
    sigma1: float = 1.0
    sigmaL: float = 0.01
    progression: str = "logarithmic"


def dict2namespace(config: dict) -> argparse.Namespace:
    """reference: train_utils.py:123-131."""
    ns = argparse.Namespace()
    for key, value in config.items():
        setattr(ns, key, dict2namespace(value) if isinstance(value, dict) else value)
    return ns


def get_config(config_path: str) -> argparse.Namespace:
    """YAML -> Namespace (reference: train_utils.py:114-120; uses safe_load because
    PyYAML >= 6 rejects the reference's loader-less ``yaml.load``)."""
    with open(config_path, "r") as f:
        return dict2namespace(yaml.safe_load(f))
