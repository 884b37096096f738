#!/usr/bin/env python
"""Noise-conditioned Glow fine-tuning CLI; see audiosourcesep_b200/train_noisy_glow.py (reference: train_noisy_glow.py)."""
from audiosourcesep_b200.train_noisy_glow import build_parser, main

if __name__ == "__main__":
    main(build_parser().parse_args())
