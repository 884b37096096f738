#!/usr/bin/env python
"""Glow / noisy-Glow training CLI; see audiosourcesep_b200/train_glow.py (reference: train_glow.py, train_noisy_glow.py)."""
from audiosourcesep_b200.train_glow import build_parser, main

if __name__ == "__main__":
    main(build_parser().parse_args())
