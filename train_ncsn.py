#!/usr/bin/env python
"""NCSN training CLI; see audiosourcesep_b200/train_ncsn.py (reference: train_ncsn.py)."""
from audiosourcesep_b200.train_ncsn import build_parser, main

if __name__ == "__main__":
    main(build_parser().parse_args())
