#!/usr/bin/env python
"""Unconditional annealed-Langevin sampling CLI; see audiosourcesep_b200/ncsn_generate_samples.py
(reference: ncsn_generate_samples.py)."""
from audiosourcesep_b200.ncsn_generate_samples import build_parser, main

if __name__ == "__main__":
    main(build_parser().parse_args())
