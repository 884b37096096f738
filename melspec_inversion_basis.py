#!/usr/bin/env python
"""Spectrogram inversion CLI; see audiosourcesep_b200/melspec_inversion_basis.py (reference: melspec_inversion_basis.py)."""
from audiosourcesep_b200.melspec_inversion_basis import build_parser, main

if __name__ == "__main__":
    main(build_parser().parse_args())
