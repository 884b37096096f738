#!/usr/bin/env python
"""BASIS separation CLI (same flags as the reference's run_basis_sep.py:453-523); see
audiosourcesep_b200/run_basis_sep.py."""
from audiosourcesep_b200.run_basis_sep import cli

if __name__ == "__main__":
    cli()
