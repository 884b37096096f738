"""Host-side mirror of the reference's ``train_noisy_glow.py``: one Glow fine-tuned per noise level, serially, each level
warm-started from the previous one, the loss evaluated on ``X + sigma * N(0, 1)`` in raw data units
(reference: train_noisy_glow.py:30-33, :211, :309-358); the weights of level sigma go to
``<output>/sigma_<round(sigma, 2)>/weights.npz``, the layout run_basis_sep reads (run_basis_sep.py:284-285).

The loop itself lives in :mod:`audiosourcesep_b200.train_glow` (``--noisy``); this module is the reference's entry point.
"""
from .train_glow import build_parser as _build_parser
from .train_glow import distributed_train_step, main as _main, setUp_optimizer, train  # noqa: F401


def build_parser():
    p = _build_parser()
    p.set_defaults(noisy=True)
    return p


def main(args):
    args.noisy = True
    return _main(args)


if __name__ == "__main__":
    main(build_parser().parse_args())
