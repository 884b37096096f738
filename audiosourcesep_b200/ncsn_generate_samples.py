"""Unconditional annealed-Langevin sampling CLI -- host-side mirror of the reference's ``ncsn_generate_samples.py``.

Same flags (reference: ncsn_generate_samples.py:119-173), same config precedence (:26-32), same print-outs
("SAMPLING PARAMETERS", "Weights loaded", "Start Generating ...", "Done. Duration: ...", "Shape: ...",
"Generated Samples saved at ..."), same output: ``<filename>.npy`` holding the post-processed samples after every
noise level, shape ``[L + 1, n_samples, H, W, 1]`` (:98-116).  The sampler is ``ncsn.utils.anneal_langevin_dynamics``
(reference ncsn/utils.py:17-38): the score network and the fused Langevin update run in libasep.so.

Weights: ``RESTORE`` is a directory holding ``weights.npz`` (the repo's checkpoint container) or, with
``--random_init SEED``, the seeded synthetic weights (there is no trained checkpoint in the reference tree).
Additions: ``--random_init``, ``--seed`` (Philox seed of the Langevin noise and of the uniform start), ``--fast``.
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np

from .config import get_config
from .ncsn.utils import anneal_langevin_dynamics, get_sigmas, get_uncompiled_model, get_uncompiled_model_v2


def setUp_optimizer(args):
    """reference: ncsn_generate_samples.py:12-21 -- only validates the name (no optimizer state is restored here)."""
    if args.optimizer not in ("adam", "adamax"):
        raise ValueError("optimizer argument should be adam or adamax")
    return args.optimizer


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Sample from NCSN model")
    p.add_argument("RESTORE", type=str, default=None, help="directory of saved weights")
    p.add_argument("--filename", type=str, default=None, help="filename for savings")
    p.add_argument("--dataset", type=str, default="melspec")
    p.add_argument("--n_samples", type=int, default=32)
    p.add_argument("--config", type=str, help="path to the config file. Overwrite all other parameters below")
    p.add_argument("--version", type=str, default="v2")
    p.add_argument("--ema", action="store_true", help="accepted; weights.npz already holds the averaged weights")
    p.add_argument("--T", type=int, default=100)
    p.add_argument("--step_lr", type=float, default=2e-5)
    p.add_argument("--return_last_point", action="store_false")
    p.add_argument("--height", type=int, default=96)
    p.add_argument("--width", type=int, default=64)
    p.add_argument("--scale", type=str, default="dB")
    p.add_argument("--n_filters", type=int, default=192)
    p.add_argument("--sigma1", type=float, default=1.0)
    p.add_argument("--sigmaL", type=float, default=0.01)
    p.add_argument("--num_classes", type=int, default=10)
    p.add_argument("--use_logit", action="store_true")
    p.add_argument("--alpha", type=float, default=1e-6)
    p.add_argument("--optimizer", type=str, default="adam")
    # additions
    p.add_argument("--random_init", type=int, default=None, metavar="SEED")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--fast", action="store_true", help="one bf16 tensor-core product per convolution (default: the parity mode)")
    return p


def main(args):
    import torch
    if args.config is not None:                                   # ncsn_generate_samples.py:26-32
        new_args = get_config(args.config)
        for k in ("dataset", "filename", "RESTORE", "n_samples", "random_init", "seed", "fast"):
            setattr(new_args, k, getattr(args, k))
        for k, v in vars(build_parser().parse_args([args.RESTORE])).items():      # defaults for keys the YAML lacks
            if not hasattr(new_args, k):
                setattr(new_args, k, v)
        new_args.num_classes = int(new_args.num_classes)
        args = new_args
    print("SAMPLING PARAMETERS")
    print("\t " + "".join("{} = {} \n\t ".format(k, v) for k, v in vars(args).items()))
    print("_" * 100)
    sigmas_np = get_sigmas(args.sigma1, args.sigmaL, args.num_classes,
                           progression=getattr(args, "progression", "geometric"))
    if args.dataset in ("mnist", "cifar10"):
        raise NotImplementedError("image toy datasets are outside the separation hot path (SURVEY.md section 2)")
    args.data_shape = [args.height, args.width, 1]
    args.data_type = "melspec"
    if args.scale == "dB":
        args.minval, args.maxval = -100.0, 20.0
    elif args.scale == "power":
        args.minval, args.maxval = 1e-10, 100.0
    else:
        raise ValueError("scale should be 'power' or 'dB'")

    def post_processing(x):                                       # ncsn_generate_samples.py:68-79
        if args.use_logit:
            x = 1.0 / (1.0 + np.exp(-x))
            x = (x - args.alpha) / (1.0 - 2.0 * args.alpha)
        x = x * (args.maxval - args.minval) + args.minval
        return np.clip(x, args.minval, args.maxval)

    abs_restore_path = os.path.abspath(args.RESTORE)
    params = None
    if args.random_init is None:
        with np.load(os.path.join(abs_restore_path, "weights.npz")) as z:
            params = {k: z[k] for k in z.files}
    if args.version == "v2":
        model = get_uncompiled_model_v2(args, sigmas=sigmas_np, params=params, seed=args.random_init)
    else:
        model = get_uncompiled_model(args, params=params, seed=args.random_init)
    setUp_optimizer(args)
    print("Weights loaded")

    print("Start Generating {} samples....".format(args.n_samples))
    t0 = time.time()
    g = torch.Generator().manual_seed(int(args.seed))
    x_mod = torch.rand([args.n_samples] + args.data_shape, generator=g)
    if args.use_logit:
        x_mod = (1.0 - 2 * args.alpha) * x_mod + args.alpha
        x_mod = torch.log(x_mod) - torch.log(1.0 - x_mod)
    x_arr = anneal_langevin_dynamics(x_mod, args.data_shape, model, args.n_samples, sigmas_np, n_steps_each=args.T,
                                     step_lr=args.step_lr, return_arr=True, verbose=True, seed=int(args.seed))
    x_arr = post_processing(x_arr)
    print("Done. Duration: {} seconds".format(round(time.time() - t0, 2)))
    print("Shape: {}".format(x_arr.shape))
    if args.filename is None:
        head, ckpt_name = os.path.split(abs_restore_path)
        args.filename = os.path.join(head, "generated_samples" + "_" + ckpt_name)
    try:
        np.save(args.filename, x_arr)
        print("Generated Samples saved at {}".format(args.filename + ".npy"))
    except FileNotFoundError:
        np.save("generated_samples", x_arr)
        print("Generated Samples saved at {}".format("generated_samples.npy"))
    return x_arr


if __name__ == "__main__":
    main(build_parser().parse_args())
