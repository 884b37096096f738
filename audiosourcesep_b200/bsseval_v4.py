"""BSS Eval v4 -- host-side mirror of the reference's ``bsseval_v4.py`` (same function names, arguments and return
values); the correlations, the block-Toeplitz solves and the four-way decomposition run in libasep.so (csrc/bsseval.cu).

Kept on the host, restated from the reference: input validation (bsseval_v4.py:21-70), framing (:377-418), the
permutation search and the result selection (:202-213, :281-300).  Limits: mono images (nchan = 1) -- what the separation
path produces (flattened mel patches, SURVEY 8(d), or 16 kHz mono audio).
"""
from __future__ import annotations

import ctypes
import itertools

import numpy as np
import torch

from . import _lib

MAX_SOURCES = 100


def validate(reference_sources, estimated_sources):
    """bsseval_v4.py:21-70."""
    if reference_sources.shape != estimated_sources.shape:
        raise ValueError("The shape of estimated sources and the true sources should match. reference_sources.shape = {}, "
                         "estimated_sources.shape = {}".format(reference_sources.shape, estimated_sources.shape))
    if reference_sources.ndim > 3 or estimated_sources.ndim > 3:
        raise ValueError("The number of dimensions is too high (must be less than 3). reference_sources.ndim = {}, "
                         "estimated_sources.ndim = {}".format(reference_sources.ndim, estimated_sources.ndim))
    if estimated_sources.shape[0] > MAX_SOURCES or reference_sources.shape[0] > MAX_SOURCES:
        raise ValueError("The supplied matrices should be of shape (nsrc, nsampl, nchan) but the number of sources exceeds "
                         "bsseval.MAX_SOURCES = {}".format(MAX_SOURCES))


class Framing:
    """bsseval_v4.py:377-418: windows [t*hop, min(t*hop + window, length))."""

    def __init__(self, window, hop, length):
        self.window, self.hop, self.length = window, hop, length

    @property
    def nwin(self):
        if self.window < self.length:
            return int(np.floor((self.length - self.window + self.hop) / self.hop))
        return 1

    def __iter__(self):
        for t in range(self.nwin):
            start, stop = t * self.hop, min(t * self.hop + self.window, self.length)
            start = 0 if (np.isnan(start) or np.isinf(start)) else start
            stop = self.length if (np.isnan(stop) or np.isinf(stop)) else stop
            yield slice(int(np.floor(start)), int(np.floor(stop)))


def _device_eval(refs: torch.Tensor, ests: torch.Tensor, filters_len, filt, wins, sources_version):
    nsrc, nwin = refs.shape[0], len(wins)
    out = torch.empty((4, nsrc, nsrc, nwin), dtype=torch.float64, device=refs.device)
    w0 = (ctypes.c_int64 * nwin)(*[w.start for w in wins])
    w1 = (ctypes.c_int64 * nwin)(*[w.stop for w in wins])
    dr, de, do = _lib.dl(refs), _lib.dl(ests), _lib.dl(out)
    _lib.check(_lib.load().asep_bss_eval(dr.ptr, de.ptr, int(filters_len), int(filt.start), int(filt.stop), w0, w1, nwin,
                                         int(bool(sources_version)), do.ptr, _lib.stream_ptr()))
    return out.cpu().numpy()


def bss_eval(reference_sources, estimated_sources, window=2 * 44100, hop=1.5 * 44100, compute_permutation=False,
             filters_len=512, framewise_filters=False, bsseval_sources_version=False, device=None):
    """reference: bsseval_v4.py:79-300.  Returns (sdr, isr, sir, sar, perm) with the reference's shapes."""
    estimated_sources = np.atleast_3d(np.asarray(estimated_sources))
    reference_sources = np.atleast_3d(np.asarray(reference_sources))
    validate(reference_sources, estimated_sources)
    if reference_sources.size == 0 or estimated_sources.size == 0:
        return np.array([]), np.array([]), np.array([]), np.array([]), np.array([])
    nsrc, nsampl, nchan = estimated_sources.shape
    if nchan != 1:
        raise NotImplementedError("the device implementation evaluates mono images (nchan = 1)")
    dev = torch.device("cuda", _lib.init(device))
    refs = torch.as_tensor(np.ascontiguousarray(reference_sources[..., 0], dtype=np.float64)).to(dev)
    ests = torch.as_tensor(np.ascontiguousarray(estimated_sources[..., 0], dtype=np.float64)).to(dev)
    if compute_permutation:
        candidate_permutations = np.array(list(itertools.permutations(list(range(nsrc)))))
    else:
        candidate_permutations = np.array(np.arange(nsrc))[None, :]
    wins = list(Framing(window, hop, nsampl))
    nwin = len(wins)
    if not framewise_filters:                                    # one set of filters from the whole signals (:238-242)
        s_r = _device_eval(refs, ests, filters_len, slice(0, nsampl), wins, bsseval_sources_version)
    else:                                                        # new filters for every window (:247-250)
        s_r = np.concatenate([_device_eval(refs, ests, filters_len, w, [w], bsseval_sources_version) for w in wins], axis=-1)
    SIR = 2
    # the reference only evaluates the (jtrue, jest) pairs its candidate permutations contain; the others stay unset
    if framewise_filters:
        mean_sir = np.empty((len(candidate_permutations), nwin))
        axis_mean = 0
    else:
        mean_sir = np.empty((len(candidate_permutations), 1))
        axis_mean = None
    dum = np.arange(nsrc)
    for i, perm in enumerate(candidate_permutations):
        mean_sir[i] = np.mean(s_r[SIR, dum, perm, :], axis=axis_mean)
    popt = candidate_permutations[np.argmax(mean_sir, axis=0)].T
    if not framewise_filters:
        result = s_r[:, dum, popt[:, 0], :]
    else:
        result = np.empty((4, nsrc, nwin))
        for m, t in itertools.product(range(4), range(nwin)):
            result[m, :, t] = s_r[m, dum, popt[:, t], t]
    return result[0], result[1], result[2], result[3], popt


def bss_eval_sources(reference_sources, estimated_sources, compute_permutation=True):
    """BSS Eval v3 bss_eval_sources (bsseval_v4.py:303-321)."""
    sdr, _, sir, sar, perm = bss_eval(reference_sources, estimated_sources, window=np.inf, hop=np.inf,
                                      compute_permutation=compute_permutation, filters_len=512, framewise_filters=True,
                                      bsseval_sources_version=True)
    return sdr, sir, sar, perm


def bss_eval_sources_framewise(reference_sources, estimated_sources, window=30 * 44100, hop=15 * 44100,
                               compute_permutation=False):
    """bsseval_v4.py:324-342."""
    sdr, _, sir, sar, perm = bss_eval(reference_sources, estimated_sources, window=window, hop=hop,
                                      compute_permutation=compute_permutation, filters_len=512, framewise_filters=True,
                                      bsseval_sources_version=True)
    return sdr, sir, sar, perm


def bss_eval_images(reference_sources, estimated_sources, compute_permutation=True):
    """bsseval_v4.py:345-358."""
    return bss_eval(reference_sources, estimated_sources, window=np.inf, hop=np.inf, compute_permutation=compute_permutation,
                    filters_len=512, framewise_filters=True, bsseval_sources_version=False)


def bss_eval_images_framewise(reference_sources, estimated_sources, window=30 * 44100, hop=15 * 44100,
                              compute_permutation=False):
    """bsseval_v4.py:361-374."""
    return bss_eval(reference_sources, estimated_sources, window=window, hop=hop, compute_permutation=compute_permutation,
                    filters_len=512, framewise_filters=True, bsseval_sources_version=False)
