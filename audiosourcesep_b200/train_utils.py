"""Host-side mirror of the hot-path part of the reference's ``train_utils.py``.

``setUp_optimizer`` (train_utils.py:23-41), ``get_config`` / ``dict2namespace`` (:114-131) and a checkpoint manager with
the reference's ``setUp_checkpoint`` call shape (:62-75) over ``.npz`` weight containers (keeping ``max_to_keep`` files,
``latest_checkpoint`` / ``save`` / ``restore``).  TensorBoard and the matplotlib helpers (:44-59, :78-111) are outside
the hot path.
"""
from __future__ import annotations

import glob
import os
import re
from typing import Optional

import numpy as np

from .config import dict2namespace, get_config                      # noqa: F401  (re-exported, train_utils.py:114-131)
from .train_glow import setUp_optimizer                             # noqa: F401  (train_utils.py:23-41)


class CheckpointManager:
    """``tf.train.CheckpointManager(ckpt, path, max_to_keep)`` over ``ckpt-<n>.npz`` files: ``variables=model.variables``
    are stored under their parameter names (audiosourcesep_b200/weights.py); the optimizer slots live in the library
    handle and are not persisted (a restored run restarts its moment estimates, as a reference run restored with
    ``expect_partial`` does)."""

    def __init__(self, model, path: str = "./tf_ckpts", max_to_keep: int = 5):
        self.model, self.path, self.max_to_keep = model, path, int(max_to_keep)

    def _files(self):
        fs = glob.glob(os.path.join(self.path, "ckpt-*.npz"))
        return sorted(fs, key=lambda f: int(re.search(r"ckpt-(\d+)\.npz$", f).group(1)))

    @property
    def latest_checkpoint(self) -> Optional[str]:
        fs = self._files()
        return fs[-1] if fs else None

    def save(self) -> str:
        os.makedirs(self.path, exist_ok=True)
        fs = self._files()
        n = 1 + (int(re.search(r"ckpt-(\d+)\.npz$", fs[-1]).group(1)) if fs else 0)
        if hasattr(self.model, "sync_host"):
            self.model.sync_host()
        out = os.path.join(self.path, f"ckpt-{n}.npz")
        np.savez(out, **self.model.variables)
        for old in self._files()[:-self.max_to_keep]:
            os.remove(old)
        return out

    def restore(self, path: Optional[str] = None) -> Optional[str]:
        path = path or self.latest_checkpoint
        if path is None:
            return None
        with np.load(path) as z:
            self.model.set_params({k: z[k] for k in z.files})
        self.model.prepare()
        return path


def setUp_checkpoint(mirrored_strategy, model, optimizer, max_to_keep=5, path="./tf_ckpts"):
    """reference: train_utils.py:62-75 -- returns (ckpt, manager); here both are the manager."""
    manager = CheckpointManager(model, path=path, max_to_keep=max_to_keep)
    return manager, manager
