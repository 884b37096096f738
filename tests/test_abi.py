"""The C-ABI library loads without a GPU and exports every symbol include/asep.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "asep.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(asep_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from audiosourcesep_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(names) == _lib.EXPORTED_SYMBOLS     # the Python binding covers the whole header
    assert lib.asep_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    from audiosourcesep_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.asep_init(0) != 0
    assert b"no CPU fallback" in lib.asep_last_error()
    with pytest.raises(RuntimeError):
        _lib.init()
    from audiosourcesep_b200 import GlowConfig
    from audiosourcesep_b200.glow import Glow
    with pytest.raises(RuntimeError):
        Glow(GlowConfig(H=8, W=8, L=2, K=1, n_filters=64))


def test_dlpack_capsule_view():
    import torch
    from audiosourcesep_b200 import _lib
    t = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    d = _lib.dl(t)
    v = d.ptr.contents
    assert v.ndim == 3 and [v.shape[i] for i in range(3)] == [2, 3, 4]
    assert v.dtype.code == 2 and v.dtype.bits == 32 and v.data == t.data_ptr()
    assert not _lib.dl(None).ptr
