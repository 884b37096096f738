"""TensorFlow checkpoint interchange: reader / writer of the TensorBundle format ``tf.train.Checkpoint`` saves.

The reference persists its models with ``tf.train.Checkpoint(variables=model.variables, optimizer=optimizer)`` +
``CheckpointManager`` (train_utils.py:62-75), restores them with ``ckpt.restore(path)`` (run_basis_sep.py:28-38) and, for
the noise-conditioned Glow priors, keeps one checkpoint directory per noise level,
``<RESTORE>/sigma_<round(sigma, 2)>/tf_ckpts`` (run_basis_sep.py:284-285, train_noisy_glow.py:309-358).  This module
reads and writes those files WITHOUT TensorFlow, so that reference-trained weights can drop into libasep.so and weights
trained here can go back:

* ``<prefix>.index``  -- a LevelDB-format sorted string table: key -> serialized ``BundleEntryProto`` (dtype, shape,
  shard, offset, size, masked crc32c); the empty key holds the ``BundleHeaderProto``.  Blocks are prefix-compressed
  (shared / non_shared / value_length varints + restart array), each followed by a 1-byte compression tag and a masked
  crc32c; a 48-byte footer (metaindex handle, index handle, magic 0xdb4775248b80fb57) closes the file.
* ``<prefix>.data-00000-of-00001`` -- the raw little-endian tensor bytes.
* object-graph keys: the list passed as ``variables=`` is saved as ``variables/<i>/.ATTRIBUTES/VARIABLE_VALUE`` with
  ``i`` the position in ``model.variables``; optimizer slots live under ``optimizer/...`` and
  ``variables/<i>/.OPTIMIZER_SLOT/...`` and are ignored on import.

STATUS: format-checked only.  There is no checkpoint in the reference tree ("ground facts") and TensorFlow
cannot be installed here, so the byte format is pinned by hand-built fixtures (tests/test_tf_checkpoint.py) and by a
write -> read round trip, NOT by a file TensorFlow wrote.  The order of ``model.variables`` (tf.Module attribute
traversal of the TFP distribution / Keras model) is likewise restated, not observed: ``variable_order`` returns this
package's construction order and ``import_variables`` validates the shape of every entry against it; a different order
can be supplied as a list of parameter names (``order=``).
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

MAGIC = 0xDB4775248B80FB57
VAR_KEY = "variables/{}/.ATTRIBUTES/VARIABLE_VALUE"
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"), 6: np.dtype("i1"),
           9: np.dtype("<i8"), 10: np.dtype("bool"), 19: np.dtype("<f2"), 17: np.dtype("<u2"), 22: np.dtype("<u4"), 23: np.dtype("<u8")}
_DTYPE_IDS = {v: k for k, v in _DTYPES.items()}


# ------------------------------------------------------------------ crc32c (Castagnoli), masked as LevelDB / TF do
def _crc_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_TAB = _crc_table()


def crc32c(data, crc: int = 0) -> int:
    """CRC-32C of ``data``; buffers of 4 KiB and more go through libasep.so's slicing-by-8 host routine."""
    data = bytes(data)
    if len(data) >= 4096:
        try:
            from . import _lib
            return int(_lib.load().asep_crc32c(data, len(data), crc)) & 0xFFFFFFFF
        except (OSError, RuntimeError, AttributeError):
            pass
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = _TAB[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------ varints / minimal protobuf
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift, val = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_fields(buf: bytes) -> List[Tuple[int, int, object]]:
    """[(field number, wire type, value)] of one protobuf message (varint, fixed32/64, length-delimited)."""
    out, pos = [], 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _get_varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            val = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.append((num, wt, val))
    return out


def _parse_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    for num, _, val in _parse_fields(buf):
        if num == 2:                                            # repeated Dim dim = 2 { int64 size = 1; }
            size = 0
            for n2, _, v2 in _parse_fields(val):
                if n2 == 1:
                    size = v2
            dims.append(int(size))
    return tuple(dims)


def _entry_proto(dtype_id: int, shape: Sequence[int], offset: int, size: int, crc: int) -> bytes:
    shp = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    out = b"\x08" + _put_varint(dtype_id) + b"\x12" + _put_varint(len(shp)) + shp
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return out


# ------------------------------------------------------------------ sorted string table
def _read_block(buf: bytes, offset: int, size: int, verify: bool) -> bytes:
    block, tag = buf[offset:offset + size], buf[offset + size]
    if verify:
        want = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if mask_crc(crc32c(bytes(buf[offset:offset + size + 1]))) != want:
            raise ValueError("checkpoint index: block checksum mismatch")
    if tag != 0:
        raise NotImplementedError("compressed index blocks (snappy) are not supported; TensorBundle writes them uncompressed")
    return bytes(block)


def _block_entries(block: bytes) -> Iterable[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _build_block(entries: Sequence[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def read_index(path: str, verify: bool = True) -> Dict[str, bytes]:
    """key -> serialized BundleEntryProto (the header sits under the empty key)."""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError(f"{path}: not a TensorFlow checkpoint index (bad table magic)")
    footer = buf[-48:]
    _, p = _get_varint(footer, 0)
    _, p = _get_varint(footer, p)                               # metaindex handle (unused)
    ioff, p = _get_varint(footer, p)
    isize, p = _get_varint(footer, p)
    out = {}
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, q = _get_varint(handle, 0)
        size, _ = _get_varint(handle, q)
        for k, v in _block_entries(_read_block(buf, off, size, verify)):
            out[k.decode("utf-8")] = v
    return out


def read_checkpoint(prefix: str, verify: bool = True, keys: Optional[Iterable[str]] = None) -> Dict[str, np.ndarray]:
    """All numeric tensors of ``<prefix>.index`` / ``<prefix>.data-*`` as numpy arrays (string tensors such as the
    object graph are skipped)."""
    index = read_index(prefix + ".index", verify)
    header = index.pop("", b"")
    num_shards = 1
    for num, _, val in _parse_fields(header):
        if num == 1:
            num_shards = int(val)
        if num == 2 and val != 0:
            raise NotImplementedError("big-endian checkpoints are not supported")
    shards = {}
    out = {}
    want = None if keys is None else set(keys)
    for key, raw in index.items():
        if want is not None and key not in want:
            continue
        f = {num: val for num, _, val in _parse_fields(raw)}
        dt = _DTYPES.get(int(f.get(1, 0)))
        if dt is None or 7 in f:                                # strings / sliced tensors: not weights
            continue
        shape = _parse_shape(f.get(2, b""))
        shard, offset, size = int(f.get(3, 0)), int(f.get(4, 0)), int(f.get(5, 0))
        if shard not in shards:
            shards[shard] = np.memmap(f"{prefix}.data-{shard:05d}-of-{num_shards:05d}", dtype=np.uint8, mode="r")
        data = shards[shard][offset:offset + size]
        if verify and 6 in f and mask_crc(crc32c(data.tobytes())) != int(f[6]):
            raise ValueError(f"{prefix}: tensor '{key}' fails its crc32c")
        out[key] = np.frombuffer(data.tobytes(), dtype=dt).reshape(shape).copy()
    return out


def write_checkpoint(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Writes ``<prefix>.index`` + ``<prefix>.data-00000-of-00001`` holding ``tensors`` (key -> array)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)) or ".", exist_ok=True)
    entries, offset = [], 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for key in sorted(tensors, key=lambda k: k.encode("utf-8")):
            a = np.asarray(tensors[key], order="C")            # (ascontiguousarray would turn a scalar into shape [1])
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            dtype_id = _DTYPE_IDS.get(np.dtype(dt.str.replace(">", "<")) if dt.kind not in "b" else np.dtype("bool"))
            if dtype_id is None:
                raise TypeError(f"tensor '{key}': dtype {a.dtype} has no checkpoint encoding here")
            raw = a.astype(dt, copy=False).tobytes()
            f.write(raw)
            entries.append((key.encode("utf-8"), _entry_proto(dtype_id, a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"                  # num_shards = 1, endianness LITTLE (default), version.producer = 1
    entries = [(b"", header)] + entries
    out = bytearray()

    def emit(block: bytes) -> bytes:
        off = len(out)
        out.extend(block + b"\x00")
        out.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    index_entries = []
    for i in range(0, len(entries), 256):
        chunk = entries[i:i + 256]
        index_entries.append((chunk[-1][0], emit(_build_block(chunk))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index_entries, restart_interval=1))
    footer = meta + idx
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC))
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))


# ------------------------------------------------------------------ variables/<i> <-> parameter names
def variable_order(shapes: Dict[str, Tuple[int, ...]]) -> List[str]:
    """Names in the order this package creates them (weights.py: glow_param_shapes / ncsn_param_shapes walk the
    reference constructors).  RESTATED, not observed: see the module docstring."""
    return list(shapes)


def import_variables(prefix: str, shapes: Dict[str, Tuple[int, ...]], order: Optional[Sequence[str]] = None,
                     verify: bool = True) -> Dict[str, np.ndarray]:
    """``variables/<i>`` of a reference checkpoint -> {parameter name: array}; every entry's shape is checked."""
    order = list(order) if order is not None else variable_order(shapes)
    ck = read_checkpoint(prefix, verify, keys=[VAR_KEY.format(i) for i in range(len(order))])
    out = {}
    for i, name in enumerate(order):
        key = VAR_KEY.format(i)
        if key not in ck:
            raise KeyError(f"{prefix}: '{key}' missing ({len(order)} variables expected)")
        a = ck[key]
        want = tuple(shapes[name])
        if int(np.prod(a.shape)) != int(np.prod(want)) or (a.ndim == len(want) and tuple(a.shape) != want):
            raise ValueError(f"{prefix}: variable {i} has shape {tuple(a.shape)}, parameter '{name}' expects {want} -- "
                             "the model.variables order differs; pass order=[...]")
        out[name] = a.reshape(want).astype(np.float32)
    return out


def export_variables(prefix: str, params: Dict[str, np.ndarray], order: Optional[Sequence[str]] = None) -> None:
    """{parameter name: array} -> a checkpoint whose ``variables/<i>`` follow ``order`` (plus ``save_counter``)."""
    order = list(order) if order is not None else list(params)
    tensors = {VAR_KEY.format(i): np.asarray(params[n], np.float32) for i, n in enumerate(order)}
    tensors["save_counter/.ATTRIBUTES/VARIABLE_VALUE"] = np.asarray(1, np.int64)
    write_checkpoint(prefix, tensors)


def latest_checkpoint(directory: str) -> Optional[str]:
    """``tf.train.latest_checkpoint``: the prefix named by the ``checkpoint`` state file, else the highest ckpt-N."""
    state = os.path.join(directory, "checkpoint")
    if os.path.exists(state):
        for line in open(state):
            if line.startswith("model_checkpoint_path:"):
                name = line.split(":", 1)[1].strip().strip('"')
                return name if os.path.isabs(name) else os.path.join(directory, name)
    best = None
    for f in os.listdir(directory) if os.path.isdir(directory) else []:
        if f.startswith("ckpt-") and f.endswith(".index"):
            n = int(f[5:-6])
            if best is None or n > best[0]:
                best = (n, os.path.join(directory, f[:-6]))
    return None if best is None else best[1]
