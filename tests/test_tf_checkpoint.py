"""TensorFlow checkpoint (TensorBundle) interchange -- format-checked only (no TensorFlow here, no checkpoint in the
reference tree): hand-built bytes pin the table / protobuf / checksum encodings, a write -> read round trip pins the rest.
Reference: train_utils.py:62-75 (tf.train.Checkpoint(variables=model.variables, optimizer=...)), run_basis_sep.py:28-38."""
import os
import struct

import numpy as np
import pytest

from audiosourcesep_b200 import tf_checkpoint as tc


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors of CRC-32C (Castagnoli)
    assert tc.crc32c(b"") == 0
    assert tc.crc32c(b"123456789") == 0xE3069283
    assert tc.crc32c(bytes(32)) == 0x8A9136AA
    assert tc.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert tc.crc32c(bytes(range(32))) == 0x46DD794E
    # the library routine (slicing-by-8, >= 4 KiB) and the byte-wise table agree, also when chained
    blob = np.random.default_rng(0).integers(0, 256, 10000, dtype=np.uint8).tobytes()
    slow = 0xFFFFFFFF
    for b in blob:
        slow = tc._TAB[(slow ^ b) & 0xFF] ^ (slow >> 8)
    assert tc.crc32c(blob) == slow ^ 0xFFFFFFFF
    assert tc.crc32c(blob[5000:], tc.crc32c(blob[:5000])) == tc.crc32c(blob)
    # LevelDB's mask: rotate right by 15 and add a constant
    assert tc.mask_crc(0) == 0xA282EAD8 and tc.mask_crc(0x00008000) == 0xA282EAD9


def _hand_built_bundle(tmp_path):
    """One float32 [2,3] tensor under variables/0 and an int64 scalar, every byte written explicitly."""
    prefix = str(tmp_path / "ckpt-1")
    w = np.arange(6, dtype="<f4").reshape(2, 3)
    sc = np.asarray(7, "<i8")
    data = w.tobytes() + sc.tobytes()
    open(prefix + ".data-00000-of-00001", "wb").write(data)

    def varint(v):                                               # protobuf / LevelDB base-128 varint, written out here
        out = b""
        while v >= 0x80:
            out += bytes([(v & 0x7F) | 0x80])
            v >>= 7
        return out + bytes([v])

    def entry(dtype, dims, offset, size, payload):
        shape = b"".join(b"\x12\x02\x08" + bytes([d]) for d in dims)          # Dim { size = d }
        out = b"\x08" + bytes([dtype]) + b"\x12" + bytes([len(shape)]) + shape
        if offset:
            out += b"\x20" + bytes([offset])
        return out + b"\x28" + bytes([size]) + b"\x35" + struct.pack("<I", tc.mask_crc(tc.crc32c(payload)))

    k1 = b"save_counter/.ATTRIBUTES/VARIABLE_VALUE"
    k2 = b"variables/0/.ATTRIBUTES/VARIABLE_VALUE"
    ents = [(b"", b"\x08\x01\x1a\x02\x08\x01"), (k1, entry(9, [], 24, 8, sc.tobytes())), (k2, entry(1, [2, 3], 0, 24, w.tobytes()))]
    # data block: prefix-compressed entries (restart interval 16 -> one restart), then restarts + count
    blk, prev = b"", b""
    for i, (k, v) in enumerate(ents):
        shared = 0
        if i:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        blk += bytes([shared, len(k) - shared, len(v)]) + k[shared:] + v
        prev = k
    blk += struct.pack("<II", 0, 1)
    out = blk + b"\x00" + struct.pack("<I", tc.mask_crc(tc.crc32c(blk + b"\x00")))
    meta = struct.pack("<II", 0, 1)
    meta_off = len(out)
    out += meta + b"\x00" + struct.pack("<I", tc.mask_crc(tc.crc32c(meta + b"\x00")))
    handle = varint(0) + varint(len(blk))
    idx = bytes([0, len(k2), len(handle)]) + k2 + handle + struct.pack("<II", 0, 1)
    idx_off = len(out)
    out += idx + b"\x00" + struct.pack("<I", tc.mask_crc(tc.crc32c(idx + b"\x00")))
    footer = varint(meta_off) + varint(len(meta)) + varint(idx_off) + varint(len(idx))
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    open(prefix + ".index", "wb").write(out)
    return prefix, w


def test_reader_on_hand_built_bundle(tmp_path):
    prefix, w = _hand_built_bundle(tmp_path)
    ck = tc.read_checkpoint(prefix)
    assert sorted(ck) == ["save_counter/.ATTRIBUTES/VARIABLE_VALUE", "variables/0/.ATTRIBUTES/VARIABLE_VALUE"]
    assert np.array_equal(ck["variables/0/.ATTRIBUTES/VARIABLE_VALUE"], w) and ck["variables/0/.ATTRIBUTES/VARIABLE_VALUE"].dtype == np.float32
    assert int(ck["save_counter/.ATTRIBUTES/VARIABLE_VALUE"]) == 7
    got = tc.import_variables(prefix, {"w": (2, 3)})
    assert np.array_equal(got["w"], w)
    with pytest.raises(ValueError, match="order differs"):
        tc.import_variables(prefix, {"w": (4, 2)})
    # a flipped data byte fails the tensor checksum; a flipped index byte fails the block checksum
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[3] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="crc32c"):
        tc.read_checkpoint(prefix)
    assert tc.read_checkpoint(prefix, verify=False)["variables/0/.ATTRIBUTES/VARIABLE_VALUE"].shape == (2, 3)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[10] ^= 1
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError, match="checksum"):
        tc.read_index(prefix + ".index")


def test_writer_output_is_byte_compatible_with_the_hand_built_layout(tmp_path):
    """The writer's index for the same two tensors parses to the same entries as the hand-built one."""
    prefix, w = _hand_built_bundle(tmp_path)
    p2 = str(tmp_path / "mine")
    tc.write_checkpoint(p2, {"variables/0/.ATTRIBUTES/VARIABLE_VALUE": w, "save_counter/.ATTRIBUTES/VARIABLE_VALUE": np.asarray(7, np.int64)})
    a, b = tc.read_index(prefix + ".index"), tc.read_index(p2 + ".index")
    assert a.keys() == b.keys()
    # (the hand-built bundle stores the tensors in the other order, so offsets differ; dtype / shape / size / crc agree)
    for k in a:
        fa = {n: v for n, _, v in tc._parse_fields(a[k]) if n != 4}
        fb = {n: v for n, _, v in tc._parse_fields(b[k]) if n != 4}
        assert fa == fb, k


def test_glow_and_ncsn_round_trip_through_the_bundle_format(tmp_path):
    """All parameters of a (small) Glow and of the NCSN v2 network: export as variables/<i>, import, bit-identical;
    multi-block index (> 256 entries), latest_checkpoint, per-sigma directory layout of run_basis_sep.py:284-285."""
    from audiosourcesep_b200 import GlowConfig, NCSNConfig
    from audiosourcesep_b200.weights import glow_param_shapes, init_glow_params, init_ncsn_params, ncsn_param_shapes
    cfg = GlowConfig(H=16, W=8, C=1, L=3, K=8, n_filters=32)
    p = init_glow_params(cfg, seed=1, mode="perturbed")
    shapes = glow_param_shapes(cfg)
    assert len(shapes) > 256
    d = tmp_path / "sigma_0.6" / "tf_ckpts"
    prefix = str(d / "ckpt-3")
    tc.export_variables(prefix, p, order=list(shapes))
    open(d / "checkpoint", "w").write('model_checkpoint_path: "ckpt-3"\nall_model_checkpoint_paths: "ckpt-3"\n')
    assert tc.latest_checkpoint(str(d)) == prefix
    back = tc.import_variables(prefix, shapes)
    assert back.keys() == p.keys()
    for k in p:
        assert np.array_equal(back[k], p[k]), k
    ncfg = NCSNConfig(version="v2", ngf=32, num_classes=5)
    q = init_ncsn_params(ncfg, seed=2)
    nprefix = str(tmp_path / "ncsn" / "ckpt-1")
    tc.export_variables(nprefix, q, order=list(ncsn_param_shapes(ncfg)))
    back = tc.import_variables(nprefix, ncsn_param_shapes(ncfg))
    assert all(np.array_equal(back[k], q[k]) for k in q)
    os.remove(d / "checkpoint")
    assert tc.latest_checkpoint(str(d)) == prefix


def test_cli_weight_loader_reads_either_container(tmp_path):
    """run_basis_sep's weight loader: weights.npz, or the reference's tf_ckpts/ckpt-N (run_basis_sep.py:28-38, 284-285)."""
    from audiosourcesep_b200 import GlowConfig
    from audiosourcesep_b200.run_basis_sep import _npz_loader, load_weights
    from audiosourcesep_b200.weights import glow_param_shapes, init_glow_params
    cfg = GlowConfig(H=16, W=8, C=1, L=3, K=1, n_filters=8)
    shapes = glow_param_shapes(cfg)
    p = init_glow_params(cfg, seed=0)
    tc.export_variables(str(tmp_path / "sigma_0.6" / "tf_ckpts" / "ckpt-2"), p, order=list(shapes))
    np.savez(tmp_path / "sigma_1.0" / "weights.npz" if (tmp_path / "sigma_1.0").mkdir() is None else None, **p)
    for sigma in (0.5994843, 1.0):
        q = _npz_loader(str(tmp_path), shapes)(sigma)
        assert all(np.array_equal(p[k], q[k]) for k in p)
    with pytest.raises(FileNotFoundError):
        load_weights(str(tmp_path / "nothing"), shapes)



def test_masked_crc32c_against_a_tensorflow_written_file():
    """The one TensorFlow-written binary the reference ships: a TensorBoard event file (TFRecord framing: length, masked
    crc32c of the length, payload, masked crc32c of the payload -- the SAME masked-crc32c convention the TensorBundle
    checkpoint format uses for its tensors and index blocks).  All 82 checksums TensorFlow wrote are reproduced, which pins
    crc32c() / mask_crc() on bytes this repo did not produce (fixture: tests/golden/tf_written_events.tfrecord, copied from
    trained_ncsn/ncsn_piano_192_32_dB_custom_loop/tensorboard_logs/.../events.out.tfevents.*)."""
    import os
    import struct
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_written_events.tfrecord")
    blob = open(path, "rb").read()
    off = n = 0
    while off < len(blob):
        (length,) = struct.unpack("<Q", blob[off:off + 8])
        (c_len,) = struct.unpack("<I", blob[off + 8:off + 12])
        data = blob[off + 12:off + 12 + length]
        (c_data,) = struct.unpack("<I", blob[off + 12 + length:off + 16 + length])
        assert tc.mask_crc(tc.crc32c(blob[off:off + 8])) == c_len
        assert tc.mask_crc(tc.crc32c(data)) == c_data
        off += 16 + length
        n += 1
    assert off == len(blob) and n == 41
