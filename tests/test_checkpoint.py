"""Checkpoint formats (pamrec_b200/checkpoint.py): CRC-32C known answers, the safetensors file against the `safetensors` package,
the TensorFlow tensor-bundle writer byte for byte on a hand-assembled case, round trips, and corruption detection."""
import os
import struct

import numpy as np
import pytest

from pamrec_b200 import checkpoint as CK


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as G
    G.build()
    from pamrec_b200 import _lib
    return _lib.load()


def _tensors(seed=0):
    rng = np.random.default_rng(seed)
    t = {f"sequential/pamrec/num_blocks_{b}/self_attention/{q}_timeaware_embedding": rng.standard_normal((10, 1600)).astype(np.float32)
         for b in range(2) for q in "QKV"}
    t.update({f"sequential/logit_fcn/nn_part/w_nn_layer{i}": rng.standard_normal((84, 100)).astype(np.float32) for i in range(20)})
    t["sequential/embedding/item_embedding"] = rng.standard_normal((1000, 16)).astype(np.float32)
    t["scalar"] = np.float32(3.5)
    t["empty"] = np.zeros((0, 4), np.float32)
    t["steps"] = np.arange(7, dtype=np.int64)
    t["zeta/é"] = np.asarray([1, 2, 3], np.int32)          # non-ASCII name: keys sort by UTF-8 bytes
    return t


def _same(a, b):
    assert sorted(a) == sorted(b)
    for k in a:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.dtype == y.dtype and x.shape == y.shape and x.tobytes() == y.tobytes(), k


def test_crc32c_known_answers(lib):
    # RFC 3720 appendix B.4 and the classic check value
    cases = [(b"123456789", 0xE3069283), (bytes(32), 0x8A9136AA), (b"\xff" * 32, 0x62A8AB43), (bytes(range(32)), 0x46DD794E),
             (bytes(range(31, -1, -1)), 0x113FDB5C), (b"", 0)]
    for data, want in cases:
        assert CK.crc32c(data) == want
        assert lib.pamrec_crc32c_portable(0, data, len(data)) == want
    blob = np.random.default_rng(1).integers(0, 256, 100_003, dtype=np.uint8).tobytes()
    whole = CK.crc32c(blob)
    assert lib.pamrec_crc32c_portable(0, blob, len(blob)) == whole
    for cut in (0, 1, 7, 8, 9, 50_000, len(blob)):            # continuation: crc(a + b) = crc(b, crc(a))
        assert CK.crc32c(blob[cut:], CK.crc32c(blob[:cut])) == whole
    # the masking of tensorflow/core/lib/hash/crc32c.h (leveldb's): invertible, and not the identity
    for c in (0, 1, 0xE3069283, 0xffffffff):
        assert CK._unmask(CK._mask(c)) == c and CK._mask(c) != c
    assert CK._mask(0) == 0xa282ead8


def test_safetensors_round_trip_and_the_official_reader(tmp_path):
    t = _tensors()
    p = str(tmp_path / "m.safetensors")
    CK.save_safetensors(p, t, metadata={"step": 7})
    got, meta = CK.load_safetensors(p)
    _same(t, got)
    assert meta == {"step": "7"}
    st = pytest.importorskip("safetensors.numpy")
    _same(t, st.load_file(p))                                   # our file through the package
    q = str(tmp_path / "theirs.safetensors")
    st.save_file({k: np.asarray(v) for k, v in t.items()}, q)
    _same(t, CK.load_safetensors(q)[0])                         # the package's file through our reader
    # damaged files are refused
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-5])
    with pytest.raises(ValueError):
        CK.load_safetensors(p)
    open(p, "wb").write(struct.pack("<Q", 1 << 40) + raw[8:])
    with pytest.raises(ValueError):
        CK.load_safetensors(p)


def test_tf_bundle_bytes_of_a_small_case(tmp_path, lib):
    """One float32 vector named "a": every byte of both files follows from the format description."""
    a = np.asarray([1.0, -2.5], np.float32)
    prefix = str(tmp_path / "ck")
    CK.save_tf_bundle(prefix, {"a": a})
    assert open(prefix + ".data-00000-of-00001", "rb").read() == a.tobytes()
    crc = CK._mask(CK.crc32c(a.tobytes()))
    header = b"\x08\x01\x1a\x02\x08\x01"                         # num_shards: 1, version { producer: 1 }
    entry = b"\x08\x01" + b"\x12\x04\x12\x02\x08\x02" + b"\x28\x08" + b"\x35" + struct.pack("<I", crc)   # DT_FLOAT, shape {dim {size: 2}}, size: 8
    block = (b"\x00\x00" + bytes([len(header)]) + header +       # key "" (shared 0, unshared 0)
             b"\x00\x01" + bytes([len(entry)]) + b"a" + entry +  # key "a"
             struct.pack("<II", 0, 1))                           # one restart point at 0
    trailer = lambda c: b"\x00" + struct.pack("<I", CK._mask(CK.crc32c(c + b"\x00")))
    meta = struct.pack("<II", 0, 1)
    meta_off = len(block) + 5
    index = b"\x00\x01\x02" + b"a" + bytes([0, len(block)]) + struct.pack("<II", 0, 1)     # last key of the block -> handle (0, size)
    idx_off = meta_off + len(meta) + 5
    foot = bytes([meta_off, len(meta), idx_off, len(index)])
    want = block + trailer(block) + meta + trailer(meta) + index + trailer(index) + foot + bytes(40 - len(foot)) + bytes.fromhex("57fb808b247547db")
    assert open(prefix + ".index", "rb").read() == want
    got = CK.load_tf_bundle(prefix)
    assert list(got) == ["a"] and got["a"].tobytes() == a.tobytes()


def test_tf_bundle_round_trip_many_blocks_and_corruption(tmp_path, monkeypatch):
    t = _tensors(3)
    prefix = str(tmp_path / "big")
    CK.save_tf_bundle(prefix, t)
    _same(t, CK.load_tf_bundle(prefix))
    one_block = open(prefix + ".index", "rb").read()
    monkeypatch.setattr(CK, "_BLOCK_SIZE", 200)                 # several data blocks, restart points inside them
    CK.save_tf_bundle(prefix, t)
    assert open(prefix + ".index", "rb").read() != one_block
    rows = CK._read_table(prefix + ".index")
    assert [k for k, _ in rows] == sorted([b""] + [k.encode() for k in t])
    _same(t, CK.load_tf_bundle(prefix))
    # a flipped bit in the data file or in the index is detected
    data = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data, "rb").read())
    raw[100] ^= 1
    open(data, "wb").write(raw)
    with pytest.raises(ValueError, match="checksum"):
        CK.load_tf_bundle(prefix)
    assert len(CK.load_tf_bundle(prefix, verify=False)) == len(t)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[10] ^= 1
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(ValueError, match="checksum"):
        CK.load_tf_bundle(prefix)
    idx[10] ^= 1
    idx[-1] ^= 1
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(ValueError, match="magic"):
        CK.load_tf_bundle(prefix)


def test_tf_bundle_reader_skips_what_it_cannot_hold(tmp_path):
    """A TF2-written checkpoint also carries a DT_STRING object graph; partitioned variables carry slices."""
    prefix = str(tmp_path / "x")
    CK.save_tf_bundle(prefix, {"w": np.ones(3, np.float32)})
    rows = CK._read_table(prefix + ".index")
    string_entry = b"\x08\x07" + b"\x12\x00" + b"\x28\x03" + b"\x35" + struct.pack("<I", 0)     # DT_STRING = 7
    CK._write_table(prefix + ".index", sorted(rows + [(b"_CHECKPOINTABLE_OBJECT_GRAPH", string_entry)]))
    got = CK.load_tf_bundle(prefix)
    assert list(got) == ["w"] and got.skipped == ["_CHECKPOINTABLE_OBJECT_GRAPH"]
    sliced = rows[1][1] + b"\x3a\x00"                                                            # slices { }
    CK._write_table(prefix + ".index", [rows[0], (b"w", sliced)])
    with pytest.raises(ValueError, match="partitioned"):
        CK.load_tf_bundle(prefix)


@pytest.mark.parametrize("fmt", CK.FORMATS)
def test_save_load_remove_in_every_format(tmp_path, fmt):
    t = {k: v for k, v in _tensors(5).items() if k not in ("zeta/é",)}
    opt = {"sequential/embedding/item_embedding/Adam": np.full((1000, 16), 0.25, np.float32),
           "sequential/embedding/item_embedding/Adam_1": np.full((1000, 16), 0.5, np.float32), "step": 12}
    path = str(tmp_path / "dir" / "step_40")
    os.makedirs(os.path.dirname(path))
    assert not CK.exists(path)
    CK.save(path, t, fmt=fmt, optimizer=opt)
    assert CK.exists(path)
    variables, got_opt = CK.load(path)
    _same(t, variables)
    assert int(got_opt["step"]) == 12 and got_opt["sequential/embedding/item_embedding/Adam_1"][3, 3] == 0.5
    CK.save(path + "0", t, fmt=fmt)                              # "step_400": remove("step_40") must not touch it
    assert CK.load(path + "0")[1] is None
    CK.remove(path)
    assert not CK.exists(path) and CK.exists(path + "0")
    with pytest.raises(FileNotFoundError):
        CK.load(path)
    with pytest.raises(ValueError):
        CK.save(path, t, fmt="hdf5")
