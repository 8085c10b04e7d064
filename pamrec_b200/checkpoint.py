"""Checkpoint files keyed by the reference's TF variable names.

The reference saves and restores through ``tf.train.Saver`` (BM = models/base_model.py: creation BM:62, ``saver.save`` at
SBM:330-333 / SBM:214, ``saver.restore`` in ``load_model`` BM:401-417); what lands on disk is the variable list that existed
when the Saver was built - model variables and BN moving statistics, no Adam slots.  Two formats are implemented here:

``safetensors`` (default for new checkpoints)
    one ``<path>.safetensors`` file: 8-byte little-endian header length, a JSON header ``{name: {dtype, shape, data_offsets}}``
    (+ ``__metadata__``), then the raw little-endian tensor bytes.  Written and read by this module (no dependency); the tests
    cross-check it with the ``safetensors`` package when that is installed.  Optimizer state, which the reference never saves,
    goes under the separate ``optimizer/`` key-space (``optimizer/<var>/Adam``, ``optimizer/<var>/Adam_1``, ``optimizer/step``).

``tf`` (TensorFlow tensor-bundle, "checkpoint V2": ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``)
    what the reference's Saver itself writes, so that weights can move between a real TF run of the reference and this
    implementation by variable NAME.  The index is a LevelDB-format sorted table (prefix-compressed blocks with restart points,
    a 5-byte trailer of compression type + masked CRC-32C per block, index block, 48-byte footer ending in the table magic);
    key "" holds ``BundleHeaderProto`` and every other key a ``BundleEntryProto`` (dtype, shape, shard, offset, size, masked
    CRC-32C of the tensor bytes).  PARITY NOTE: written from the published format (tensorflow/core/util/tensor_bundle,
    tensorflow/core/lib/io/format.cc, table_builder.cc); TensorFlow is not installable offline, so no TF-written file was
    available to test the reader against - the tests check the KATs of CRC-32C, the structure byte by byte on a small case, and
    round trips.

A legacy ``<path>.npz`` (names with ``/`` replaced by ``|``) written by earlier versions is still read.
"""
import ctypes as C
import json
import os
import struct

import numpy as np

__all__ = ["save_safetensors", "load_safetensors", "save_tf_bundle", "load_tf_bundle", "save", "load", "exists", "remove",
           "crc32c", "OPT"]

OPT = "optimizer/"          # key-space of the Adam slots / step (never written by the reference)

# ----------------------------------------------------------------------------- safetensors
_ST_DTYPES = {"F64": np.float64, "F32": np.float32, "F16": np.float16, "I64": np.int64, "I32": np.int32, "I16": np.int16,
              "I8": np.int8, "U8": np.uint8, "BOOL": np.bool_, "U16": np.uint16, "U32": np.uint32, "U64": np.uint64}
_ST_NAMES = {np.dtype(v): k for k, v in _ST_DTYPES.items()}


def _raw(a):
    """The bytes of a C-contiguous array without a copy where the buffer protocol allows it."""
    return a.tobytes() if a.ndim == 0 or a.size == 0 else memoryview(a).cast("B")


def save_safetensors(path, tensors, metadata=None):
    """tensors: name -> array (written in sorted name order, C-contiguous, little-endian)."""
    header, offset, arrays = {}, 0, []
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    for name in sorted(tensors):
        a = np.asarray(tensors[name])                       # ascontiguousarray would turn a 0-d array into 1-d
        a = np.ascontiguousarray(a).reshape(a.shape)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        if a.dtype not in _ST_NAMES:
            raise TypeError(f"{name}: dtype {a.dtype} has no safetensors name")
        header[name] = {"dtype": _ST_NAMES[a.dtype], "shape": list(a.shape), "data_offsets": [offset, offset + a.nbytes]}
        offset += a.nbytes
        arrays.append(a)
    blob = json.dumps(header, separators=(",", ":")).encode("utf-8")
    blob += b" " * (-len(blob) % 8)                         # data starts 8-byte aligned (the format allows trailing spaces)
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(struct.pack("<Q", len(blob)))
        f.write(blob)
        for a in arrays:
            f.write(_raw(a))
    os.replace(tmp, path)                                   # a reader never sees a half-written checkpoint
    return path


def load_safetensors(path):
    """-> (name -> array, metadata dict).  Validates the header against the file size like the format asks."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        head = f.read(8)
        if len(head) != 8:
            raise ValueError(f"{path}: not a safetensors file (shorter than its length field)")
        (n,) = struct.unpack("<Q", head)
        if n > size - 8 or n > 100 * 1024 * 1024:
            raise ValueError(f"{path}: header length {n} exceeds the file")
        header = json.loads(f.read(n).decode("utf-8"))
        meta = header.pop("__metadata__", {}) or {}
        data_start, out, spans = 8 + n, {}, []
        for name, d in header.items():
            dt = np.dtype(_ST_DTYPES[d["dtype"]])
            a, b = d["data_offsets"]
            count = int(np.prod(d["shape"], dtype=np.int64)) if d["shape"] else 1
            if not (0 <= a <= b <= size - data_start) or b - a != count * dt.itemsize:
                raise ValueError(f"{path}: tensor {name} has inconsistent offsets")
            spans.append((a, b))
            f.seek(data_start + a)
            out[name] = np.frombuffer(f.read(b - a), dtype=dt).reshape(d["shape"]).copy()
        spans.sort()
        if spans and (spans[0][0] != 0 or any(x[1] != y[0] for x, y in zip(spans, spans[1:])) or spans[-1][1] != size - data_start):
            raise ValueError(f"{path}: tensor data does not tile the file")
    return out, meta


# ----------------------------------------------------------------------------- CRC-32C
def crc32c(data, crc=0):
    """CRC-32C of a bytes-like / contiguous array, continuing from `crc` (native: csrc/crc32c.cu)."""
    from . import _lib
    a = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).reshape(-1).view(np.uint8)
    return int(_lib.load().pamrec_crc32c(C.c_uint32(crc), a.ctypes.data if a.size else None, a.size))


def _mask(crc):
    """tensorflow/core/lib/hash/crc32c.h: rotate right by 15, add a constant (CRCs of data that embeds CRCs)."""
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xffffffff


def _unmask(m):
    rot = (m - 0xa282ead8) & 0xffffffff
    return ((rot >> 17) | (rot << 15)) & 0xffffffff


# ----------------------------------------------------------------------------- protobuf wire format (the few fields used)
def _varint(n):
    n &= (1 << 64) - 1                                      # negative int64 -> 10-byte two's complement, as protobuf does
    out = bytearray()
    while True:
        b = n & 0x7f
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _read_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7f) << shift
        if not b & 0x80:
            return val, pos
        shift += 7
        if shift > 70:
            raise ValueError("varint too long")


def _fields(buf):
    """Yield (field number, wire type, value) of one message; value is int (varint / fixed) or bytes (length-delimited)."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _read_varint(buf, pos)
        elif wt == 1:
            val, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            val, pos = bytes(buf[pos:pos + ln]), pos + ln
        elif wt == 5:
            val, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, val


# tensorflow/core/framework/types.proto
_TF_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
              17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_TF_ENUM = {np.dtype(v): k for k, v in _TF_DTYPES.items()}


def _entry_proto(a, offset, crc):
    """BundleEntryProto (tensor_bundle.proto): dtype=1, shape=2 {dim=2 {size=1}}, shard_id=3, offset=4, size=5, crc32c=6 (fixed32).
    proto3: zero-valued scalars are not written."""
    shape = b"".join(b"\x12" + _varint(len(d)) + d for d in ((b"\x08" + _varint(s) if s else b"") for s in a.shape))
    out = b"\x08" + _varint(_TF_ENUM[a.dtype]) + b"\x12" + _varint(len(shape)) + shape
    if offset:
        out += b"\x20" + _varint(offset)
    if a.nbytes:
        out += b"\x28" + _varint(a.nbytes)
    return out + b"\x35" + struct.pack("<I", crc)


def _parse_entry(buf):
    e = dict(dtype=0, shape=[], shard=0, offset=0, size=0, crc=0, slices=0, unknown_rank=False)
    for num, _, val in _fields(buf):
        if num == 1:
            e["dtype"] = val
        elif num == 2:
            for n2, _, v2 in _fields(val):
                if n2 == 2:
                    e["shape"].append(next((v3 for n3, _, v3 in _fields(v2) if n3 == 1), 0))
                elif n2 == 3:
                    e["unknown_rank"] = bool(v2)
        elif num == 3:
            e["shard"] = val
        elif num == 4:
            e["offset"] = val
        elif num == 5:
            e["size"] = val
        elif num == 6:
            e["crc"] = val
        elif num == 7:
            e["slices"] += 1
    return e


# ----------------------------------------------------------------------------- LevelDB-format table (the .index file)
_MAGIC = 0xdb4775248b80fb57
_RESTART_INTERVAL = 16
_BLOCK_SIZE = 256 * 1024      # the default of TF's table::Options; any size reads back


class _BlockBuilder:
    def __init__(self):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""

    def add(self, key, value):
        shared = 0
        if self.count % _RESTART_INTERVAL == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4


def _write_block(f, contents):
    """contents + 1-byte compression type (0: none) + masked CRC-32C of both; returns the BlockHandle (offset, size)."""
    off = f.tell()
    f.write(contents)
    f.write(b"\x00" + struct.pack("<I", _mask(crc32c(contents + b"\x00"))))
    return off, len(contents)


def _write_table(path, items):
    """items: sorted [(key bytes, value bytes)]."""
    with open(path, "wb") as f:
        index, blk = _BlockBuilder(), _BlockBuilder()
        pending = None

        def flush():
            nonlocal blk, pending
            if blk.count:
                handle = _write_block(f, blk.finish())
                pending = (blk.last, handle)                # index key: any key >= the block's last and < the next block's first
                blk = _BlockBuilder()
        for key, value in items:
            if pending:
                index.add(pending[0], _varint(pending[1][0]) + _varint(pending[1][1]))
                pending = None
            blk.add(key, value)
            if blk.size() >= _BLOCK_SIZE:
                flush()
        flush()
        if pending:
            index.add(pending[0], _varint(pending[1][0]) + _varint(pending[1][1]))
        meta = _write_block(f, _BlockBuilder().finish())    # empty metaindex block (no filter policy)
        idx = _write_block(f, index.finish())
        foot = _varint(meta[0]) + _varint(meta[1]) + _varint(idx[0]) + _varint(idx[1])
        f.write(foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", _MAGIC))


def _read_block(buf, offset, size, path):
    contents, kind = buf[offset:offset + size], buf[offset + size]
    (stored,) = struct.unpack_from("<I", buf, offset + size + 1)
    if _unmask(stored) != crc32c(bytes(buf[offset:offset + size + 1])):
        raise ValueError(f"{path}: block at {offset} fails its checksum")
    if kind != 0:
        raise ValueError(f"{path}: compressed table block (type {kind}); tensor-bundle indexes are written uncompressed")
    (n_restarts,) = struct.unpack_from("<I", contents, len(contents) - 4)
    end = len(contents) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _read_varint(contents, pos)
        unshared, pos = _read_varint(contents, pos)
        vlen, pos = _read_varint(contents, pos)
        key = key[:shared] + bytes(contents[pos:pos + unshared])
        pos += unshared
        out.append((key, bytes(contents[pos:pos + vlen])))
        pos += vlen
    return out


def _read_table(path):
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != _MAGIC:
        raise ValueError(f"{path}: not a table file (bad magic)")
    foot = buf[-48:-8]
    pos = 0
    _, pos = _read_varint(foot, pos)
    _, pos = _read_varint(foot, pos)
    idx_off, pos = _read_varint(foot, pos)
    idx_size, pos = _read_varint(foot, pos)
    out = []
    for _, handle in _read_block(buf, idx_off, idx_size, path):
        off, p = _read_varint(handle, 0)
        size, _ = _read_varint(handle, p)
        out.extend(_read_block(buf, off, size, path))
    return out


# ----------------------------------------------------------------------------- tensor bundle
def _data_path(prefix, shard, n):
    return "{}.data-{:05d}-of-{:05d}".format(prefix, shard, n)


def save_tf_bundle(prefix, tensors):
    """Write ``prefix.index`` + ``prefix.data-00000-of-00001`` holding `tensors` (name -> array) like BundleWriter does for a
    tf.train.Saver: one shard, tensors in name order, little-endian, every entry with its masked CRC-32C."""
    names = sorted(tensors, key=lambda s: s.encode("utf-8"))
    if "" in tensors:
        raise ValueError('the empty name is the bundle header\'s key')
    items = [(b"", b"\x08\x01" + b"\x1a\x02\x08\x01")]      # BundleHeaderProto: num_shards = 1, (LITTLE = 0), version.producer = 1
    offset = 0
    tmp = _data_path(prefix, 0, 1) + ".tmp"
    with open(tmp, "wb") as f:
        for name in names:
            a = np.asarray(tensors[name])
            a = np.ascontiguousarray(a).reshape(a.shape)
            if a.dtype.byteorder == ">":
                a = a.astype(a.dtype.newbyteorder("<"))
            if a.dtype not in _TF_ENUM:
                raise TypeError(f"{name}: dtype {a.dtype} has no TensorFlow DataType here")
            f.write(_raw(a))
            items.append((name.encode("utf-8"), _entry_proto(a, offset, _mask(crc32c(a)))))
            offset += a.nbytes
    os.replace(tmp, _data_path(prefix, 0, 1))
    _write_table(prefix + ".index.tmp", items)
    os.replace(prefix + ".index.tmp", prefix + ".index")
    return prefix


def load_tf_bundle(prefix, verify=True):
    """name -> array for every tensor of a tensor-bundle checkpoint (TF "V2" format).  Sliced (partitioned) variables and
    big-endian bundles are refused; string / resource / variant tensors are skipped with their names in ``.skipped``."""
    rows = _read_table(prefix + ".index")
    if not rows or rows[0][0] != b"":
        raise ValueError(f"{prefix}.index: no bundle header")
    num_shards, endian = 0, 0
    for num, _, val in _fields(rows[0][1]):
        if num == 1:
            num_shards = val
        elif num == 2:
            endian = val
    if endian != 0:
        raise ValueError(f"{prefix}: big-endian bundle")
    files, out = {}, _Loaded()
    try:
        for key, val in rows[1:]:
            name, e = key.decode("utf-8"), _parse_entry(val)
            if e["slices"]:
                raise ValueError(f"{prefix}: {name} is a partitioned variable (slices are not supported)")
            if e["dtype"] not in _TF_DTYPES:
                out.skipped.append(name)                     # e.g. the object-graph proto a TF2 checkpoint carries (DT_STRING)
                continue
            dt = np.dtype(_TF_DTYPES[e["dtype"]])
            count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
            if e["unknown_rank"] or count * dt.itemsize != e["size"]:
                raise ValueError(f"{prefix}: {name} has shape {e['shape']} but {e['size']} bytes")
            if e["shard"] not in files:
                files[e["shard"]] = open(_data_path(prefix, e["shard"], max(num_shards, 1)), "rb")
            f = files[e["shard"]]
            f.seek(e["offset"])
            raw = f.read(e["size"])
            if len(raw) != e["size"]:
                raise ValueError(f"{prefix}: {name} runs past the end of its data file")
            if verify and _unmask(e["crc"]) != crc32c(raw):
                raise ValueError(f"{prefix}: {name} fails its checksum")
            out[name] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    finally:
        for f in files.values():
            f.close()
    return out


class _Loaded(dict):
    def __init__(self):
        super().__init__()
        self.skipped = []


# ----------------------------------------------------------------------------- one entry point for the Saver
FORMATS = ("safetensors", "tf", "npz")


def _files(path):
    """Files that make up a checkpoint `path` in any of the formats."""
    d, base = os.path.split(path)
    out = [path + ".safetensors", path + ".npz", path + ".index"]
    if os.path.isdir(d or "."):
        out += [os.path.join(d, n) for n in os.listdir(d or ".") if n.startswith(base + ".data-")]
    return [p for p in out if os.path.exists(p)]


def exists(path):
    return any(p.endswith((".safetensors", ".npz", ".index")) for p in _files(path))


def remove(path, keep_fmt=None):
    """Delete the files of checkpoint `path`; with keep_fmt, only the files of the OTHER formats (called after a successful
    save so that a crash between the two never leaves `path` without a complete checkpoint)."""
    keep = {"safetensors": (".safetensors",), "npz": (".npz",), "tf": (".index", ".data-")}.get(keep_fmt, ())
    base = os.path.basename(path)
    for p in _files(path):
        tail = os.path.basename(p)[len(base):]
        if any(tail.startswith(k) for k in keep):
            continue
        os.remove(p)


def save(path, variables, fmt="safetensors", optimizer=None, metadata=None):
    """variables: TF name -> array; optimizer: optional TF slot name -> array (+ "step"), stored under ``optimizer/``."""
    if fmt not in FORMATS:
        raise ValueError(f"checkpoint_format must be one of {FORMATS}")
    tensors = dict(variables)
    for k, v in (optimizer or {}).items():
        tensors[OPT + k] = np.asarray(v, dtype=np.int64) if k == "step" else v
    if fmt == "safetensors":
        save_safetensors(path + ".safetensors", tensors, metadata={"format": "pamrec_b200", **(metadata or {})})
    elif fmt == "tf":
        save_tf_bundle(path, tensors)
    else:
        tmp = path + ".tmp.npz"
        np.savez(tmp, **{k.replace("/", "|"): v for k, v in tensors.items()})
        os.replace(tmp, path + ".npz")
    return path


def load(path):
    """-> (variables, optimizer state or None), whichever format `path` was written in."""
    if os.path.exists(path + ".safetensors"):
        tensors, _ = load_safetensors(path + ".safetensors")
    elif os.path.exists(path + ".index"):
        tensors = load_tf_bundle(path)
    elif os.path.exists(path + ".npz"):
        with np.load(path + ".npz") as z:
            tensors = {k.replace("|", "/"): z[k] for k in z.files}
    else:
        raise FileNotFoundError(path)
    variables = {k: v for k, v in tensors.items() if not k.startswith(OPT)}
    opt = {k[len(OPT):]: v for k, v in tensors.items() if k.startswith(OPT)}
    return variables, (opt or None)
