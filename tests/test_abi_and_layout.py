"""CPU-side checks: the C-ABI library loads, exports every symbol of include/pamrec_b200.h, and its
variable inventory matches the oracle's restatement of the reference graph."""
import ctypes
import os
import re

import pytest

from oracle import pamrec_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(lib_built):
    from pamrec_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pamrec_b200.h")).read()
    declared = set(re.findall(r"\b(pamrec_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(lib_built)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTS), "python binding and header disagree"
    assert b"sm_100a" in _lib.load().pamrec_version()


def test_struct_layouts_match_the_header(lib_built):
    """The ctypes mirrors have the sizes the compiler gave the header's structs (checked again at every load)."""
    import ctypes as C
    from pamrec_b200 import _lib as L
    lib = L.load()
    sizes = (C.c_int64 * 6)()
    assert lib.pamrec_abi_sizes(sizes) == 0
    assert list(sizes) == [C.sizeof(t) for t in (L.PamrecConfig, L.PamrecBatch, L.PamrecBuffers, L.PamrecTensorInfo, L.PamrecLines, L.PamrecVocab)]
    # field offsets that a binding in another language is most likely to get wrong (padding after int32 members)
    assert L.PamrecBatch.item_history.offset == 8 and L.PamrecBatch.global_batch.offset == 8 + 10 * 8
    assert L.PamrecTensorInfo.offset.offset % 8 == 0 and L.PamrecLines.offsets.offset == 8


def test_inventory_matches_oracle(lib_built):
    from pamrec_b200.engine import Engine
    eng = Engine(n_users=37, n_items=211, n_cates=13, max_seq_len=50, max_batch=25)
    shapes = eng.variable_shapes()
    spec, bn = O.param_spec(37, 211, 13, 50)
    want = {n: tuple(s) for n, s, _, _ in spec}
    for scope, c in bn:
        want[scope + "/moving_mean"] = (c,)
        want[scope + "/moving_variance"] = (c,)
    assert set(shapes) == set(want)
    for n in want:
        assert tuple(shapes[n]) == want[n], n
    # L2 / position flags follow sequential_base_model.py:714-721
    from pamrec_b200 import _lib as L
    grp = {n: g for n, _, _, g in spec}
    for n, d in eng.info[L.POOL_DENSE].items():
        assert bool(d["flags"] & L.SEG_L2) == (grp[n] == "layer"), n
        assert bool(d["flags"] & L.SEG_POS) == (grp[n] == "pos"), n
    # segments tile the dense pool without gaps
    segs = sorted((d["offset"], d["numel"]) for d in eng.info[L.POOL_DENSE].values())
    pos = 0
    for off, n in segs:
        assert off == pos
        pos += n
    assert pos == eng.dense_numel
    eng.close()


def test_create_rejects_bad_config(lib_built):
    from pamrec_b200.engine import Engine, PamrecError
    with pytest.raises(PamrecError):
        Engine(n_users=10, n_items=10, n_cates=10, max_seq_len=1000, max_batch=5)
    with pytest.raises(PamrecError):
        Engine(n_users=0, n_items=10, n_cates=10, max_seq_len=10, max_batch=5)


def test_no_cpu_fallback(lib_built):
    import torch
    from pamrec_b200.engine import Engine, PamrecError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    eng = Engine(n_users=10, n_items=10, n_cates=10, max_seq_len=10, max_batch=5)
    with pytest.raises(PamrecError):
        eng.allocate()
