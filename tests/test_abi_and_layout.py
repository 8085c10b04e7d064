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


@pytest.mark.parametrize("model", ["mmoe", "ple", "sharebottom", "sasrec"])
def test_sibling_inventory_matches_oracle(lib_built, model):
    """MMoEModel_original / PLEModel / ShareBottomModel: the library's variables, by TF name, are the oracle's param_spec."""
    from oracle import siblings_oracle as S
    from pamrec_b200 import _lib as L
    from pamrec_b200.engine import Engine, PamrecError
    eng = Engine(n_users=37, n_items=211, n_cates=13, max_seq_len=50, max_batch=25, model=model)
    shapes = eng.variable_shapes()
    spec, bn = S.sasrec_param_spec(37, 211, 13, 50) if model == "sasrec" else S.param_spec(model, 37, 211, 13)
    want = {n: tuple(s) for n, s, _, _ in spec}
    for scope, c in bn:
        want[scope + "/moving_mean"] = (c,)
        want[scope + "/moving_variance"] = (c,)
    assert set(shapes) == set(want)
    for n in want:
        assert tuple(shapes[n]) == want[n], n
    for n, d in eng.info[L.POOL_DENSE].items():                                      # everything outside sequential/embedding gets L2
        assert bool(d["flags"] & L.SEG_L2) == (not n.startswith("sequential/embedding/")), n
    segs = sorted((d["offset"], d["numel"]) for d in eng.info[L.POOL_DENSE].values())
    pos = 0
    for off, n in segs:
        assert off == pos
        pos += n
    assert pos == eng.dense_numel
    eng.close()
    with pytest.raises(PamrecError):
        Engine(n_users=37, n_items=211, n_cates=13, max_seq_len=50, max_batch=25, model=model, world_size=2, tables="replicated")
    with pytest.raises(PamrecError):
        Engine(n_users=37, n_items=211, n_cates=13, max_seq_len=50, max_batch=25, model="din")


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


def test_config_validation_and_calls_before_bind(lib_built):
    """Error behaviour of the C ABI that needs no device: return codes of pamrec_create for every rejected field, the message of
    a device call on an unbound handle, and the host-only entry points handed nonsense."""
    import ctypes as C
    import numpy as np
    from pamrec_b200 import _lib as L
    lib = L.load()
    base = dict(n_users=10, n_items=20, n_cates=5, max_seq_len=8, max_batch=10, learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8,
                max_grad_norm=2.0, is_clip_norm=1, fuzhu_weight=0.5, order_weight=0.1, world_size=1, rank=0)

    def create(**kw):
        cfg = L.PamrecConfig(**{**base, **kw})
        h = C.c_void_p()
        rc = lib.pamrec_create(C.byref(cfg), C.byref(h))
        if rc == 0:
            lib.pamrec_destroy(h)
        else:
            assert not h.value                                    # no handle on failure
        return rc
    assert create() == 0
    assert create(n_cates=0) == -2 and create(max_seq_len=257) == -3 and create(max_batch=0) == -4
    assert create(world_size=2, rank=2, table_mode=L.TABLES_SHARDED) == -5 and create(world_size=65, table_mode=L.TABLES_SHARDED) == -5
    assert create(table_mode=7) == -6 and create(world_size=2, table_mode=L.TABLES_LOCAL) == -6
    assert create(loss_kind=2) == -8
    assert create(loss_kind=L.LOSS_SOFTMAX, softmax_group=3) == 0                                   # one GPU: any group
    assert create(loss_kind=L.LOSS_SOFTMAX, softmax_group=3, world_size=2, table_mode=L.TABLES_SHARDED) == -8   # ranks hold groups of 5
    assert create(loss_kind=L.LOSS_SOFTMAX, softmax_group=5, world_size=2, table_mode=L.TABLES_SHARDED) == 0
    assert create(model_kind=5) == -9 and create(model_kind=L.MODEL_PLE) == 0 and create(model_kind=L.MODEL_SASREC) == 0
    assert create(model_kind=L.MODEL_MMOE, world_size=2, table_mode=L.TABLES_REPLICATED) == -9     # sibling models: one GPU
    assert create(model_kind=L.MODEL_MMOE, loss_kind=L.LOSS_SOFTMAX, softmax_group=5) == -9
    assert lib.pamrec_create(None, None) == -1
    # a device call before pamrec_bind: error code + message, no crash
    cfg, h = L.PamrecConfig(**base), C.c_void_p()
    assert lib.pamrec_create(C.byref(cfg), C.byref(h)) == 0
    batch = L.PamrecBatch(batch=5)
    assert lib.pamrec_train_step(h, C.byref(batch), 1, None, None) != 0
    assert b"pamrec_bind" in lib.pamrec_last_error(h)
    info = L.PamrecTensorInfo()
    assert lib.pamrec_tensor_info(h, L.POOL_DENSE, 10 ** 6, C.byref(info)) != 0
    lib.pamrec_destroy(h)
    # batcher / tokenizer misuse
    off = np.zeros(1, np.int64)
    lines = L.PamrecLines(n_lines=0, offsets=off.ctypes.data)
    borders = np.asarray([0.0, 1.0])
    bh = C.c_void_p()
    assert lib.pamrec_batcher_create(C.byref(lines), borders.ctypes.data, 2, 0, C.byref(bh)) == -1          # max_seq_len < 1
    assert lib.pamrec_batcher_create(C.byref(lines), borders.ctypes.data, 2, 8, C.byref(bh)) == 0
    ptrs = (C.c_void_p * 19)()
    assert lib.pamrec_batcher_next(bh, 10, ptrs) == -1                                                       # no pass begun
    assert lib.pamrec_batcher_begin_eval(bh, 1) == -1                                                        # a train file has no label columns
    assert lib.pamrec_batcher_begin_train(bh, None, None, 0) == 0
    assert lib.pamrec_batcher_next(bh, 7, ptrs) == -3                                                        # not a multiple of 5
    n = C.c_int(-1)
    assert lib.pamrec_batcher_next_shard(bh, 10, 2, 2, ptrs, None, None, C.byref(n)) == -1                  # rank outside the world
    assert lib.pamrec_batcher_next_shard(bh, 10, 2, 1, ptrs, None, None, C.byref(n)) == 0 and n.value == 0  # empty pass
    assert lib.pamrec_batcher_destroy(bh) == 0
    assert lib.pamrec_tokenize_file(None, 1, None, None, None, 0, None, None, None) == -1
    assert lib.pamrec_tokens_read(None, None) == -1 and lib.pamrec_tokens_free(None) == 0
