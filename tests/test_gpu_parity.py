"""GPU parity: every stage of the CUDA step, called through the C ABI, against the fp64 oracle on the
same injected weights and the same seeded batch.  Tolerance from BASELINE.json north_star:
1e-5 relative for fp32 logits and loss (tensors are compared relative to their max magnitude)."""
import math

import numpy as np
import pytest
import torch

from oracle import pamrec_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _close(name, got, ref, rtol=RTOL, report=None):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64).reshape(got.shape)
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(got - ref).max() / scale
    if report is not None:
        report.append((name, err))
    assert np.isfinite(got).all(), f"{name}: non-finite values"
    assert err <= rtol, f"{name}: max err / max|ref| = {err:.3e} > {rtol:.1e}"


def _setup(nu, ni, nc, T, B, seed, hp=None, sparse_adam="dense_exact"):
    from pamrec_b200.engine import Engine
    om = O.OracleModel(nu, ni, nc, T, hp=hp, seed=seed)
    O.perturb_params(om.params, om.bn_state, seed=seed + 1)
    eng = Engine(nu, ni, nc, T, B, hp=hp, sparse_adam=sparse_adam).allocate()
    eng.set_variables({n: t.numpy() for n, t in om.params.items()})
    eng.set_variables({n: t.numpy() for n, t in om.bn_state.items()})
    return om, eng


CASES = [
    # nu, ni, nc, T, B
    (50, 300, 20, 12, 20),
    (400, 5000, 60, 50, 130),     # B*T not a multiple of the 128-token tile; T=50 like cfg-B
    (200, 2000, 40, 100, 35),     # quick-start T
    (100, 1000, 30, 200, 10),     # long history (cfg-4 T)
    # the BASELINE.json shapes themselves (bench.py WORKLOADS): multi-wave grids, k_pos_grad grid.y > 1, > 64 K-row dense
    # tiles, the T = 200 shared-memory footprints at full occupancy
    (50000, 30000, 50, 50, 1025),      # configs[1] takatak_b1025_t50
    (20000, 100000, 500, 100, 500),    # configs[0] quick start, wechat_b500_t100
    (50000, 30000, 50, 200, 4095),     # configs[3] long_b4095_t200 (fp64 oracle: ~40 s, 16 GB on 8 host cores)
]


@pytest.mark.parametrize("nu,ni,nc,T,B", CASES)
def test_train_step_stagewise(nu, ni, nc, T, B):
    from pamrec_b200 import _lib as L
    from relu_masks import engine_relu_masks
    om, eng = _setup(nu, ni, nc, T, B, seed=11)
    om.proj = "grouped"                     # same sums as the reference's [B,T,40,40] gather without materialising it
    batch = O.make_batch(5, B, T, nu, ni, nc)
    keep = ("x0", "new_long", "blk0.out", "blk1.out", "logits")
    db = eng.upload(batch)
    # The oracle differentiates on the ENGINE's ReLU pattern (tests/relu_masks.py): gradient parity is then defined at any batch
    # size instead of only where no pre-activation happens to fall inside fp32 rounding of a kink.
    eng.set_debug(L.DEBUG_SAVE_FFN_HIDDEN)
    eng.forward(db, training=True, want_pred=False)
    masks = engine_relu_masks(eng, B)
    eng.set_debug(0)
    ref = om.train_step(batch, apply=False, keep=keep, relu_masks=masks)
    n_units = sum(int(np.prod(m.shape)) for m in masks.values())
    print(f"\n[{nu},{ni},{nc},T={T},B={B}] ReLU units {n_units}, on opposite sides in fp32 / fp64: {ref['relu_forced']}")
    assert ref["relu_forced"] <= max(4, n_units // 100000), "the two forward passes disagree on far more units than rounding explains"
    t = ref["t"]
    rep = []
    # ---- gather alone (bit-exact: pure loads + one add)
    x0 = eng.gather(db).cpu().numpy()
    x0_ref32 = (torch.cat([om.params[O_EMB + "item_embedding"][torch.as_tensor(batch["item_history"]).long()],
                           om.params[O_EMB + "cate_embedding"][torch.as_tensor(batch["item_cate_history"]).long()],
                           torch.cat([om.params[O_EMB + "item_embedding"][torch.as_tensor(batch["items"]).long()],
                                      om.params[O_EMB + "cate_embedding"][torch.as_tensor(batch["cates"]).long()]], -1)[:, None, :].expand(B, T, 20)], 2)
                + om.params[O_EMB + "position_embedding"][None]).numpy()
    assert np.array_equal(x0, x0_ref32), "gather is not bit-exact"
    # ---- forward
    eng.forward(db, training=True, want_pred=False)
    torch.cuda.synchronize()
    for k in range(2):
        for nm in ("qin", "Q", "K", "V", "y", "out"):
            _close(f"blk{k}.{nm}", eng.ws(f"blk{k}.{nm}", B).cpu().numpy(), t[f"blk{k}.{nm}"].detach().numpy(), report=rep)
    _close("z1", eng.ws("z1", B).cpu().numpy(), t["z1"].detach().numpy(), report=rep)
    _close("z2", eng.ws("z2", B).cpu().numpy(), t["z2"].detach().numpy(), report=rep)
    _close("new_long", eng.ws("new_long", B).cpu().numpy(), t["new_long"].detach().numpy(), report=rep)
    ze0 = np.concatenate([t[f"expert{j}.z0"].detach().numpy() for j in range(5)], 1)
    ze1 = np.concatenate([t[f"expert{j}.z1"].detach().numpy() for j in range(5)], 1)
    _close("ze0", eng.ws("ze0", B).cpu().numpy(), ze0, report=rep)
    _close("ze1", eng.ws("ze1", B).cpu().numpy(), ze1, report=rep)
    zg1 = np.concatenate([t["gate_main.z1"].detach().numpy(), t["gate_sub.z1"].detach().numpy()], 1)
    _close("zg1", eng.ws("zg1", B).cpu().numpy(), zg1, report=rep)
    u = eng.ws("u", B).cpu().numpy()
    _close("main", u[:, :64], t["main"].detach().numpy(), report=rep)
    _close("sub", u[:, 84:148], t["sub"].detach().numpy(), report=rep)
    zt1 = np.concatenate([t[f"tower{g}.z1"].detach().numpy() for g in range(3)], 1)
    _close("zt1", eng.ws("zt1", B).cpu().numpy(), zt1, report=rep)
    _close("logits", eng.ws("logits", B).cpu().numpy(), t["logits"].detach().numpy(), report=rep)
    # ---- backward
    P0 = eng.pool["dense_param"].clone()
    eng.backward(db)
    torch.cuda.synchronize()
    _close("d_logits", eng.ws("d_logits", B).cpu().numpy(), t["logits"].grad.numpy(), rtol=2e-5, report=rep)
    _close("d_new_long", eng.ws("d_new_long", B).cpu().numpy(), t["new_long"].grad.numpy(), rtol=5e-5, report=rep)
    _close("d_x0", eng.ws("g_a", B).cpu().numpy(), t["x0"].grad.numpy(), rtol=1e-4, report=rep)
    l2 = om.hp["layer_l2"]
    # Parameter gradients are sums over B*T tokens that can cancel almost completely (e.g. a bias in front of a
    # mean-removing BN): the fp32 rounding noise of such a sum scales with the summands, not with the result, so
    # the absolute floor is tied to the largest gradient entry of the whole model.
    gmax = max(float(ref["grads"][n].abs().max()) for n in eng.info[L.POOL_DENSE])
    bad = []
    for name, d in eng.info[L.POOL_DENSE].items():
        g = eng.dense(name, "dense_grad").cpu().numpy().astype(np.float64)
        if d["flags"] & L.SEG_L2:
            g = g + l2 * eng.dense(name).cpu().numpy().astype(np.float64)
        gr = ref["grads"][name].numpy().reshape(g.shape)
        err = np.abs(g - gr).max()
        lim = 1e-4 * np.abs(gr).max() + 2e-6 * gmax
        rep.append(("grad " + name, err / max(np.abs(gr).max(), 1e-30)))
        if not (np.isfinite(g).all() and err <= lim):
            bad.append((name, err, np.abs(gr).max()))
    assert not bad, f"gmax={gmax:.3e} " + "; ".join(f"{n}: err {e:.3e} max|ref| {m:.3e}" for n, e, m in bad)
    # ---- apply: sparse gradients, clip norms, losses
    tables0 = {k: eng.pool[k].clone() for k in ("item_w", "cate_w", "ulong_w", "ushort_w")}
    M0, V0 = eng.pool["dense_m"].clone(), eng.pool["dense_v"].clone()
    G0 = eng.pool["dense_grad"].clone()
    losses = eng.apply_gradients(db).cpu().numpy()
    torch.cuda.synchronize()
    lr = ref["losses"]
    for i, k in enumerate(("loss", "data_loss", "regular_loss", "auxiliary_data_loss", "order_loss")):
        assert abs(losses[i] - lr[k]) <= 1e-5 * max(abs(lr[k]), 1e-3), (k, losses[i], lr[k])
    nun = eng.ws("sp.nuniq").cpu().numpy()
    el2 = om.hp["embed_l2"]
    for tab, idx, name in (("item", 0, O_EMB + "item_embedding"), ("cate", 1, O_EMB + "cate_embedding")):
        n = int(nun[idx])
        uk = eng.ws(f"sp.{tab}.ukeys").cpu().numpy()[:n]
        acc = eng.ws(f"sp.{tab}.accum").cpu().numpy()[:n].astype(np.float64)
        gref = ref["grads"][name].numpy()
        touched = np.unique(np.concatenate([batch["item_history" if tab == "item" else "item_cate_history"].reshape(-1),
                                            batch["items" if tab == "item" else "cates"]]))
        assert np.array_equal(uk, touched), "unique ids differ"        # sortedness + uniqueness, bit exact
        g = acc + el2 * tables0[tab + "_w"].cpu().numpy()[uk]
        _close(f"sparse grad {tab}", g, gref[uk], rtol=1e-4, report=rep)
        mask = np.ones(gref.shape[0], bool); mask[uk] = False
        assert not gref[mask].any()
    spn = eng.ws("sp_normsq").cpu().numpy()
    for i, name in enumerate(("item_embedding", "cate_embedding", "user_long_embedding", "user_short_embedding", "position_embedding")):
        want = ref["sqnorms"][O_EMB + name]
        assert abs(spn[i] - want) <= 2e-4 * want + 1e-30, (name, spn[i], want)
    segn = eng.ws("seg_normsq").cpu().numpy()
    for s, (name, d) in enumerate(eng.info[L.POOL_DENSE].items()):
        want = ref["sqnorms"][name]
        assert abs(segn[s] - want) <= 2e-4 * want + d["numel"] * (2e-6 * gmax) ** 2, (name, segn[s], want)
    # ---- optimiser kernels against the TF formulas applied to the engine's own gradients
    hp = om.hp
    b1, b2, eps = hp["beta1"], hp["beta2"], hp["epsilon"]
    f32 = lambda x: float(np.float32(x))
    lr_t = np.float32(f32(hp["learning_rate"]) * math.sqrt(1 - f32(b2)) / (1 - f32(b1)))
    for s, (name, d) in enumerate(eng.info[L.POOL_DENSE].items()):
        o, n = d["offset"], d["numel"]
        p0 = P0[o:o + n].cpu().numpy(); g = G0[o:o + n].cpu().numpy()
        if d["flags"] & L.SEG_L2:
            g = g + np.float32(l2) * p0
        norm = np.float32(math.sqrt(segn[s]))
        g = g * np.float32(2.0) / max(norm, np.float32(2.0))
        m = (M0[o:o + n].cpu().numpy() + (g - M0[o:o + n].cpu().numpy()) * (np.float32(1) - np.float32(b1))).astype(np.float32)
        v = (V0[o:o + n].cpu().numpy() + (g * g - V0[o:o + n].cpu().numpy()) * (np.float32(1) - np.float32(b2))).astype(np.float32)
        p = p0 - lr_t * m / (np.sqrt(v) + np.float32(eps))
        got = eng.dense(name).cpu().numpy().reshape(-1)
        assert np.allclose(got, p, rtol=1e-6, atol=1e-9), name
    # untouched table rows still decay/move (dense_exact); with zero slots they must stay bit-identical
    for tab in ("item", "cate", "ulong", "ushort"):
        w0 = tables0[tab + "_w"].cpu().numpy(); w1 = eng.pool[tab + "_w"].cpu().numpy()
        assert np.isfinite(w1).all()
        assert (w0 != w1).any()
    print(f"\n[{nu},{ni},{nc},T={T},B={B}] worst relative errors:")
    print("\n".join(f"  {n:70s} {e:.2e}" for n, e in sorted(rep, key=lambda x: -x[1])[:10]))
    print("  forward / activation-gradient stages:")
    print("\n".join(f"  {n:70s} {e:.2e}" for n, e in sorted((x for x in rep if not x[0].startswith("grad ")), key=lambda x: -x[1])[:8]))
    eng.close()


O_EMB = "sequential/embedding/"


def test_multi_step_losses_and_eval():
    """Five optimisation steps then a scoring pass: losses and predictions track the fp64 oracle."""
    nu, ni, nc, T, B = 300, 3000, 50, 50, 100
    om, eng = _setup(nu, ni, nc, T, B, seed=3)
    for step in range(5):
        batch = O.make_batch(100 + step, B, T, nu, ni, nc)
        ref = om.train_step(batch)
        got = eng.train_step(eng.upload(batch)).cpu().numpy()
        for i, k in enumerate(("loss", "data_loss", "regular_loss", "auxiliary_data_loss", "order_loss")):
            r = ref["losses"][k]
            assert abs(got[i] - r) <= 5e-5 * max(abs(r), 1e-3), (step, k, got[i], r)   # trajectories drift: per-step parity is the stagewise test
    ev = O.make_batch(999, 77, T, nu, ni, nc, grouped=False)
    pred = eng.forward(eng.upload(ev, training=False), training=False).cpu().numpy()
    want = om.eval_forward(ev).t["pred"].numpy().reshape(-1)
    assert np.abs(pred - want).max() <= 1e-4
    # A bias in front of a BN layer has a mathematically zero data gradient, so Adam moves it along fp32 rounding
    # noise; moving_mean tracks that bias (and cancels it again at inference), hence the looser absolute floor.
    for name in eng.info[1]:
        got = eng.bn(name).cpu().numpy()
        atol = 2e-4 if name.endswith("moving_mean") else 1e-6
        assert np.allclose(got, om.bn_state[name].numpy(), rtol=1e-4, atol=atol), name
    eng.close()


@pytest.mark.parametrize("group", [5, 1, 2])
def test_softmax_loss_branch(group):
    """hparams.loss == "softmax" (BM:222-242, PAM:97-105) with groups of train_num_ngs + 1 rows: the two softmax terms, their
    gradient at the logits, every dense gradient and three steps of losses against the oracle."""
    nu, ni, nc, T, B = 120, 900, 25, 30, 40
    hp = dict(loss="softmax", softmax_group=group)
    om, eng = _setup(nu, ni, nc, T, B, seed=21, hp=hp)
    batch = O.make_batch(9, B, T, nu, ni, nc)
    if group == 5:                                                   # the in-batch sampler's layout: one positive, four negatives
        for k in ("labels_satisfied", "labels_play"):
            batch[k] = np.tile(np.asarray([1, 0, 0, 0, 0], np.float32), B // 5).reshape(batch[k].shape)
    ref = om.train_step(batch, apply=False, keep=("logits",))
    db = eng.upload(batch)
    eng.forward(db, training=True, want_pred=False)
    eng.backward(db)
    torch.cuda.synchronize()
    _close("d_logits", eng.ws("d_logits", B).cpu().numpy(), ref["t"]["logits"].grad.numpy(), rtol=2e-5)
    from pamrec_b200 import _lib as L
    gmax = max(float(ref["grads"][n].abs().max()) for n in eng.info[L.POOL_DENSE])
    for name, d in eng.info[L.POOL_DENSE].items():
        g = eng.dense(name, "dense_grad").cpu().numpy().astype(np.float64)
        if d["flags"] & L.SEG_L2:
            g = g + om.hp["layer_l2"] * eng.dense(name).cpu().numpy().astype(np.float64)
        gr = ref["grads"][name].numpy().reshape(g.shape)
        assert np.abs(g - gr).max() <= 1e-4 * np.abs(gr).max() + 2e-6 * gmax, name
    eng2 = _setup(nu, ni, nc, T, B, seed=21, hp=hp)[1]
    for step in range(3):
        b = O.make_batch(30 + step, B, T, nu, ni, nc)
        want = om.train_step(b)["losses"]
        got = eng2.train_step(eng2.upload(b)).cpu().numpy()
        for i, k in enumerate(("loss", "data_loss", "regular_loss", "auxiliary_data_loss", "order_loss")):
            assert abs(got[i] - want[k]) <= 1e-4 * max(abs(want[k]), 1e-3), (step, k, got[i], want[k])
    # a batch that is not a multiple of the group is refused like the reference's reshape would
    if group == 2:
        from pamrec_b200.engine import PamrecError
        odd = O.make_batch(1, 15, T, nu, ni, nc)
        with pytest.raises(PamrecError, match="softmax group"):
            eng2.train_step(eng2.upload(odd))
    eng.close(); eng2.close()


def test_lazy_mode_touches_only_looked_up_rows():
    nu, ni, nc, T, B = 100, 4000, 30, 20, 25
    om, eng = _setup(nu, ni, nc, T, B, seed=5, sparse_adam="lazy")
    batch = O.make_batch(1, B, T, nu, ni, nc)
    w0 = eng.pool["item_w"].clone()
    eng.train_step(eng.upload(batch))
    torch.cuda.synchronize()
    changed = (eng.pool["item_w"] != w0).any(1).cpu().numpy()
    touched = np.zeros(ni, bool)
    touched[np.unique(np.concatenate([batch["item_history"].reshape(-1), batch["items"]]))] = True
    assert np.array_equal(changed, touched)
    eng.close()


@pytest.mark.parametrize("shape", [(100, 4000, 30, 20, 25), (300, 900, 7, 50, 400)])
def test_lazy_fused_adam_equals_dense_exact_on_the_first_step(shape):
    """LAZY applies Adam inside the run walk (kernels_sparse2.cu: complete runs directly, runs that cross a warp block through the
    compact accumulator + k_sp2_lazy_finish).  From zero Adam slots one step of DENSE_EXACT leaves untouched rows alone, so after
    the first step both modes must agree on every row of every table and slot (the second shape has few items and long histories:
    hot rows whose runs span many warp blocks, and 7 categories with runs of thousands of positions)."""
    nu, ni, nc, T, B = shape
    om, lazy = _setup(nu, ni, nc, T, B, seed=5, sparse_adam="lazy")
    om2, exact = _setup(nu, ni, nc, T, B, seed=5, sparse_adam="dense_exact")
    batch = O.make_batch(1, B, T, nu, ni, nc)
    la = lazy.train_step(lazy.upload(batch)).cpu().numpy()
    ex = exact.train_step(exact.upload(batch)).cpu().numpy()
    assert np.allclose(la, ex, rtol=1e-6, atol=1e-8), (la, ex)
    for tab in ("item", "cate", "ulong", "ushort"):
        for sfx in ("w", "m", "v"):
            a, b = lazy.pool[f"{tab}_{sfx}"].cpu().numpy(), exact.pool[f"{tab}_{sfx}"].cpu().numpy()
            # summation order of a run that crosses warp blocks is not fixed (atomics): agreement to fp32 rounding of the sum
            assert np.allclose(a, b, rtol=2e-5, atol=1e-9), (tab, sfx, float(np.abs(a - b).max()))
    # second step: rows looked up in both steps keep agreeing only where the first step touched them too; check the lazy rule
    batch2 = O.make_batch(2, B, T, nu, ni, nc)
    m0 = lazy.pool["item_m"].clone()
    lazy.train_step(lazy.upload(batch2))
    torch.cuda.synchronize()
    touched = np.zeros(ni, bool)
    touched[np.unique(np.concatenate([batch2["item_history"].reshape(-1), batch2["items"]]))] = True
    same = (lazy.pool["item_m"] == m0).all(1).cpu().numpy()
    assert same[~touched].all(), "LAZY moved the first moment of a row that was not looked up"
    lazy.close(); exact.close()


def test_error_paths():
    from pamrec_b200.engine import PamrecError
    nu, ni, nc, T, B = 50, 300, 20, 12, 20
    om, eng = _setup(nu, ni, nc, T, B, seed=1)
    bad = O.make_batch(1, 10, T, nu, ni, nc)
    for k in list(bad):
        bad[k] = bad[k][:7]
    with pytest.raises(PamrecError):          # 7 rows: not a multiple of the listwise group
        eng.train_step(eng.upload(bad))
    big = O.make_batch(1, 40, T, nu, ni, nc)
    with pytest.raises(PamrecError):          # larger than max_batch
        eng.forward(eng.upload(big), training=False)
    eng.close()


@pytest.mark.parametrize("T,B", [(12, 20), (50, 130), (100, 35), (200, 10), (256, 5)])
def test_attention_mma_and_ffma_paths_agree(T, B, monkeypatch):
    """The default attention kernels (warp-level 3xTF32 MMAs over the live keys, kernels_attn_mma.cu) against the FFMA kernels
    (PAMREC_ATTN=ffma, kernels_encoder.cu) on the same batch - including a sample without any live key (uniform weights) and
    arbitrary, non-prefix masks: outputs, saved row statistics and all three gradients."""
    nu, ni, nc = 200, 2000, 40
    batch = O.make_batch(21, B, T, nu, ni, nc)
    rng = np.random.default_rng(3)
    batch["mask"][5:10] = 0                                             # one listwise group without history
    batch["mask"][10:15] = (rng.random((1, T)) < 0.5).astype(np.int32)  # holes in the middle of a history
    got = {}
    for mode in ("ffma", "mma"):
        monkeypatch.setenv("PAMREC_ATTN", mode)
        om, eng = _setup(nu, ni, nc, T, B, seed=11)
        db = eng.upload(batch)
        eng.forward(db, training=True, want_pred=False)
        eng.backward(db)
        torch.cuda.synchronize()
        got[mode] = {k: eng.ws(k, B).cpu().numpy().copy() for k in ("blk0.y", "blk1.y", "blk0.ml", "blk1.ml", "d_Q", "d_K", "d_V", "g_a")}
        got[mode]["logits"] = eng.ws("logits", B).cpu().numpy().copy()
        eng.close()
    for k in got["mma"]:
        a, b = got["ffma"][k].astype(np.float64), got["mma"][k].astype(np.float64)
        if k.endswith(".ml"):
            # row maximum and row sum: the sum is relative to the maximum, which the two kernels round differently
            assert np.allclose(a[..., 0], b[..., 0], rtol=1e-5, atol=1e-5), k
            assert np.allclose(a[..., 1], b[..., 1], rtol=1e-4), k
            continue
        scale = max(np.abs(a).max(), 1e-30)
        assert np.abs(a - b).max() <= 1e-5 * scale, (k, float(np.abs(a - b).max() / scale))
