"""GPU parity of the sibling multi-task baselines (MMoEModel_original, PLEModel, ShareBottomModel)
against oracle/siblings_oracle.py, through the C ABI: gather bit-exact, every forward stage, activation gradients, all dense
gradients, merged sparse gradients, clip norms, the four losses, several optimisation steps and inference predictions.
The fp64 oracle differentiates on the engine's ReLU pattern (see tests/relu_masks.py for why)."""
import numpy as np
import pytest
import torch

from oracle import pamrec_oracle as O
from oracle import siblings_oracle as S

pytestmark = pytest.mark.gpu

EMB = "sequential/embedding/"
CLSR = "sequential/clsr/"
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


def expert_scopes(model):
    if model == "mmoe":
        return [f"{CLSR}expert_{j}" for j in range(5)]
    if model == "ple":
        return [f"{CLSR}share_expert_{j}" for j in range(3)] + [f"{CLSR}main_expert_{j}" for j in range(2)] + [f"{CLSR}sub_expert_{j}" for j in range(2)]
    return []


def _setup(model, nu, ni, nc, T, B, seed):
    from pamrec_b200.engine import Engine
    om = S.SiblingOracleModel(model, nu, ni, nc, T, seed=seed)
    O.perturb_params(om.params, om.bn_state, seed=seed + 1)
    eng = Engine(nu, ni, nc, T, B, model=model).allocate()
    eng.set_variables({n: t.numpy() for n, t in om.params.items()})
    eng.set_variables({n: t.numpy() for n, t in om.bn_state.items()})
    return om, eng


def _batch(seed, B, T, nu, ni, nc, min_len=1, grouped=True):
    return S.add_satisfied_fields(O.make_batch(seed, B, T, nu, ni, nc, min_len=min_len, grouped=grouped), seed=seed + 1)


def _on(z, stat, gamma, beta):
    xh = (z.astype(np.float32) - stat[:, 0].astype(np.float32)) * stat[:, 1].astype(np.float32)
    return gamma.astype(np.float64) * xh.astype(np.float64) + beta.astype(np.float64) > 0


def sibling_relu_masks(eng, model, B, T):
    ex = expert_scopes(model)
    att = [f"{CLSR}{b}/attention_fcn/att_fcn" for b in ("long_term", "short_term")]
    towers = ["sequential/logit_fcn", "sequential/valid_logit_fcn"]
    gates = [CLSR + "gate_main", CLSR + "gate_sub"]
    table = [("z1", "s0", 80, [s + "/nn_part/batch_normalization" for s in att], B * T, (B, T)),
             ("z2", "s1", 40, [s + "/nn_part/batch_normalization_1" for s in att], B * T, (B, T)),
             ("zt0", "t0", 100, [s + "/nn_part/batch_normalization" for s in towers], B, (B,)),
             ("zt1", "t1", 64, [s + "/nn_part/batch_normalization_1" for s in towers], B, (B,))]
    if ex:
        table += [("ze0", "e0", 100, [s + "/nn_part/batch_normalization" for s in ex], B, (B,)),
                  ("ze1", "e1", 64, [s + "/nn_part/batch_normalization_1" for s in ex], B, (B,)),
                  ("zg0", "g0", 64, [s + "/nn_part/batch_normalization" for s in gates], B, (B,)),
                  ("zg1", "g1", 5, [s + "/nn_part/batch_normalization_1" for s in gates], B, (B,))]
    out = {}
    for zbuf, bn, width, scopes, rows, lead in table:
        z = eng.ws(zbuf, rows).cpu().numpy().reshape(rows, len(scopes) * width)
        st = eng.ws(f"bn.{bn}.stat").cpu().numpy()
        for m, scope in enumerate(scopes):
            sl = slice(m * width, (m + 1) * width)
            g, b = eng.dense(scope + "/gamma").cpu().numpy(), eng.dense(scope + "/beta").cpu().numpy()
            out[scope] = _on(z[:, sl], st[sl], g, b).reshape(lead + (width,))
    return out


def _branches(eng, name, N, width):
    """[2, N, width] view of a branch-major workspace tensor (branch stride = the batch's own token count)."""
    return eng.ws(name).reshape(-1)[:2 * N * width].view(2, N, width).cpu().numpy()


CASES = [
    # model, nu, ni, nc, T, B, min_len
    ("mmoe", 50, 300, 20, 12, 20, 1),
    ("ple", 50, 300, 20, 12, 20, 1),
    ("sharebottom", 50, 300, 20, 12, 20, 1),
    ("mmoe", 400, 5000, 60, 50, 130, 50),          # full-length histories: id 0 only through the satisfied-only padding (no L2 on row 0)
    ("ple", 200, 2000, 40, 100, 35, 1),            # quick-start T
    ("sharebottom", 100, 1000, 30, 200, 10, 1),    # T = 200
    ("mmoe", 50000, 30000, 50, 50, 1025, 1),       # the takatak bench shape
    ("ple", 20000, 100000, 500, 100, 500, 1),      # the quick-start shape
]


@pytest.mark.parametrize("model,nu,ni,nc,T,B,min_len", CASES)
def test_sibling_step_stagewise(model, nu, ni, nc, T, B, min_len):
    from pamrec_b200 import _lib as L
    om, eng = _setup(model, nu, ni, nc, T, B, seed=11)
    batch = _batch(5, B, T, nu, ni, nc, min_len=min_len)
    N = B * T
    db = eng.upload(batch)
    eng.forward(db, training=True, want_pred=False)
    torch.cuda.synchronize()
    masks = sibling_relu_masks(eng, model, B, T)
    keep = ("x", "logits")
    ref = om.train_step(batch, apply=False, keep=keep, relu_masks=masks)
    n_units = sum(int(np.prod(m.shape)) for m in masks.values())
    print(f"\n[{model},{nu},{ni},{nc},T={T},B={B}] ReLU units {n_units}, on opposite sides in fp32 / fp64: {ref['relu_forced']}")
    assert ref["relu_forced"] <= max(4, n_units // 100000)
    t = ref["t"]
    rep = []
    # ---- gather: bit exact
    item_w, cate_w = om.params[EMB + "item_embedding"].numpy(), om.params[EMB + "cate_embedding"].numpy()
    h = _branches(eng, "sib.h", N, 20)
    want = [np.concatenate([item_w[batch["satisfied_item_history"].reshape(-1)], cate_w[batch["satisfied_cate_history"].reshape(-1)]], 1),
            np.concatenate([item_w[batch["item_history"].reshape(-1)], cate_w[batch["item_cate_history"].reshape(-1)]], 1)]
    assert np.array_equal(h[0], want[0]) and np.array_equal(h[1], want[1]), "gather is not bit-exact"
    tg = np.concatenate([item_w[batch["items"]], cate_w[batch["cates"]]], 1)
    assert np.array_equal(eng.ws("tgt", B).cpu().numpy(), tg)
    # ---- forward stages
    feat = eng.ws("sib.feat", N).cpu().numpy()
    z1, z2 = eng.ws("z1", N).cpu().numpy(), eng.ws("z2", N).cpu().numpy()
    score = eng.ws("sib.score", N).cpu().numpy()
    aw = eng.ws("sib.aw").reshape(-1)[:2 * N].view(2, N).cpu().numpy()
    for r in range(2):
        _close(f"att{r}.feat", feat[:, 80 * r:80 * r + 80], t[f"att{r}.feat"].detach().numpy(), report=rep)
        _close(f"att{r}.z0", z1[:, 80 * r:80 * r + 80], t[f"att{r}.z0"].detach().numpy(), report=rep)
        _close(f"att{r}.z1", z2[:, 40 * r:40 * r + 40], t[f"att{r}.z1"].detach().numpy(), report=rep)
        _close(f"att{r}.score", score[:, r], t[f"att{r}.score"].detach().numpy(), report=rep)
        _close(f"att{r}.w", aw[r], t[f"att{r}.w"].detach().numpy(), report=rep)
    _close("x", eng.ws("x", B).cpu().numpy(), t["x"].detach().numpy(), report=rep)
    ex = expert_scopes(model)
    if ex:
        nE = len(ex)
        _close("ze0", eng.ws("ze0", B).cpu().numpy(), np.concatenate([t[f"expert{j}.z0"].detach().numpy() for j in range(nE)], 1), report=rep)
        _close("ze1", eng.ws("ze1", B).cpu().numpy(), np.concatenate([t[f"expert{j}.z1"].detach().numpy() for j in range(nE)], 1), report=rep)
        _close("zg1", eng.ws("zg1", B).cpu().numpy(), np.concatenate([t[f"gate{k}.z1"].detach().numpy() for k in range(2)], 1), report=rep)
        u = eng.ws("u", B).cpu().numpy()
        _close("main", u[:, :64], t["main"].detach().numpy(), report=rep)
        _close("sub", u[:, 84:148], t["sub"].detach().numpy(), report=rep)
    _close("zt1", eng.ws("zt1", B).cpu().numpy(), np.concatenate([t[f"tower{k}.z1"].detach().numpy() for k in range(2)], 1), report=rep)
    _close("logits", eng.ws("logits", B).cpu().numpy(), t["logits"].detach().numpy(), report=rep)
    # ---- backward
    eng.backward(db)
    torch.cuda.synchronize()
    _close("d_logits", eng.ws("d_logits", B).cpu().numpy(), t["logits"].grad.numpy(), rtol=2e-5, report=rep)
    _close("d_x", eng.ws("d_x", B).cpu().numpy(), t["x"].grad.numpy(), rtol=5e-5, report=rep)
    l2 = om.hp["layer_l2"]
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
    # ---- apply: merged sparse gradients, clip norms, losses
    tables0 = {k: eng.pool[k].clone() for k in ("item_w", "cate_w", "ulong_w", "ushort_w")}
    losses = eng.apply_gradients(db).cpu().numpy()
    torch.cuda.synchronize()
    lr = ref["losses"]
    for i, k in enumerate(("loss", "data_loss", "regular_loss", "auxiliary_data_loss")):
        assert abs(losses[i] - lr[k]) <= 1e-5 * max(abs(lr[k]), 1e-3), (k, losses[i], lr[k])
    assert losses[4] == 0.0
    nun = eng.ws("sp.nuniq").cpu().numpy()
    has0 = eng.ws("sib.has0").cpu().numpy()
    el2 = om.hp["embed_l2"]
    for tab, idx, name in (("item", 0, EMB + "item_embedding"), ("cate", 1, EMB + "cate_embedding")):
        n = int(nun[idx])
        uk = eng.ws(f"sp.{tab}.ukeys").cpu().numpy()[:n]
        acc = eng.ws(f"sp.{tab}.accum").cpu().numpy()[:n].astype(np.float64)
        gref = ref["grads"][name].numpy()
        hk, sk, tk = (("item_history", "satisfied_item_history", "items") if tab == "item" else ("item_cate_history", "satisfied_cate_history", "cates"))
        involved = np.unique(np.concatenate([batch[hk].reshape(-1), batch[tk].reshape(-1)]))
        touched = np.unique(np.concatenate([involved, batch[sk].reshape(-1)]))
        assert np.array_equal(uk, touched), "unique ids differ"
        assert int(has0[idx]) == int(0 in involved)
        l2row = np.isin(uk, involved).astype(np.float64)[:, None]
        g = acc + el2 * l2row * tables0[tab + "_w"].cpu().numpy()[uk]
        _close(f"sparse grad {tab}", g, gref[uk], rtol=1e-4, report=rep)
        rest = np.ones(gref.shape[0], bool); rest[uk] = False
        assert not gref[rest].any()
    if min_len == T:
        assert has0[0] == 0 and has0[1] == 0, "this case is meant to exercise the row-0-without-L2 path"
    spn = eng.ws("sp_normsq").cpu().numpy()
    for i, name in enumerate(("item_embedding", "cate_embedding", "user_long_embedding", "user_short_embedding")):
        want_sq = ref["sqnorms"][EMB + name]
        assert abs(spn[i] - want_sq) <= 2e-4 * want_sq + 1e-30, (name, spn[i], want_sq)
    segn = eng.ws("seg_normsq").cpu().numpy()
    for s, (name, d) in enumerate(eng.info[L.POOL_DENSE].items()):
        want_sq = ref["sqnorms"][name]
        assert abs(segn[s] - want_sq) <= 2e-4 * want_sq + d["numel"] * (2e-6 * gmax) ** 2, (name, segn[s], want_sq)
    for tab in ("item", "cate", "ulong", "ushort"):
        w0 = tables0[tab + "_w"].cpu().numpy(); w1 = eng.pool[tab + "_w"].cpu().numpy()
        assert np.isfinite(w1).all() and (w0 != w1).any()
    print("\n".join(f"  {n:80s} {e:.2e}" for n, e in sorted(rep, key=lambda x: -x[1])[:10]))
    eng.close()


@pytest.mark.parametrize("model", ["mmoe", "ple", "sharebottom"])
def test_sibling_multi_step_and_eval(model):
    """Five optimisation steps (tables, dense variables, BN moving statistics all move), then a scoring pass."""
    nu, ni, nc, T, B = 300, 3000, 50, 50, 100
    om, eng = _setup(model, nu, ni, nc, T, B, seed=3)
    for step in range(5):
        batch = _batch(100 + step, B, T, nu, ni, nc, min_len=(T if step == 2 else 1))
        ref = om.train_step(batch)
        got = eng.train_step(eng.upload(batch)).cpu().numpy()
        for i, k in enumerate(("loss", "data_loss", "regular_loss", "auxiliary_data_loss")):
            r = ref["losses"][k]
            assert abs(got[i] - r) <= 5e-5 * max(abs(r), 1e-3), (step, k, got[i], r)
    got_vars = eng.get_variables()
    lr = om.hp["learning_rate"]
    for name in (EMB + "item_embedding", EMB + "cate_embedding", EMB + "user_long_embedding"):
        d = np.abs(got_vars[name] - om.params[name].numpy()).max()
        assert d <= 2.5 * lr, (name, d)          # Adam moves a weight by at most ~lr per step; trajectories differ by rounding only
    for name, tval in om.bn_state.items():
        assert np.allclose(got_vars[name], tval.numpy(), rtol=2e-4, atol=2e-4), name
    ev = _batch(999, 77, T, nu, ni, nc, grouped=False)
    pred = eng.forward(eng.upload(ev, training=False), training=False).cpu().numpy()
    want = om.eval_forward(ev).t["pred"].numpy().reshape(-1)
    assert np.abs(pred - want).max() <= 1e-4, float(np.abs(pred - want).max())
    eng.close()
