"""oracle/siblings_oracle.py (MMoEModel_original / PLEModel / ShareBottomModel) checked for internal
consistency on CPU: inventory, the structural identities between the three models, masking, gradients.  There is no device
path for these models yet - this pins the restatement the kernels will be built against."""
import numpy as np
import pytest
import torch

from oracle import pamrec_oracle as O
from oracle import siblings_oracle as S

NU, NI, NC, T, B = 30, 120, 9, 12, 20


def _batch(seed=3):
    return S.add_satisfied_fields(O.make_batch(seed, B, T, NU, NI, NC), seed=seed)


def _randomise(m, seed):
    g = torch.Generator().manual_seed(seed)
    for n, t in m.params.items():
        m.params[n] = t + 0.1 * torch.randn(t.shape, generator=g)
    for n, t in m.bn_state.items():
        m.bn_state[n] = t + 0.1 * torch.rand(t.shape, generator=g)


@pytest.mark.parametrize("model", S.MODELS)
def test_inventory_and_forward_shapes(model):
    m = S.SiblingOracle(model, NU, NI, NC)
    names = [n for n, _, _, _ in m.spec]
    assert len(names) == len(set(names))
    n_mlps = {"mmoe": 2 + 5 + 2 + 2, "ple": 2 + 3 + 4 + 2 + 2, "sharebottom": 2 + 2}[model]
    assert sum(n.endswith("w_nn_layer0") for n in names) == n_mlps
    assert sum(n.endswith("w_nn_output") for n in names) == 4                     # two att_fcn scorers + two towers
    shape = dict((n, s) for n, s, _, _ in m.spec)
    assert shape["sequential/clsr/long_term/attention_fcn/att_fcn/nn_part/w_nn_layer0"] == (80, 80)
    assert shape["sequential/logit_fcn/nn_part/w_nn_layer0"] == ((60, 100) if model == "sharebottom" else (84, 100))
    assert {s + "/moving_mean" for s, _ in m.bn_spec} | {s + "/moving_variance" for s, _ in m.bn_spec} == set(m.bn_state)
    _randomise(m, 1)
    batch = _batch()
    out, grads, ctx = m.loss_and_grads(batch)
    assert ctx.t["logits"].shape == (B, 2) and ctx.t["x"].shape == (B, 60)
    assert abs(out["loss"] - (out["data_loss"] + out["regular_loss"] + out["auxiliary_data_loss"])) < 1e-12
    # every BN layer saw a batch; variables that are created but never read get no gradient
    assert set(ctx.new_bn) == {s for s, _ in m.bn_spec}
    for n, _, _, grp in m.spec:
        if grp == "frozen":
            assert not grads[n].any(), n
        elif grp == "layer":
            assert grads[n].abs().max() > 0, n
    # user_long / user_short: L2 on the batch's users only (MM:140-149), no data gradient
    users = np.unique(batch["users"])
    g = grads["sequential/embedding/user_long_embedding"]
    assert torch.allclose(g[users], m.hp["embed_l2"] * m.params["sequential/embedding/user_long_embedding"].double()[users])
    mask = np.ones(NU, bool); mask[users] = False
    assert not g[mask].any()


def test_ple_with_only_shared_experts_is_mmoe():
    """PLE with 5 shared and 0 task experts routes both gates over the same 5 experts: MMoE under a renaming of scopes."""
    a = S.SiblingOracle("mmoe", NU, NI, NC, expert_num=5)
    b = S.SiblingOracle("ple", NU, NI, NC, share_expert_num=5, independent_expert_num=0)
    _randomise(a, 2)
    for n in a.params:
        m = n.replace("/clsr/expert_", "/clsr/share_expert_")
        assert m in b.params, m
        b.params[m] = a.params[n].clone()
    for n in a.bn_state:
        b.bn_state[n.replace("/clsr/expert_", "/clsr/share_expert_")] = a.bn_state[n].clone()
    batch = _batch(5)
    la, ga, ca = a.loss_and_grads(batch)
    lb, gb, cb = b.loss_and_grads(batch)
    assert torch.equal(ca.t["logits"], cb.t["logits"]) and la == lb
    assert torch.equal(a.eval_forward(batch).t["pred"], b.eval_forward(batch).t["pred"])


def test_attention_masking_and_padding():
    m = S.SiblingOracle("sharebottom", NU, NI, NC)
    _randomise(m, 4)
    batch = _batch(6)
    base = m.eval_forward(batch).t["logits"]
    # inference mode (moving statistics): ids under the mask do not matter ...
    other = dict(batch)
    for ids, mask, n in (("item_history", "mask", NI), ("satisfied_item_history", "satisfied_mask", NI),
                         ("item_cate_history", "mask", NC), ("satisfied_cate_history", "satisfied_mask", NC)):
        a = np.asarray(batch[ids]).copy()
        pad = np.asarray(batch[mask]) == 0
        a[pad] = np.random.default_rng(1).integers(1, n, size=int(pad.sum()))
        other[ids] = a
    assert torch.allclose(m.eval_forward(other).t["logits"], base, atol=1e-12)
    # ... while in training mode they do, through the batch statistics over all B*T positions (parity trap 5 of the survey)
    l0, _, c0 = m.loss_and_grads(batch)
    l1, _, c1 = m.loss_and_grads(other)
    assert not torch.allclose(c0.t["logits"], c1.t["logits"], atol=1e-9)
    # a row with no satisfied item at all: softmax over equal -2^32+1 scores = uniform weights over the T padding rows
    empty = dict(batch)
    sm = np.asarray(batch["satisfied_mask"]).copy(); sm[0] = 0
    si = np.asarray(batch["satisfied_item_history"]).copy(); si[0] = 0
    sc = np.asarray(batch["satisfied_cate_history"]).copy(); sc[0] = 0
    empty.update(satisfied_mask=sm, satisfied_item_history=si, satisfied_cate_history=sc)
    x = m.eval_forward(empty).t["x"]
    row0 = torch.cat([m.params["sequential/embedding/item_embedding"][0], m.params["sequential/embedding/cate_embedding"][0]]).double()
    assert torch.allclose(x[0, :20], row0, atol=1e-12)               # T copies of row 0, each weighted 1 / T


@pytest.mark.parametrize("model", S.MODELS)
def test_gradients_match_finite_differences(model):
    m = S.SiblingOracle(model, NU, NI, NC)
    _randomise(m, 7)
    batch = _batch(8)
    _, grads, _ = m.loss_and_grads(batch)
    for name, idx in (("sequential/clsr/short_term/attention_fcn/attention_mat", (3, 5)),
                      ("sequential/valid_logit_fcn/nn_part/w_nn_layer1", (10, 3)),
                      ("sequential/embedding/item_embedding", (int(batch["items"][0]), 2))):
        base = m.params[name].clone()
        vals = []
        for sgn in (+1, -1):
            m.params[name] = base.clone().double()
            m.params[name][idx] += sgn * 1e-6
            vals.append(m.loss_and_grads(batch)[0]["loss"])
        m.params[name] = base
        fd = (vals[0] - vals[1]) / 2e-6
        assert abs(fd - float(grads[name][idx])) <= 1e-5 * max(abs(fd), 1e-5), (name, fd, float(grads[name][idx]))


def test_sasrec_restatement():
    spec, bn = S.sasrec_param_spec(NU, NI, NC, T)
    params, bn_state = S.sasrec_init(spec, bn, seed=3)
    assert [n for n, _, _, _ in spec] == list(params) and len(set(params)) == len(params)
    assert params["sequential/sasrec/num_blocks_1/self_attention/dense_2/kernel"].shape == (20, 20)
    assert params["sequential/logit_fcn/nn_part/w_nn_layer0"].shape == (40, 100)
    g = torch.Generator().manual_seed(5)
    params = {n: t + 0.1 * torch.randn(t.shape, generator=g) for n, t in params.items()}
    batch = _batch(9)
    p = {n: t.double().requires_grad_(True) for n, t in params.items()}
    ctx = S.sasrec_forward(p, bn_state, batch, True)
    out = S.sasrec_losses(ctx, spec, batch, dict(embed_l2=1e-4, layer_l2=1e-4))
    out["loss"].backward()
    assert ctx.t["logits"].shape == (B, 1)
    # the read-out is the block output at the last satisfied position; rows without one read zeros
    length = np.asarray(batch["satisfied_mask"]).sum(1)
    b = int(np.argmax(length > 0))
    assert torch.equal(ctx.t["final_state"][b], ctx.t["blk1.out"][b, int(length[b]) - 1])
    none = dict(batch)
    sm = np.asarray(batch["satisfied_mask"]).copy(); sm[0] = 0
    none["satisfied_mask"] = sm
    with torch.no_grad():
        assert not S.sasrec_forward({n: t.double() for n, t in params.items()}, bn_state, none, False).t["final_state"][0].any()
    # key mask only: every query row (padded ones too) attends, so padded positions of the LAST block never reach the loss, while
    # the keys they would contribute are masked out -> gradient of the position rows beyond every length is zero
    longest = int(length.max())
    gpos = p["sequential/embedding/position_embedding"].grad
    assert gpos[:longest].abs().max() > 0 and (longest == T or not gpos[longest:].any())
    # finite differences on a projection weight
    name, idx = "sequential/sasrec/num_blocks_0/self_attention/dense_1/kernel", (4, 7)
    vals = []
    for sgn in (+1, -1):
        q = {n: t.double().clone() for n, t in params.items()}
        q[name][idx] += sgn * 1e-6
        with torch.no_grad():
            c = S.sasrec_forward(q, bn_state, batch, True)
            vals.append(float(S.sasrec_losses(c, spec, batch, dict(embed_l2=1e-4, layer_l2=1e-4))["loss"]))
    fd = (vals[0] - vals[1]) / 2e-6
    assert abs(fd - float(p[name].grad[idx])) <= 1e-5 * max(abs(fd), 1e-5)


@pytest.mark.parametrize("model", S.MODELS + ("sasrec",))
def test_full_train_step_of_every_sibling(model):
    """SiblingOracleModel = the PAMRec oracle's update rule (per-lookup sparse gradients, per-tensor clip, TF Adam, BN moving
    averages) around a sibling's graph: gradients agree with plain autograd through the tables, two steps move every live variable,
    frozen ones stay put."""
    hp = dict(embed_l2=1e-4, layer_l2=1e-4)
    m = S.SiblingOracleModel(model, NU, NI, NC, T, hp=hp, seed=4)
    O.perturb_params(m.params, m.bn_state, seed=5)
    batch = _batch(11)
    ref = m.train_step(batch, apply=False)
    # the same loss differentiated straight through the embedding tables
    p = {n: t.double().clone().requires_grad_(True) for n, t in m.params.items()}
    if model == "sasrec":
        ctx = S.sasrec_forward(p, m.bn_state, batch, True)
        want = S.sasrec_losses(ctx, m.spec, batch, hp)
    else:
        ctx = S.forward(model, p, m.bn_state, batch, True)
        want = S.losses(ctx, m.spec, batch, hp)
    want["loss"].backward()
    for k, v in want.items():
        assert abs(float(v.detach()) - ref["losses"][k]) < 1e-12, k
    for n, g in ref["grads"].items():
        assert torch.allclose(g, p[n].grad, rtol=1e-10, atol=1e-14), n
    # the clip norm of a table is taken over the un-deduplicated lookup rows (BM:297-303), not over the summed gradient
    name = "sequential/embedding/item_embedding"
    assert ref["sqnorms"][name] > 0 and abs(ref["sqnorms"][name] - float((ref["grads"][name] ** 2).sum())) > 0
    before = {n: t.clone() for n, t in m.params.items()}
    m.train_step(batch)
    m.train_step(_batch(12))
    assert m.step == 2
    for n, _, _, grp in m.spec:
        moved = bool((m.params[n] != before[n]).any())
        assert moved == (grp != "frozen"), (n, grp)
    assert all(bool((m.bn_state[s + "/moving_mean"] != 0).any()) for s, _ in m.bn_spec)
