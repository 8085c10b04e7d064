"""TEST INFRASTRUCTURE - CPU restatement (torch, fp64 by default) of the reference's sibling multi-task baselines that share
PAMRec's input pipeline, MLP blocks and towers:

    MMoEModel_original   models/sequential/mmoe.py        (MM)
    PLEModel             models/sequential/ple.py         (PLE)
    ShareBottomModel     models/sequential/sharebottom.py (SB)
    SASRecModel          models/sequential/sasrec.py      (SAS; single task, see the second half of this file)

The first three are: DIN-style attention pooling of the satisfied-only history ("long_term") and of the full history ("short_term")
against the target item (`_attention_fcn`, MM:299-337), a mixing layer over concat(long, short, target) - MMoE (MM:26-50), PLE
(PLE:25-59) or none (SB:196-203) -, and two towers (`logit_fcn` on the satisfied label, `valid_logit_fcn` on the play label,
MM:175-179) trained with  loss = data + regular + 0.5 * auxiliary  (MM:52-82; the 0.5 is a literal, not hparams.fuzhu_weight).
Optimizer, clipping, BN and L2 semantics are those of the shared base classes, restated in oracle/pamrec_oracle.py; this file
adds only what differs: the variable inventory, the forward pass and the losses.

PARITY UNPINNED: like the float half of oracle/pamrec_oracle.py (TensorFlow cannot run here; the reference has no tests or golden
vectors for these models).  No product code exists for these models yet - this is the oracle the device path will be built
against; only tests/ may import it.
"""
import numpy as np
import torch

from . import pamrec_oracle as O

E_DIM = O.I_DIM + O.C_DIM                  # 20: item + category embedding = target / history token width
ATT_SIZES = (80, 40)                       # config/mmoe.yaml att_fcn_layer_sizes
MODELS = ("mmoe", "ple", "sharebottom")


def param_spec(model, n_users, n_items, n_cates, expert_num=5, share_expert_num=3, independent_expert_num=2,
               expert_sizes=(100, 64), gate_sizes=(64, 5), tower_sizes=(100, 64)):
    """TF variable inventory -> (params [(name, shape, init, group)], bn layers [(scope, channels)]).
    Groups as in pamrec_oracle.param_spec: `embed` tables get L2 on the rows a batch involves (SBM:651-664, MM:140-149),
    `layer` variables get full L2 (everything trainable outside sequential/embedding, SBM:714-721), `frozen` ones neither
    gradient nor L2 (user_embedding and play_lookup are created but never read, SBM:571, MM:118-122)."""
    assert model in MODELS
    emb = "sequential/embedding/"
    P = [(emb + "user_embedding", (n_users, O.U_DIM), "tn", "frozen"),
         (emb + "item_embedding", (n_items, O.I_DIM), "tn", "embed"),
         (emb + "cate_embedding", (n_cates, O.C_DIM), "tn", "embed"),
         (emb + "looptimes_embedding", (10, O.C_DIM), "tn", "frozen"),
         (emb + "user_long_embedding", (n_users, O.U_DIM), "tn", "embed_l2only"),
         (emb + "user_short_embedding", (n_users, O.U_DIM), "tn", "embed_l2only"),
         (emb + "play_lookup", (10, 40), "tn", "frozen")]
    BN = []

    def mlp(prefix, in_dim, sizes, out=False):
        s, b = O._mlp_spec(prefix, in_dim, sizes, out=out)
        P.extend(s)
        BN.extend(b)
    for branch in ("long_term", "short_term"):                                    # MM:217-226
        pre = f"sequential/clsr/{branch}/attention_fcn"
        P.append((pre + "/attention_mat", (E_DIM, E_DIM), "tn", "layer"))         # MM:315-319
        mlp(pre + "/att_fcn", 4 * E_DIM, ATT_SIZES, out=True)                     # MM:327-329 (_fcn_net: linear output)
    x_dim = 3 * E_DIM
    if model == "mmoe":
        for j in range(expert_num):
            mlp(f"sequential/clsr/expert_{j}", x_dim, expert_sizes)               # MM:38-41
        for g in ("gate_main", "gate_sub"):
            mlp(f"sequential/clsr/{g}", x_dim, gate_sizes)                        # MM:43-44
        tower_in = expert_sizes[-1] + E_DIM                                       # MM:231-232
    elif model == "ple":
        for j in range(share_expert_num):
            mlp(f"sequential/clsr/share_expert_{j}", x_dim, expert_sizes)         # PLE:38-40
        for task in ("main", "sub"):
            for j in range(independent_expert_num):
                mlp(f"sequential/clsr/{task}_expert_{j}", x_dim, expert_sizes)    # PLE:44-49
        for task in ("main", "sub"):
            mlp(f"sequential/clsr/gate_{task}", x_dim, gate_sizes)                # PLE:53-54
        tower_in = expert_sizes[-1] + E_DIM                                       # PLE:230-231
    else:
        tower_in = x_dim                                                          # SB:200-201
    for tw in ("sequential/valid_logit_fcn", "sequential/logit_fcn"):             # MM:175-177 (creation order)
        mlp(tw, tower_in, tower_sizes, out=True)
    return P, BN


def init_params(spec, bn_spec, seed=8, init_value=0.01):
    g = torch.Generator().manual_seed(seed)
    params, bn_state = {}, {}
    for name, shape, init, _ in spec:
        if init == "zeros":
            t = torch.zeros(shape)
        elif init == "ones":
            t = torch.ones(shape)
        else:
            t = torch.empty(shape)
            torch.nn.init.trunc_normal_(t, std=init_value, a=-2 * init_value, b=2 * init_value, generator=g)
        params[name] = t
    for scope, c in bn_spec:
        bn_state[scope + "/moving_mean"] = torch.zeros(c)
        bn_state[scope + "/moving_variance"] = torch.ones(c)
    return params, bn_state


def add_satisfied_fields(batch, seed=0):
    """make_batch of pamrec_oracle leaves out the satisfied-only copies of the history (dead for PAMRec, IT:1069-1103): build
    them the way `_convert_data` does - the satisfied entries of each row, compacted to the left, zero padded."""
    rng = np.random.default_rng(seed)
    ih, ch, mask = np.asarray(batch["item_history"]), np.asarray(batch["item_cate_history"]), np.asarray(batch["mask"])
    sat = (rng.random(ih.shape) < 0.5) & (mask == 1)
    order = np.argsort(~sat, axis=1, kind="stable")
    keep = np.arange(ih.shape[1])[None, :] < sat.sum(1)[:, None]
    out = dict(batch)
    out["satisfied_item_history"] = np.where(keep, np.take_along_axis(ih, order, 1), 0).astype(np.int32)
    out["satisfied_cate_history"] = np.where(keep, np.take_along_axis(ch, order, 1), 0).astype(np.int32)
    out["satisfied_mask"] = keep.astype(np.int32)
    return out


def _attention_fcn(ctx, query, hist, mask, scope, tag=None):
    """MM:299-337.  query [B, 20], hist [B, T, 20], mask [B, T] -> hist * softmax weights [B, T, 20]."""
    B, T, _ = hist.shape
    att_inputs = hist @ ctx.p[scope + "/attention_mat"]                                    # tensordot over the feature axis
    q = query[:, None, :].expand(B, T, query.shape[1])
    feats = torch.cat([att_inputs, q, att_inputs - q, att_inputs * q], -1)                # [B, T, 80]
    score = O._mlp(ctx, feats, scope + "/att_fcn", ATT_SIZES, out=True, tag=tag).squeeze(-1)   # BN statistics over all B*T rows
    pad = torch.full_like(score, float(-(2 ** 32) + 1))
    w = torch.softmax(torch.where(mask == 1, score, pad), dim=-1)                          # an all-padding row: uniform weights
    if tag:
        ctx.t[tag + ".feat"], ctx.t[tag + ".score"], ctx.t[tag + ".w"] = feats, score, w
    return hist * w[..., None]


def forward(model, p, bn_state, batch, training, dtype=torch.float64, rows=None, relu_masks=None, **sizes):
    """-> ctx with ctx.t["logits"] [B, 2] = (logit of `labels` (satisfied), valid_logit of `labels_play`) and ctx.t["pred"].
    relu_masks: activation pattern to differentiate on, keyed by BN scope (pamrec_oracle._act)."""
    expert_num = sizes.get("expert_num", 5)
    share_n, indep_n = sizes.get("share_expert_num", 3), sizes.get("independent_expert_num", 2)
    expert_sizes, gate_sizes = sizes.get("expert_sizes", (100, 64)), sizes.get("gate_sizes", (64, 5))
    tower_sizes = sizes.get("tower_sizes", (100, 64))
    ctx = O._Ctx(p, bn_state, training, dtype, relu_masks=relu_masks)
    emb = "sequential/embedding/"
    idx = lambda k: torch.as_tensor(np.asarray(batch[k])).long()
    item_w, cate_w = p[emb + "item_embedding"], p[emb + "cate_embedding"]
    if rows is None:
        rows = {"hist_item": item_w[idx("item_history")], "hist_cate": cate_w[idx("item_cate_history")],
                "sat_item": item_w[idx("satisfied_item_history")], "sat_cate": cate_w[idx("satisfied_cate_history")],
                "tgt_item": item_w[idx("items")], "tgt_cate": cate_w[idx("cates")]}
    target = torch.cat([rows["tgt_item"], rows["tgt_cate"]], -1)                            # SBM:669-671
    long_in = torch.cat([rows["sat_item"], rows["sat_cate"]], -1)                           # MM:199-201
    short_in = torch.cat([rows["hist_item"], rows["hist_cate"]], -1)                        # MM:208-210
    long = _attention_fcn(ctx, target, long_in, idx("satisfied_mask"), "sequential/clsr/long_term/attention_fcn", tag="att0").sum(1)
    short = _attention_fcn(ctx, target, short_in, idx("mask"), "sequential/clsr/short_term/attention_fcn", tag="att1").sum(1)
    x = torch.cat([long, short, target], -1)                                                # [B, 60]
    ctx.t["x"] = x

    # a gate is a BN + ReLU MLP like any other (no softmax, MM:43-46): [B, E] x experts [B, E, 64] -> [B, 64]
    if model == "mmoe":
        scopes = [f"sequential/clsr/expert_{j}" for j in range(expert_num)]
        experts = torch.stack([O._mlp(ctx, x, s, expert_sizes, tag=f"expert{j}") for j, s in enumerate(scopes)], 1)
        main = (O._mlp(ctx, x, "sequential/clsr/gate_main", gate_sizes, tag="gate0")[:, None, :] @ experts).squeeze(1)
        sub = (O._mlp(ctx, x, "sequential/clsr/gate_sub", gate_sizes, tag="gate1")[:, None, :] @ experts).squeeze(1)
    elif model == "ple":
        # tags number the experts in the device layout's order: shared 0-2, main 3-4, sub 5-6
        share = [O._mlp(ctx, x, f"sequential/clsr/share_expert_{j}", expert_sizes, tag=f"expert{j}") for j in range(share_n)]
        own = {t: [O._mlp(ctx, x, f"sequential/clsr/{t}_expert_{j}", expert_sizes, tag=f"expert{share_n + k * indep_n + j}")
                   for j in range(indep_n)] for k, t in enumerate(("main", "sub"))}
        outs = {}
        for k, t in enumerate(("main", "sub")):                                            # PLE:51-58: shared experts first
            gate = O._mlp(ctx, x, f"sequential/clsr/gate_{t}", gate_sizes, tag=f"gate{k}")
            outs[t] = (gate[:, None, :] @ torch.stack(share + own[t], 1)).squeeze(1)
        main, sub = outs["main"], outs["sub"]
    else:
        main = sub = None
    if model == "sharebottom":
        model_output = valid_output = x                                                    # SB:200-201
    else:
        model_output, valid_output = torch.cat([main, target], -1), torch.cat([sub, target], -1)
    valid_logit = O._mlp(ctx, valid_output, "sequential/valid_logit_fcn", tower_sizes, out=True, tag="tower1")    # MM:175
    logit = O._mlp(ctx, model_output, "sequential/logit_fcn", tower_sizes, out=True, tag="tower0")               # MM:177
    if main is not None:
        ctx.t["main"], ctx.t["sub"] = main, sub
    ctx.t["logits"] = torch.cat([logit, valid_logit], -1)
    ctx.t["pred"] = torch.sigmoid(logit)                                                           # BM:93-113
    return ctx


def losses(ctx, spec, batch, hp):
    """MM:52-82 + BM:122-134 / SBM:714-721.  hp: embed_l2, layer_l2."""
    p, dtype = ctx.p, ctx.dtype
    logits = ctx.t["logits"]
    y_sat = torch.as_tensor(np.asarray(batch["labels_satisfied"])).to(dtype).reshape(-1)
    y_play = torch.as_tensor(np.asarray(batch["labels_play"])).to(dtype).reshape(-1)
    data = O._sigmoid_xent(logits[:, 0], y_sat).mean()
    aux = 0.5 * O._sigmoid_xent(logits[:, 1], y_play).mean()
    emb = "sequential/embedding/"
    cat = lambda *ks: torch.unique(torch.cat([torch.as_tensor(np.asarray(batch[k])).long().reshape(-1) for k in ks]))
    involved = {emb + "item_embedding": cat("item_history", "items"),                       # SBM:640-664: the FULL history and the
                emb + "cate_embedding": cat("item_cate_history", "cates"),                  # target; satisfied ids are a subset
                emb + "user_long_embedding": cat("users"), emb + "user_short_embedding": cat("users")}
    reg = torch.zeros((), dtype=dtype)
    for name, ids in involved.items():
        reg = reg + hp["embed_l2"] * 0.5 * (p[name][ids] ** 2).sum()
    for name, _, _, grp in spec:
        if grp == "layer":
            reg = reg + hp["layer_l2"] * 0.5 * (p[name] ** 2).sum()
    return {"loss": data + reg + aux, "data_loss": data, "regular_loss": reg, "auxiliary_data_loss": aux}


class SiblingOracle:
    """Weights + one differentiable evaluation of the loss; the update rule is pamrec_oracle.OracleModel's."""

    def __init__(self, model, n_users, n_items, n_cates, hp=None, seed=8, dtype=torch.float64, **sizes):
        self.model, self.sizes, self.dtype = model, sizes, dtype
        self.hp = dict(embed_l2=1e-4, layer_l2=1e-4)
        self.hp.update(hp or {})
        self.spec, self.bn_spec = param_spec(model, n_users, n_items, n_cates,
                                             **{k: v for k, v in sizes.items() if k in ("expert_num", "share_expert_num",
                                                                                       "independent_expert_num", "expert_sizes",
                                                                                       "gate_sizes", "tower_sizes")})
        self.params, self.bn_state = init_params(self.spec, self.bn_spec, seed=seed)

    def loss_and_grads(self, batch):
        p = {n: t.detach().clone().to(self.dtype).requires_grad_(True) for n, t in self.params.items()}
        ctx = forward(self.model, p, self.bn_state, batch, True, self.dtype, **self.sizes)
        out = losses(ctx, self.spec, batch, self.hp)
        out["loss"].backward()
        grads = {n: (t.grad if t.grad is not None else torch.zeros_like(t)) for n, t in p.items()}
        return {k: float(v.detach()) for k, v in out.items()}, grads, ctx

    def eval_forward(self, batch):
        with torch.no_grad():
            p = {n: t.to(self.dtype) for n, t in self.params.items()}
            return forward(self.model, p, self.bn_state, batch, False, self.dtype, **self.sizes)


# ===================================================================================================================== SASRec
# SASRecModel (SAS:16-96): the satisfied-only history (item || category, 20 wide) plus a learned position table goes through two
# pre-LN self-attention blocks with DENSE projections (tf.layers.dense with bias, SAS:268-270 - PAMRec's time-aware tables
# replace exactly these), one head, key mask = satisfied_mask, no query mask, no causality (SAS:89); the state at the last
# satisfied position (SAS:72-78) is concatenated with the target and fed to one tower (`logit_fcn`, SBM:76-79).  Loss = data +
# regular (BM:136-151 / SBM single task).
SAS_D = E_DIM


def sasrec_param_spec(n_users, n_items, n_cates, T, tower_sizes=(100, 64)):
    emb = "sequential/embedding/"
    P = [(emb + "user_embedding", (n_users, O.U_DIM), "tn", "frozen"),
         (emb + "item_embedding", (n_items, O.I_DIM), "tn", "embed"),
         (emb + "cate_embedding", (n_cates, O.C_DIM), "tn", "embed"),
         (emb + "looptimes_embedding", (10, O.C_DIM), "tn", "frozen"),
         (emb + "position_embedding", (T, SAS_D), "tn", "pos")]                          # SAS:29-34 (add_feature False)
    for b in range(2):
        pre = f"sequential/sasrec/num_blocks_{b}/"
        P += [(pre + "ln/Variable", (SAS_D,), "zeros", "layer"), (pre + "ln/Variable_1", (SAS_D,), "ones", "layer")]
        for name in ("dense", "dense_1", "dense_2"):                                      # Q, K, V in creation order
            P += [(pre + f"self_attention/{name}/kernel", (SAS_D, SAS_D), "glorot", "layer"),
                  (pre + f"self_attention/{name}/bias", (SAS_D,), "zeros", "layer")]
        P += [(pre + "ln_1/Variable", (SAS_D,), "zeros", "layer"), (pre + "ln_1/Variable_1", (SAS_D,), "ones", "layer")]
        for name in ("conv1d", "conv1d_1"):
            P += [(pre + f"multihead_attention/{name}/kernel", (1, SAS_D, SAS_D), "glorot", "layer"),
                  (pre + f"multihead_attention/{name}/bias", (SAS_D,), "zeros", "layer")]
    s, BN = O._mlp_spec("sequential/logit_fcn", 2 * SAS_D, tower_sizes, out=True)
    return P + s, BN


def sasrec_init(spec, bn_spec, seed=8, init_value=0.01):
    params, bn_state = init_params([x for x in spec if x[2] != "glorot"], bn_spec, seed=seed, init_value=init_value)
    g = torch.Generator().manual_seed(seed + 1)
    for name, shape, init, _ in spec:
        if init == "glorot":                                                              # tf.layers default initializer
            fan_in, fan_out = shape[-2], shape[-1]
            lim = (6.0 / (fan_in + fan_out)) ** 0.5
            params[name] = (torch.rand(shape, generator=g) * 2 - 1) * lim
    return {n: params[n] for n, _, _, _ in spec}, bn_state


def sasrec_forward(p, bn_state, batch, training, dtype=torch.float64, tower_sizes=(100, 64), rows=None, relu_masks=None):
    """rows: optional pre-gathered lookups (sat_item, sat_cate, tgt_item, tgt_cate, pos [B, T, 20]) so that a caller can observe
    their per-lookup gradients; by default they are read from the tables here.  relu_masks: see pamrec_oracle._act (keys
    "blk{b}.ffn" for the point-wise FFN, the BN scopes of the tower)."""
    ctx = O._Ctx(p, bn_state, training, dtype, relu_masks=relu_masks)
    emb = "sequential/embedding/"
    idx = lambda k: torch.as_tensor(np.asarray(batch[k])).long()
    mask = idx("satisfied_mask")
    if rows is None:
        rows = {"sat_item": p[emb + "item_embedding"][idx("satisfied_item_history")],
                "sat_cate": p[emb + "cate_embedding"][idx("satisfied_cate_history")],
                "tgt_item": p[emb + "item_embedding"][idx("items")], "tgt_cate": p[emb + "cate_embedding"][idx("cates")],
                "pos": p[emb + "position_embedding"][None].expand(mask.shape[0], -1, -1)}   # SAS:39-44: tile(range(T)) lookup
    seq = torch.cat([rows["sat_item"], rows["sat_cate"]], -1) + rows["pos"]                # SAS:61-64
    ctx.t["x0"] = seq
    target = torch.cat([rows["tgt_item"], rows["tgt_cate"]], -1)
    pad = float(-(2 ** 32) + 1)
    for b in range(2):
        pre = f"sequential/sasrec/num_blocks_{b}/"
        q_in = O._ln(seq, p[pre + "ln/Variable"], p[pre + "ln/Variable_1"])               # queries = LN(seq), keys = seq (SAS:89-90)
        dense = lambda x, n: x @ p[pre + f"self_attention/{n}/kernel"] + p[pre + f"self_attention/{n}/bias"]
        Q, K, V = dense(q_in, "dense"), dense(seq, "dense_1"), dense(seq, "dense_2")
        s = Q @ K.transpose(1, 2) / (SAS_D ** 0.5)                                        # SAS:281-284
        s = torch.where(mask[:, None, :] == 0, torch.full_like(s, pad), s)                # SAS:288-293 key mask only
        y = torch.softmax(s, -1) @ V + q_in                                               # SAS:318-324 residual on the queries
        f = O._ln(y, p[pre + "ln_1/Variable"], p[pre + "ln_1/Variable_1"])
        hid = O._act(ctx, f @ p[pre + "multihead_attention/conv1d/kernel"][0] + p[pre + "multihead_attention/conv1d/bias"], f"blk{b}.ffn")
        seq = hid @ p[pre + "multihead_attention/conv1d_1/kernel"][0] + p[pre + "multihead_attention/conv1d_1/bias"] + f   # SAS:127-141
        ctx.t[f"blk{b}.qin"], ctx.t[f"blk{b}.Q"], ctx.t[f"blk{b}.K"], ctx.t[f"blk{b}.V"] = q_in, Q, K, V
        ctx.t[f"blk{b}.y"], ctx.t[f"blk{b}.f"], ctx.t[f"blk{b}.out"] = y, f, seq
    length = mask.sum(1)
    # SAS:72-78 reads seq[b, length - 1]: index -1 for a row without any satisfied item, which tf.gather_nd rejects on CPU and
    # answers with zeros on GPU.  The iterator can produce such rows; the restatement follows the GPU behaviour.
    last = torch.clamp(length - 1, min=0)
    final = seq[torch.arange(seq.shape[0]), last] * (length > 0).to(seq.dtype)[:, None]
    ctx.t["final_state"] = final
    logit = O._mlp(ctx, torch.cat([final, target], -1), "sequential/logit_fcn", tower_sizes, out=True, tag="tower0")
    ctx.t["logits"] = logit
    ctx.t["pred"] = torch.sigmoid(logit)
    return ctx


def sasrec_losses(ctx, spec, batch, hp):
    p, dtype = ctx.p, ctx.dtype
    y = torch.as_tensor(np.asarray(batch["labels_satisfied"])).to(dtype).reshape(-1)
    data = O._sigmoid_xent(ctx.t["logits"][:, 0], y).mean()                               # BM:195-205
    emb = "sequential/embedding/"
    cat = lambda *ks: torch.unique(torch.cat([torch.as_tensor(np.asarray(batch[k])).long().reshape(-1) for k in ks]))
    reg = hp["embed_l2"] * 0.5 * ((p[emb + "item_embedding"][cat("item_history", "items")] ** 2).sum()
                                  + (p[emb + "cate_embedding"][cat("item_cate_history", "cates")] ** 2).sum())
    for name, _, _, grp in spec:
        if grp == "layer":
            reg = reg + hp["layer_l2"] * 0.5 * (p[name] ** 2).sum()                       # position table is under /embedding: no L2
    return {"loss": data + reg, "data_loss": data, "regular_loss": reg}


# ===================================================================================================================== train step
class SiblingOracleModel(O.OracleModel):
    """pamrec_oracle.OracleModel (per-lookup sparse gradients, per-tensor tf.clip_by_norm, TF Adam with dense decay of sparse
    variables, BN moving averages) driven by a sibling model's inventory, forward pass and losses.
    model: "mmoe" | "ple" | "sharebottom" | "sasrec"."""

    def __init__(self, model, n_users, n_items, n_cates, T, hp=None, seed=8, dtype=torch.float64, **sizes):
        assert model in MODELS + ("sasrec",)
        self.model, self.sizes = model, sizes
        super().__init__(n_users, n_items, n_cates, T, hp=hp, seed=seed, dtype=dtype)

    def _spec(self):
        nu, ni, nc, T = self.dims
        if self.model == "sasrec":
            return sasrec_param_spec(nu, ni, nc, T, **{k: v for k, v in self.sizes.items() if k == "tower_sizes"})
        return param_spec(self.model, nu, ni, nc, **self.sizes)

    def _init(self, seed):
        return sasrec_init(self.spec, self.bn_spec, seed=seed) if self.model == "sasrec" else init_params(self.spec, self.bn_spec, seed=seed)

    def _gather(self, p, batch):
        """Every embedding_lookup as its own tensor (SBM:598-668, MM:131-149): full history, satisfied-only history, target,
        and the `involved` rows that only carry the L2 term."""
        idx = lambda k: torch.as_tensor(np.asarray(batch[k])).long()
        emb = "sequential/embedding/"
        uniq = lambda *ks: torch.unique(torch.cat([idx(k).reshape(-1) for k in ks]))
        g = {"hist_item": (emb + "item_embedding", idx("item_history")), "hist_cate": (emb + "cate_embedding", idx("item_cate_history")),
             "sat_item": (emb + "item_embedding", idx("satisfied_item_history")),
             "sat_cate": (emb + "cate_embedding", idx("satisfied_cate_history")),
             "tgt_item": (emb + "item_embedding", idx("items")), "tgt_cate": (emb + "cate_embedding", idx("cates")),
             "inv_item": (emb + "item_embedding", uniq("item_history", "items")),
             "inv_cate": (emb + "cate_embedding", uniq("item_cate_history", "cates"))}
        if self.model != "sasrec":
            g["inv_ulong"] = (emb + "user_long_embedding", uniq("users"))
            g["inv_ushort"] = (emb + "user_short_embedding", uniq("users"))
        else:                                                     # SAS:39-44: one lookup row per (sample, position)
            B, T = idx("satisfied_mask").shape
            g["pos"] = (emb + "position_embedding", torch.arange(T)[None, :].expand(B, T))
        return g, {k: p[name][i] for k, (name, i) in g.items()}

    def _forward(self, p, batch, training, rows=None, relu_masks=None):
        if self.model == "sasrec":
            return sasrec_forward(p, self.bn_state, batch, training, self.dtype, rows=rows, relu_masks=relu_masks, **self.sizes)
        return forward(self.model, p, self.bn_state, batch, training, self.dtype, rows=rows, relu_masks=relu_masks, **self.sizes)

    def _losses(self, ctx, batch, rows):
        dtype, hp = self.dtype, self.hp
        y = lambda k: torch.as_tensor(np.asarray(batch[k])).to(dtype).reshape(-1)
        out = {"data_loss": O._sigmoid_xent(ctx.t["logits"][:, 0], y("labels_satisfied")).mean()}
        if self.model != "sasrec":
            out["auxiliary_data_loss"] = 0.5 * O._sigmoid_xent(ctx.t["logits"][:, 1], y("labels_play")).mean()
        reg = torch.zeros((), dtype=dtype)
        for k in rows:
            if k.startswith("inv_"):
                reg = reg + hp["embed_l2"] * 0.5 * (rows[k] ** 2).sum()
        for name, _, _, grp in self.spec:
            if grp == "layer":
                reg = reg + hp["layer_l2"] * 0.5 * (ctx.p[name] ** 2).sum()
        out["regular_loss"] = reg
        out["loss"] = sum(out.values())
        return out
