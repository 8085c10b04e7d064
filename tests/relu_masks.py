"""The ReLU activation pattern of an engine forward pass, read back through the workspace and keyed like the oracle's
``relu_masks`` (oracle/pamrec_oracle.py:_act).

Why: the gradient of a ReLU network is only defined once the side of every kink is fixed.  A step has ~10^6 batch-normalised
pre-activations and a few always lie inside the fp32 rounding of the forward pass; when the fp32 engine and the fp64 oracle land
on different sides of one, that unit's gradient differs at O(1) and - through the batch statistics - every row's a little.
Round 1 detected such batches and retried on other seeds, which cannot work beyond B ~ 200.  Here the oracle differentiates on
the ENGINE's pattern instead, so gradient parity is defined at any batch size.

The head masks are recomputed from the stored pre-activations with the kernels' own fp32 expression
``fmaf(gamma, (z - mean) * invstd, beta) > 0`` (kernels_head.cu:bn_relu; the sign of an fma equals the sign of the exact value,
which float64 reproduces because a product of two float32 is exact in float64).  The point-wise FFN's pattern comes from the
PAMREC_DEBUG_SAVE_FFN_HIDDEN hook (its pre-activation is never stored)."""
import numpy as np

P_ = "sequential/pamrec/"
SCORE = P_ + "new_long/score_1/nn_part/"
TOWERS = ("sequential/logit_fcn", "sequential/valid_logit_fcn", "xilidu_logit_fcn")
# engine pre-activation buffer, BN set, columns per member, member scopes (oracle key = scope)
HEAD_BN = [
    ("ze0", "e0", 100, [f"{P_}expert_{j}/nn_part/batch_normalization" for j in range(5)]),
    ("ze1", "e1", 64, [f"{P_}expert_{j}/nn_part/batch_normalization_1" for j in range(5)]),
    ("zg0", "g0", 64, [f"{P_}gate_{g}/nn_part/batch_normalization" for g in ("main", "sub")]),
    ("zg1", "g1", 5, [f"{P_}gate_{g}/nn_part/batch_normalization_1" for g in ("main", "sub")]),
    ("zt0", "t0", 100, [s + "/nn_part/batch_normalization" for s in TOWERS]),
    ("zt1", "t1", 64, [s + "/nn_part/batch_normalization_1" for s in TOWERS]),
    ("z1", "s0", 20, [SCORE + "batch_normalization"]),
    ("z2", "s1", 1, [SCORE + "batch_normalization_1"]),
]


def _on(z, stat, gamma, beta):
    xh = (z.astype(np.float32) - stat[:, 0].astype(np.float32)) * stat[:, 1].astype(np.float32)      # two fp32 roundings, as on the device
    return gamma.astype(np.float64) * xh.astype(np.float64) + beta.astype(np.float64) > 0


def engine_relu_masks(eng, rows, ffn=True):
    """Masks of the last ``eng.forward(db, training=True)`` over its first `rows` rows: key -> bool array.
    ffn=True needs ``eng.set_debug(DEBUG_SAVE_FFN_HIDDEN)`` before that forward call."""
    out = {}
    if rows == 0:
        return out
    for zbuf, bn, width, scopes in HEAD_BN:
        z = eng.ws(zbuf, rows).cpu().numpy()
        st = eng.ws(f"bn.{bn}.stat").cpu().numpy()
        lead = z.shape[:-1] if zbuf != "z2" else z.shape
        z = z.reshape(-1, len(scopes) * width)
        for m, scope in enumerate(scopes):
            sl = slice(m * width, (m + 1) * width)
            g, b = eng.dense(scope + "/gamma").cpu().numpy(), eng.dense(scope + "/beta").cpu().numpy()
            out[scope] = _on(z[:, sl], st[sl], g, b).reshape(tuple(lead) + (width,))
    if ffn:
        out["blk0.ffn"] = eng.ws("d_Q", rows).cpu().numpy() > 0
        out["blk1.ffn"] = eng.ws("d_K", rows).cpu().numpy() > 0
    return out


def gather_masks(local, rows_of_rank, n_global, dist, device):
    """All ranks' masks assembled into global-batch order.  local: this rank's masks (may be {} for an empty share);
    rows_of_rank: global row indices of this rank's rows; every rank must call with the same keys known: the key list and the
    trailing shapes are taken from rank-independent metadata (HEAD_BN + the FFN keys)."""
    import torch
    out = {}
    T = None
    for k, v in local.items():
        if k.endswith(".ffn"):
            T = v.shape[1]
    t_box = torch.tensor([T or 0], device=device)
    dist.all_reduce(t_box, op=dist.ReduceOp.MAX)
    T = int(t_box.item())
    keys = [(scope, (T, width) if zbuf in ("z1", "z2") else (width,)) for zbuf, _, width, scopes in HEAD_BN for scope in scopes]
    keys += [("blk0.ffn", (T, 40)), ("blk1.ffn", (T, 40))]
    for k, tail in keys:
        full = torch.zeros((n_global,) + tail, dtype=torch.int32, device=device)
        if len(rows_of_rank):
            full[torch.as_tensor(np.asarray(rows_of_rank), device=device)] = torch.as_tensor(local[k].reshape((len(rows_of_rank),) + tail).astype(np.int32), device=device)
        dist.all_reduce(full)
        out[k] = full.cpu().numpy() > 0
    return out
