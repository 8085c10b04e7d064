"""The tcgen05 / tensor-memory attention forward kernel (csrc/kernels_attn_tc.cu) against the FFMA kernel it replaces and,
through the stage-wise tensors, against the fp64 oracle: same Q, K, V, mask in; y = softmax(QK^T/sqrt(40))V + qin and the saved
row statistics out.  Shapes cover the three tile geometries: two samples per 128-row tile (T <= 64), one tile per sample
(T <= 128) and two query tiles per sample (T <= 208)."""
import os

import numpy as np
import pytest
import torch

from oracle import pamrec_oracle as O

pytestmark = pytest.mark.gpu


def _engine(om, dims, B, attn):
    from pamrec_b200.engine import Engine
    old = os.environ.get("PAMREC_ATTN")
    os.environ["PAMREC_ATTN"] = attn
    try:
        eng = Engine(*dims, B).allocate()
    finally:
        if old is None:
            os.environ.pop("PAMREC_ATTN", None)
        else:
            os.environ["PAMREC_ATTN"] = old
    eng.set_variables({n: t.numpy() for n, t in om.params.items()})
    eng.set_variables({n: t.numpy() for n, t in om.bn_state.items()})
    return eng


@pytest.mark.parametrize("T,B", [(12, 20), (50, 35), (64, 10), (100, 15), (128, 5), (200, 10), (50, 1025)])
def test_tc_attention_forward_matches_ffma_and_oracle(T, B):
    nu, ni, nc = 200, 2000, 40
    om = O.OracleModel(nu, ni, nc, T, seed=21)
    O.perturb_params(om.params, om.bn_state, seed=22)
    om.proj = "grouped"
    batch = O.make_batch(7, B, T, nu, ni, nc)
    if B >= 20:                                            # a fully masked history would be a different code path: make one
        batch["mask"][5:10] = 0
    ref = om.train_step(batch, apply=False)["t"]
    a = _engine(om, (nu, ni, nc, T), B, "ffma")
    b = _engine(om, (nu, ni, nc, T), B, "tc")
    for eng in (a, b):
        eng.forward(eng.upload(batch), training=True, want_pred=False)
    torch.cuda.synchronize()
    for k in range(2):
        ya, yb = a.ws(f"blk{k}.y", B).cpu().numpy(), b.ws(f"blk{k}.y", B).cpu().numpy()
        yo = ref[f"blk{k}.y"].detach().numpy()
        scale = np.abs(yo).max()
        assert np.isfinite(yb).all(), f"blk{k}.y has non-finite values"
        print(f"T={T} B={B} blk{k}.y: tc vs ffma {np.abs(ya - yb).max() / scale:.2e}, tc vs oracle {np.abs(yb - yo).max() / scale:.2e}, "
              f"ffma vs oracle {np.abs(ya - yo).max() / scale:.2e}")
        assert np.abs(yb - yo).max() <= 1e-5 * scale
        ma, mb = a.ws(f"blk{k}.ml", B).cpu().numpy(), b.ws(f"blk{k}.ml", B).cpu().numpy()
        # (m, l) pairs may differ in how the maximum was found; what the backward pass uses is m + log(l)
        la, lb = ma[..., 0] + np.log(ma[..., 1]), mb[..., 0] + np.log(mb[..., 1])
        assert np.abs(la - lb).max() <= 1e-5 * max(np.abs(la).max(), 1.0)
    la, lb = a.ws("logits", B).cpu().numpy(), b.ws("logits", B).cpu().numpy()
    assert np.abs(la - lb).max() <= 1e-5 * np.abs(la).max()
    a.close(); b.close()
