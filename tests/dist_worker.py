"""One rank of the multi-GPU parity check (launched by tests/test_gpu_sharded.py through torch.distributed.run).

Every rank builds the SAME oracle and the SAME global batches from seeds, trains on its own share (listwise groups dealt
round-robin, tables row-sharded), and compares with the fp64 oracle run on the whole global batch:

* step 0 (same weights) against the fp64 oracle: losses, the all-reduced dense gradients, the global clip norms of the
  sparse tables, Adam's first moment of every table row (= the owner-side merge of the row gradients of all ranks) - Adam is
  invariant to the scale of a gradient, so these are checked directly rather than through the weights;
* every step against a SINGLE-GPU engine that rank 0 runs on the whole global batch (same kernels, same fp32 arithmetic up
  to summation order, so the two trajectories stay together where fp32-vs-fp64 ones drift): losses per step, then tables,
  dense variables and BN moving statistics after the last step;
* replicated dense parameters bit-identical on every rank; a step in which the last rank has NO rows; data-parallel scoring.

ReLU kinks.  The gradient is discontinuous where a pre-activation crosses 0, and every step has a few of the ~10^6 such values
inside the fp32 rounding of the forward pass.  The oracle therefore differentiates on the ENGINE's activation pattern: every
rank reads its masks back after the forward pass (tests/relu_masks.py), the ranks assemble the global pattern, and the fp64
step is taken on that piece of the piecewise-linear function.  No batch is ever skipped or retried."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import pamrec_oracle as O  # noqa: E402  (checker only)
from pamrec_b200 import _lib as L  # noqa: E402
from pamrec_b200 import dist as D  # noqa: E402
from pamrec_b200.engine import Engine  # noqa: E402
from relu_masks import engine_relu_masks, gather_masks  # noqa: E402

EMB = "sequential/embedding/"
NAMES = ("loss", "data_loss", "regular_loss", "auxiliary_data_loss", "order_loss")
P_ = "sequential/pamrec/"

def attempt(k, rank, world, local):
    """One full comparison on seed set k; returns a summary line."""
    nu, ni, nc, T = 301, 3001, 53, 50
    Bg = 5 * (8 * world + 3)                    # groups do not divide evenly: ranks get different shares
    om = O.OracleModel(nu, ni, nc, T, seed=3 + k)
    O.perturb_params(om.params, om.bn_state, seed=4 + k)
    cap = -(-(Bg // 5) // world) * 5
    tables = os.environ.get("PAMREC_TEST_TABLES", "sharded")       # "sharded" | "replicated"
    eng = Engine(nu, ni, nc, T, max(cap, 40), world_size=world, rank=rank, tables=tables).allocate(f"cuda:{local}")
    eng.init_comm()
    eng.set_variables({n: t.numpy() for n, t in om.params.items()})
    eng.set_variables({n: t.numpy() for n, t in om.bn_state.items()})
    dev = eng.device
    worst = {}
    solo = None
    if rank == 0:                                    # the same model on ONE GPU, fed the whole global batch
        solo = Engine(nu, ni, nc, T, Bg, tables="sharded" if tables == "sharded" else "local").allocate(f"cuda:{local}")
        solo.set_variables({n: t.numpy() for n, t in om.params.items()})
        solo.set_variables({n: t.numpy() for n, t in om.bn_state.items()})

    def note(name, got, ref, atol):
        # Adam normalises each coordinate's gradient: rounding noise on a nearly cancelling gradient moves a weight by a fraction
        # of lr (1e-3), up to 2 lr per step where a ~0 gradient changes sign.  Variables are compared on that scale:
        # 99.9 % of the entries within `atol`, none further than one step per step taken.
        got = np.asarray(got, np.float64)
        ref = np.asarray(ref, np.float64).reshape(got.shape)
        d = np.abs(got - ref)
        err = float(np.quantile(d, 0.999))
        worst[name] = max(worst.get(name, 0.0), err)
        assert np.isfinite(got).all() and err <= atol and d.max() <= 3e-3, f"{name}: q999 {err:.3e} max {d.max():.3e} > {atol:.1e}"

    forced = []

    def one_step(batch, step):
        """forward / masks / backward / apply through the phase entry points; returns (losses, ref, P0)."""
        P0 = eng.pool["dense_param"].clone()
        mine, n = D.split_feed(batch, world, rank)
        db = eng.upload(mine, global_batch=n)
        eng.set_debug(L.DEBUG_SAVE_FFN_HIDDEN)
        eng.forward(db, training=True, want_pred=False)
        rows = D.group_rows(n, world, rank)
        masks = gather_masks(engine_relu_masks(eng, len(rows)), rows, n, dist, dev)
        eng.set_debug(0)
        ref = om.train_step(batch, relu_masks=masks)           # same global step on every rank, on the engine's ReLU pattern
        forced.append(ref["relu_forced"])
        eng.backward(db)
        got = eng.apply_gradients(db).cpu().numpy()
        if solo is not None:
            one = solo.train_step(solo.upload(batch)).cpu().numpy()
            for i, name in enumerate(NAMES):
                assert abs(got[i] - one[i]) <= 1e-5 * max(abs(one[i]), 1e-3), ("vs single GPU", step, name, float(got[i]), float(one[i]))
        return got, ref, P0

    def finish(result):
        eng.close()
        if solo is not None:
            solo.close()
        return result

    for step in range(3):
        got, ref, P0 = one_step(O.make_batch(100 + step + 1000 * k, Bg, T, nu, ni, nc), step)
        tol = 2e-5 if step == 0 else 1e-3            # same weights: 2e-5; later steps vs the ORACLE are a sanity bound only
        for i, name in enumerate(NAMES):
            r = ref["losses"][name]
            assert abs(got[i] - r) <= tol * max(abs(r), 1e-3), (rank, step, name, float(got[i]), r)
        if step == 0:
            l2 = om.hp["layer_l2"]
            gmax = max(float(ref["grads"][nm].abs().max()) for nm in eng.info[L.POOL_DENSE])
            for nm, d in eng.info[L.POOL_DENSE].items():
                g = eng.dense(nm, "dense_grad").cpu().numpy().astype(np.float64)
                if d["flags"] & L.SEG_L2:
                    o, cnt = d["offset"], d["numel"]
                    g = g + l2 * P0[o:o + cnt].view(d["shape"]).cpu().numpy().astype(np.float64)
                gr = ref["grads"][nm].numpy().reshape(g.shape)
                err = np.abs(g - gr).max()
                assert np.isfinite(g).all() and err <= 1e-4 * np.abs(gr).max() + 2e-6 * gmax, ("dense grad", nm, err, float(np.abs(gr).max()))
            spn = eng.ws("sp_normsq").cpu().numpy()
            for i, nm in enumerate(("item_embedding", "cate_embedding", "user_long_embedding", "user_short_embedding", "position_embedding")):
                want = ref["sqnorms"][EMB + nm]
                assert abs(spn[i] - want) <= 2e-4 * want + 1e-30, ("clip norm", nm, float(spn[i]), want)
            st1 = eng.get_optimizer_state()
            for nm in ("item_embedding", "cate_embedding", "user_long_embedding", "user_short_embedding"):
                a = st1[EMB + nm + "/Adam"].astype(np.float64)
                b_ = om.m[EMB + nm].numpy().astype(np.float64)
                d = np.abs(a - b_)
                bad = np.nonzero(d.max(1) > 2e-4 * np.abs(b_).max())[0]
                assert bad.size == 0, (nm, "first moment after one step", float(d.max()), float(np.abs(b_).max()), "bad rows per owner",
                                       [int((bad % world == r).sum()) for r in range(world)], bad[:12].tolist())
    # variables after the steps (collective gathers) against the single-GPU run
    var = eng.get_variables()
    if solo is not None:
        one = solo.get_variables()
        for name in one:
            if name.rsplit("/", 1)[-1].startswith("b_nn_layer"):
                continue                                # bias in front of BN: zero data gradient, Adam follows fp32 noise
            if name.endswith("moving_mean") or name.endswith("moving_variance"):
                assert np.allclose(var[name], one[name], rtol=1e-5, atol=2e-5), name
            else:
                note(name, var[name], one[name], 2e-5)  # 2 % of one lr step
    # replicated dense parameters must be bit-identical on every rank
    dp = eng.pool["dense_param"].clone()
    lo, hi = dp.clone(), dp.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "dense parameters diverged between ranks"
    # a step in which the last rank has nothing to train on (B = 0): it still serves rows and joins the all-reduces
    # (batch norm over so few rows is ill-conditioned, so after this step only a loose bound on the tables is meaningful)
    got, ref, _ = one_step(O.make_batch(500 + 1000 * k, 5 * (world - 1) if world > 1 else 5, T, nu, ni, nc), 3)
    assert abs(got[0] - ref["losses"]["loss"]) <= 1e-3 * abs(ref["losses"]["loss"]), (rank, float(got[0]), ref["losses"]["loss"])
    item = eng.get_variables()[EMB + "item_embedding"]
    if solo is not None:
        assert np.abs(item - solo.get_variables()[EMB + "item_embedding"]).max() <= 3e-3
    # data-parallel scoring
    ev = O.make_batch(999, 37, T, nu, ni, nc, grouped=False)
    want = om.eval_forward(ev).t["pred"].numpy().reshape(-1)
    mine, n = D.split_feed(ev, world, rank, grouped=False)
    full = torch.zeros(n, dtype=torch.float32, device=dev)
    full[rank::world] = eng.forward(eng.upload(mine, training=False, global_batch=n), training=False)
    pred = eng.all_reduce_(full).cpu().numpy()
    assert np.abs(pred - want).max() <= 1e-4, float(np.abs(pred - want).max())
    dist.barrier()
    top = sorted(worst.items(), key=lambda kv: -kv[1])[:5]
    return finish(f"relu_units_forced_per_step={forced} " + " ".join(f"{n_.rsplit('/', 1)[-1]}={v:.1e}" for n_, v in top))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    summary = attempt(0, rank, world, local)
    if rank == 0:
        print(f"DIST_PARITY_OK world={world} {summary}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException as e:                      # one greppable line per failing rank, then the traceback
        import traceback
        print(f"DIST_PARITY_FAIL rank={os.environ.get('RANK')} {type(e).__name__}: {str(e)[:400]}", flush=True)
        traceback.print_exc()
        sys.stdout.flush()
        os._exit(1)
