"""One rank of the multi-GPU parity check (launched by tests/test_gpu_sharded.py through torch.distributed.run).

Every rank builds the SAME oracle and the SAME global batches from seeds, trains on its own share (listwise groups dealt
round-robin, tables row-sharded), and rank 0 compares losses, the gathered tables, dense variables, BN statistics and a
data-parallel scoring pass with the fp64 oracle run on the whole global batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import pamrec_oracle as O  # noqa: E402  (checker only)
from pamrec_b200 import dist as D  # noqa: E402
from pamrec_b200.engine import Engine  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nu, ni, nc, T = 301, 3001, 53, 50
    Bg = 5 * (8 * world + 3)                    # groups do not divide evenly: ranks get different shares
    steps = 3
    om = O.OracleModel(nu, ni, nc, T, seed=3)
    O.perturb_params(om.params, om.bn_state, seed=4)
    cap = -(-(Bg // 5) // world) * 5
    eng = Engine(nu, ni, nc, T, max(cap, 40), world_size=world, rank=rank).allocate(f"cuda:{local}")
    eng.init_comm()
    eng.set_variables({n: t.numpy() for n, t in om.params.items()})
    eng.set_variables({n: t.numpy() for n, t in om.bn_state.items()})
    worst = {}

    def note(name, got, ref, atol):
        # Adam normalises each coordinate's gradient: summation-order noise on a nearly cancelling gradient moves a weight by a
        # fraction of lr (1e-3), so variables are compared on the scale of one optimiser step, not of the weight.
        # (up to 2 lr where a ~0 gradient changes sign): 99.9 % of the entries within `atol`, none further than one step.
        got = np.asarray(got, np.float64)
        ref = np.asarray(ref, np.float64).reshape(got.shape)
        d = np.abs(got - ref)
        err = float(np.quantile(d, 0.999))
        worst[name] = max(worst.get(name, 0.0), err)
        assert np.isfinite(got).all() and err <= atol and d.max() <= 1e-3, f"{name}: q999 {err:.3e} max {d.max():.3e} > {atol:.1e}"

    for step in range(steps):
        batch = O.make_batch(100 + step, Bg, T, nu, ni, nc)
        ref = om.train_step(batch)
        mine, n = D.split_feed(batch, world, rank)
        got = eng.train_step(eng.upload(mine, global_batch=n)).cpu().numpy()
        tol = 2e-5 if step == 0 else 1e-4            # same weights: 2e-5; later steps: trajectories drift (see test_gpu_parity)
        for i, k in enumerate(("loss", "data_loss", "regular_loss", "auxiliary_data_loss", "order_loss")):
            r = ref["losses"][k]
            assert abs(got[i] - r) <= tol * max(abs(r), 1e-3), (rank, step, k, float(got[i]), r)
    # variables after the steps (collective gathers)
    var = eng.get_variables()
    for name in ("item_embedding", "cate_embedding", "user_long_embedding", "user_short_embedding"):
        note(name, var["sequential/embedding/" + name], om.params["sequential/embedding/" + name].numpy(), 5e-5)   # 5 % of one lr step over 3 steps
    for name in om.params:
        if name in var and "embedding/" not in name:
            ref_v = om.params[name].numpy()
            leaf = name.rsplit("/", 1)[-1]
            if leaf.startswith("b_nn_layer"):
                continue                                # bias in front of BN: zero data gradient, Adam follows fp32 noise
            note(name, var[name], ref_v, 1e-4)
    for name, t in om.bn_state.items():
        got_bn = var[name]
        atol = 2e-4 if name.endswith("moving_mean") else 1e-6
        assert np.allclose(got_bn, t.numpy(), rtol=1e-4, atol=atol), name
    # a step in which the last rank has nothing to train on (B = 0): it still serves rows and joins the all-reduces
    small = O.make_batch(500, 5 * (world - 1), T, nu, ni, nc) if world > 1 else O.make_batch(500, 5, T, nu, ni, nc)
    ref = om.train_step(small)
    mine, n = D.split_feed(small, world, rank)
    got = eng.train_step(eng.upload(mine, global_batch=n)).cpu().numpy()
    assert abs(got[0] - ref["losses"]["loss"]) <= 1e-4 * abs(ref["losses"]["loss"]), (rank, float(got[0]), ref["losses"]["loss"])
    # (batch norm over so few rows is ill-conditioned, so after this step only a loose bound on the tables is meaningful)
    item = eng.get_variables()["sequential/embedding/item_embedding"]
    assert np.abs(item - om.params["sequential/embedding/item_embedding"].numpy()).max() <= 1e-3
    # replicated dense parameters must be bit-identical on every rank
    dp = eng.pool["dense_param"].clone()
    lo, hi = dp.clone(), dp.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "dense parameters diverged between ranks"
    # data-parallel scoring
    ev = O.make_batch(999, 37, T, nu, ni, nc, grouped=False)
    want = om.eval_forward(ev).t["pred"].numpy().reshape(-1)
    mine, n = D.split_feed(ev, world, rank, grouped=False)
    full = torch.zeros(n, dtype=torch.float32, device=eng.device)
    full[rank::world] = eng.forward(eng.upload(mine, training=False, global_batch=n), training=False)
    pred = eng.all_reduce_(full).cpu().numpy()
    assert np.abs(pred - want).max() <= 1e-4, float(np.abs(pred - want).max())
    dist.barrier()
    if rank == 0:
        top = sorted(worst.items(), key=lambda kv: -kv[1])[:5]
        print("DIST_PARITY_OK world=%d" % world, " ".join(f"{k.rsplit('/', 1)[-1]}={v:.1e}" for k, v in top), flush=True)
    eng.close()
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
