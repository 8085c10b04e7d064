"""Host logic of the data-parallel / row-sharded step (pamrec_b200/dist.py), including a world_size-2 run over
``gloo`` that routes lookups to their owners exactly as csrc/api.cu:shard_exchange_fwd does on NCCL and checks the
result against a plain full-table lookup.  No GPU involved."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pamrec_b200 import dist as D


def test_shard_roundtrip_and_padding():
    rng = np.random.default_rng(0)
    for vocab in (1, 7, 64, 1001):
        full = rng.normal(size=(vocab, 16)).astype(np.float32)
        for world in (1, 2, 3, 8):
            shards = [D.shard_table(full, world, r) for r in range(world)]
            assert all(s.shape == (D.shard_rows(vocab, world), 16) for s in shards)
            assert np.array_equal(D.unshard_table(shards, vocab), full)
            ids = np.arange(vocab)
            for i in ids[:: max(vocab // 13, 1)]:
                assert np.array_equal(shards[D.owner_of(i, world)][D.local_row(i, world)], full[i])


def test_group_split_is_a_partition_of_whole_groups():
    for n_groups in (1, 2, 5, 41, 205):
        n = n_groups * 5
        for world in (1, 2, 4, 8):
            rows = [D.group_rows(n, world, r) for r in range(world)]
            assert sorted(np.concatenate(rows).tolist()) == list(range(n))
            for r in rows:
                assert len(r) % 5 == 0
                g = r.reshape(-1, 5)
                assert (g[:, 0] % 5 == 0).all() and (np.diff(g, axis=1) == 1).all()      # groups stay intact
    with pytest.raises(ValueError):
        D.group_rows(7, 2, 0)


def test_split_feed_slices_every_batch_array():
    B, T = 20, 6
    feed = {"items": np.arange(B), "cates": np.arange(B) + 100, "users": np.arange(B), "mask": np.ones((B, T), np.int32),
            "item_history": np.arange(B * T).reshape(B, T), "labels_satisfied": np.zeros((B, 1), np.float32), "scalar": 3}
    seen = []
    for r in range(3):
        loc, n = D.split_feed(feed, 3, r)
        assert n == B and loc["scalar"] == 3
        assert loc["item_history"].shape[0] == loc["items"].shape[0] == loc["labels_satisfied"].shape[0]
        assert np.array_equal(loc["item_history"][:, 0], loc["items"] * T)
        seen += loc["items"].tolist()
    assert sorted(seen) == list(range(B))
    loc, _ = D.split_feed(feed, 3, 1, grouped=False)
    assert loc["items"].tolist() == list(range(1, B, 3))


def test_exchange_plan_properties():
    rng = np.random.default_rng(1)
    vocab, world = 1000, 4
    rps = D.shard_rows(vocab, world)
    ids = rng.integers(0, vocab, size=5000)
    per_owner, inv = D.exchange_plan(ids, world, rps)
    flat = np.concatenate([np.asarray(rows, np.int64) * world + o for o, rows in enumerate(per_owner)])
    assert np.array_equal(flat[inv], ids)                       # every lookup finds its row
    assert len(np.unique(flat)) == len(flat) == len(np.unique(ids))
    for rows in per_owner:
        assert (np.diff(rows) > 0).all() and (rows < rps).all()  # sorted, in range


def test_table_placement_rule():
    assert D.choose_tables(None, 1, 50000, 30000, 50) == "local"
    assert D.choose_tables("auto", 8, 50000, 30000, 50) == "replicated"            # configs[1]: 10 MB of tables
    assert D.choose_tables("auto", 8, 50000, 10_000_000, 100_000) == "sharded"     # configs[2]: 650 MB
    assert D.choose_tables("replicated", 1, 10, 10, 10) == "local"
    assert D.choose_tables("sharded", 1, 10, 10, 10) == "sharded"                  # the exchange path also runs on one GPU
    with pytest.raises(ValueError):
        D.choose_tables("local", 2, 10, 10, 10)
    with pytest.raises(ValueError):
        D.choose_tables("mirrored", 2, 10, 10, 10)


def test_two_shot_slices_cover_the_padded_contribution():
    for n in (1, 63, 64, 1000, 252_197 + 560_250):
        for world in (2, 3, 8):
            cap, sl = D.two_shot_slices(n, world)
            assert cap >= n and cap % (64 * world) == 0 and cap - n < 64 * world
            assert sl[0][0] == 0 and sl[-1][1] == cap
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            assert all((hi - lo) % 4 == 0 for lo, hi in sl)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        vocab, width = 997, 16
        full = rng.normal(size=(vocab, width)).astype(np.float32)          # same on every rank
        shard = D.shard_table(full, world, rank)
        rps = D.shard_rows(vocab, world)
        ids = np.random.default_rng(100 + rank).integers(0, vocab, size=(15 + 10 * rank, 12))   # ragged share per rank
        per_owner, inv = D.exchange_plan(ids, world, rps)
        # ids -> owners (what the counts + ids all-to-all deliver)
        box = [None] * world
        dist.all_gather_object(box, per_owner)
        asked = [box[src][rank] for src in range(world)]                    # rows every src wants from me
        served = [shard[a] for a in asked]                                  # owner gather
        box2 = [None] * world
        dist.all_gather_object(box2, served)
        rows = np.concatenate([box2[o][rank] for o in range(world)])       # owner-major == order of my unique list
        got = rows[inv].reshape(ids.shape + (width,))
        ok_lookup = np.array_equal(got, full[ids])
        # gradient direction: per-unique-row sums go back to the owners, owners merge duplicates across ranks
        g = np.random.default_rng(200 + rank).normal(size=(ids.size, width))
        nun = sum(len(a) for a in per_owner)
        acc = np.zeros((nun, width))
        np.add.at(acc, inv, g)
        off = np.cumsum([0] + [len(a) for a in per_owner])
        box3 = [None] * world
        dist.all_gather_object(box3, [acc[off[o]:off[o + 1]] for o in range(world)])
        mine = np.zeros((rps, width))
        for src in range(world):
            np.add.at(mine, asked[src], box3[src][rank])
        # check against the dense gradient of the whole job
        box4 = [None] * world
        dist.all_gather_object(box4, (ids.reshape(-1), g))
        dense = np.zeros((vocab, width))
        for i, gg in box4:
            np.add.at(dense, i, gg)
        ok_grad = np.allclose(D.shard_table(dense, world, rank), mine, rtol=1e-12, atol=1e-12)
        # batch-norm style statistics: fp64 sums all-reduced == sums over the concatenated batch
        z = torch.from_numpy(np.random.default_rng(300 + rank).normal(size=(10 + rank, 4)))
        s = torch.stack([z.sum(0), (z * z).sum(0), torch.full((4,), float(z.shape[0]), dtype=torch.float64)])
        dist.all_reduce(s)
        box5 = [None] * world
        dist.all_gather_object(box5, z.numpy())
        zz = np.concatenate(box5)
        ok_bn = np.allclose(s[0].numpy() / s[2].numpy(), zz.mean(0)) and np.allclose(s[1].numpy() / s[2].numpy(), (zz * zz).mean(0))
        # replicated tables: every rank's dense contribution (+ touch marks) summed by the two-shot scheme of k_xr_* - rank r
        # sums slice r of all contributions in rank order, everybody receives every slice - equals the whole job's gradient,
        # and the rows touched by ANY rank are the unique ids of the global batch (the L2 rows of tf.unique)
        contrib, touch = D.replicated_contribution(ids, g, vocab)
        flat = np.concatenate([contrib.reshape(-1), touch])
        cap, slices = D.two_shot_slices(flat.size, world)
        xbuf = np.zeros(cap)
        xbuf[:flat.size] = flat
        box6 = [None] * world
        dist.all_gather_object(box6, xbuf)                                  # "peer memory": every rank can read every contribution
        lo, hi = slices[rank]
        mine_sum = np.zeros(hi - lo)
        for src in range(world):
            mine_sum += box6[src][lo:hi]
        box7 = [None] * world
        dist.all_gather_object(box7, mine_sum)                              # every rank stores its slice into every result buffer
        rbuf = np.concatenate(box7)
        ok_rep = np.allclose(rbuf[:contrib.size].reshape(vocab, width), dense, rtol=1e-12, atol=1e-12)
        all_ids = np.unique(np.concatenate([i for i, _ in box4]))
        ok_rep = ok_rep and np.array_equal(np.nonzero(rbuf[contrib.size:flat.size] > 0)[0], all_ids)
        q.put((rank, bool(ok_lookup), bool(ok_grad), bool(ok_bn), bool(ok_rep)))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    for rank, ok_lookup, ok_grad, ok_bn, ok_rep in res:
        assert ok_rep, f"rank {rank}: replicated-table contributions summed by the two-shot scheme differ from the job's gradient"
        assert ok_lookup, f"rank {rank}: sharded lookup differs from the full-table lookup"
        assert ok_grad, f"rank {rank}: merged row gradients differ from the dense gradient"
        assert ok_bn, f"rank {rank}: all-reduced statistics differ"
