"""Data-parallel input: with `iterator.shard = (world, rank)` the native batcher draws the same global batches on every rank but
materialises only that rank's rows (`LocalFeed`).  Checked against `dist.split_feed` of the global feed - the selection the model
used before and still uses for feeds it is handed whole - for training (whole groups of 5) and scoring (row by row) passes, and,
over gloo with two processes and a stand-in engine, through PAMRECModel's own loops."""
import os
import random
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gen_golden as G
from pamrec_b200 import dist as D
from pamrec_b200 import sequential_iterator as IT


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_batches_equal_split_feed(world):
    case = list(G.CASES)[0]
    with tempfile.TemporaryDirectory() as tmp:
        data_dir = G.synth_case(case, tmp)
        hp = G.hparams_for(case, data_dir)
        train, valid = os.path.join(data_dir, "train_data"), os.path.join(data_dir, "valid_data")

        def run(shard):
            random.seed(8)
            it = IT.SequentialIterator(hp, None)
            it.shard = shard
            out = [list(it.load_data_from_file(train)) for _ in range(2)]          # two epochs: the RNG state carries over
            out.append(list(it.load_data_from_file(valid, min_seq_length=3)))
            return out
        whole = run(None)
        assert all(type(b) is dict for p in whole for b in p)
        n_empty = 0
        for rank in range(world):
            mine = run((world, rank))
            for p, (gp, lp) in enumerate(zip(whole, mine)):
                assert len(gp) == len(lp) > 0                                      # every rank sees every global batch
                for g, l in zip(gp, lp):
                    want, n = D.split_feed(g, world, rank, grouped=(p < 2))
                    assert isinstance(l, IT.LocalFeed) and (l.global_rows, l.world, l.rank) == (n, world, rank)
                    assert list(l) == list(g)
                    for k in g:
                        assert l[k].dtype == want[k].dtype and l[k].shape == want[k].shape, (p, k)
                        assert np.array_equal(l[k], want[k], equal_nan=True), (p, k)
                    assert np.array_equal(l.global_users, g["users"]) and np.array_equal(l.global_labels_satisfied, g["labels_satisfied"])
                    n_empty += l["items"].shape[0] == 0
        if world == 8:
            assert n_empty > 0, "a tail batch with fewer groups than ranks leaves some ranks without rows (they still step)"
        # negatives are drawn from the whole batch: that path keeps global feeds
        random.seed(8)
        it = IT.SequentialIterator(hp, None)
        it.shard = (world, 0)
        assert type(next(it.load_data_from_file(train, batch_num_ngs=4))) is dict


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(data_dir, tmp, stub):
    from pamrec_b200 import deeprec_utils as DU
    from pamrec_b200 import models as M
    M.Engine = stub
    hp = DU.prepare_hparams(None, model_type="mmoe", dataset="wechat", bucket_num=10, method="classification", loss="cross_entropy_loss",
                            optimizer="adam", item_embedding_dim=16, cate_embedding_dim=4, user_embedding_dim=20, layer_sizes=[100, 64],
                            expert_layer_sizes=[100, 64], gate_layer_sizes=[64, 5], expert_num=5, activation=["relu", "relu"], enable_BN=True,
                            dropout=[0.0, 0.0], embedding_dropout=0.0, hidden_size=40, attention_size=40, att_fcn_layer_sizes=[20, 1],
                            fuzhu_weight=0.5, batch_size=60, max_seq_length=20, epochs=1, eval_step=10 ** 9, show_step=10 ** 9,
                            train_num_ngs=0, need_sample=False, metrics=["auc", "logloss"], pairwise_metrics=["mean_mrr", "group_auc"],
                            weighted_metrics=["wauc"], MODEL_DIR=os.path.join(tmp, "model") + "/", SUMMARIES_DIR=os.path.join(tmp, "s") + "/",
                            noise_train_hist=0, noise_train_listwise=0, noise_only_predict=0, save_model=False,
                            user_vocab=os.path.join(data_dir, "user_vocab.pkl"), item_vocab=os.path.join(data_dir, "item_vocab.pkl"),
                            cate_vocab=os.path.join(data_dir, "category_vocab.pkl"))
    return M.PAMRECModel(hp, IT.SequentialIterator, seed=8)


def _run(model, data_dir, tmp):
    """What both the single process and every rank do: some training steps, the scoring loop, predict."""
    model.engine.quality = 1.0
    seen = []
    for k, feed in enumerate(model.iterator.load_data_from_file(os.path.join(data_dir, "train_data"))):
        model.train(None, feed)
        seen.append(model.engine.last_upload)
        if k == 5:
            break
    res = model.run_weighted_eval(os.path.join(data_dir, "valid_data"), num_ngs=3)
    out = os.path.join(tmp, f"pred_{model.engine.rank}.txt")
    model.predict(os.path.join(data_dir, "test_data"), out)
    return seen, res, (open(out).read() if os.path.exists(out) else None)


def _make_stub():
    from test_host_loops import StubEngine

    class DistStub(StubEngine):
        def upload(self, feed, training=True, staged=False, global_batch=0):
            db = super().upload(feed, training, staged, global_batch)
            self.last_upload = (int(global_batch), np.asarray(feed["items"]).copy(), np.asarray(feed["item_history"]).copy())
            return db

        def all_reduce_(self, t):
            if self.world > 1:
                dist.all_reduce(t)
            return t

        def _gather_table(self, shard, vocab_rows):
            return D.unshard_table([shard.numpy()] * self.world, vocab_rows)
    return DistStub


def _worker(rank, world, port, data_dir, tmp, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model(data_dir, tmp, _make_stub())
        assert model.engine.world == world and model.iterator.shard == (world, rank)
        q.put((rank, _run(model, data_dir, tmp)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_run_the_model_loops_on_their_share(tmp_path, lib_built):
    from pamrec_b200 import synth
    data_dir = synth.generate(str(tmp_path), "wechat", n_users=80, n_items=300, n_cates=12, mean_len=40, seed=3, n_neg=3)
    solo = _model(data_dir, str(tmp_path), _make_stub())
    assert solo.engine.world == 1 and solo.iterator.shard is None
    seen1, res1, pred1 = _run(solo, data_dir, str(tmp_path))
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, data_dir, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        seen, res, pred = got[rank]
        assert res == res1                                        # every rank computes the metrics of ALL rows
        assert (pred == pred1) if rank == 0 else (pred is None)   # rank 0 writes the prediction file
        assert len(seen) == len(seen1)
        for (gb, items, hist), (_, items1, hist1) in zip(seen, seen1):
            rows = D.group_rows(len(items1), world, rank)
            assert gb == len(items1) and np.array_equal(items, items1[rows]) and np.array_equal(hist, hist1[rows])


@pytest.mark.parametrize("shard", [None, (3, 1)])
def test_eval_batches_are_the_same_with_one_or_several_batcher_threads(shard, monkeypatch, lib_built):
    """The native batcher fills the history arrays of a large evaluation batch with several threads over disjoint rows
    (csrc/batcher.cu, PAMREC_BATCHER_THREADS): every one of the 19 arrays is identical to the single-threaded result."""
    from pamrec_b200 import synth
    case = list(G.CASES)[0]
    with tempfile.TemporaryDirectory() as tmp:
        data_dir = G.synth_case(case, tmp)
        hp = G.hparams_for(case, data_dir)
        hp.batch_size = 1500                                                       # above the 512-row threshold of the threaded path
        path = os.path.join(tmp, "big_eval")
        T = hp.max_seq_length
        n_users, n_items, n_cates = (len(IT.load_dict(p)) for p in (hp.user_vocab, hp.item_vocab, hp.cate_vocab))
        synth.write_eval_file(path, 40, 99, T, n_users, n_items, n_cates, seed=3)   # 4 000 lines: 2 full batches + a tail

        def run(threads):
            monkeypatch.setenv("PAMREC_BATCHER_THREADS", str(threads))
            it = IT.SequentialIterator(hp, None)
            it.shard = shard
            return list(it.load_data_from_file(path, min_seq_length=1))
        one, four = run(1), run(4)
        assert len(one) == len(four) == 3
        for a, b in zip(one, four):
            assert list(a) == list(b)
            for k in a:
                assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k], equal_nan=True), k
