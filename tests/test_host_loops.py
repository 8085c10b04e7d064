"""The host side of PAMRECModel - fit_step / fit / run_weighted_eval / run_eval / predict / load_model / the Saver - driven on
CPU with a stand-in for the device: `StubEngine` keeps the real Engine's inventory, variables and checkpoint plumbing (host
memory instead of device memory) and replaces the three device calls by cheap deterministic functions of the feed and of one
weight.  Nothing here measures or checks kernels; those tests are marked `gpu`."""
import os
import random
import types

import numpy as np
import pytest
import torch

from pamrec_b200 import checkpoint as CK
from pamrec_b200 import deeprec_utils as DU
from pamrec_b200 import engine as E
from pamrec_b200 import models as M
from pamrec_b200 import synth
from pamrec_b200.sequential_iterator import SequentialIterator

W = "sequential/logit_fcn/nn_part/b_nn_output"        # the one weight the stub's scores depend on


class StubEngine(E.Engine):
    def allocate(self, device="cpu"):
        self.device = torch.device("cpu")
        nu, ni, nc, T, B = self.dims
        z = lambda *s: torch.zeros(*s, dtype=torch.float32)
        self.pool = {k: z(self.dense_numel) for k in ("dense_param", "dense_grad", "dense_m", "dense_v")}
        self.pool["bn_moving"] = z(self.bn_numel)
        for pre, rows, w in (("item", ni, 16), ("cate", nc, 4), ("ulong", nu, 20), ("ushort", nu, 20)):
            for s in ("w", "m", "v"):
                self.pool[f"{pre}_{s}"] = z(self.table_rows(rows), w)
        self.calls = dict(train=0, forward=0)
        return self

    def init_comm(self):
        return self

    def upload(self, feed, training=True, staged=False, global_batch=0):
        assert feed["items"].shape[0] <= self.dims[4], "the model sized the engine for the largest batch it feeds"
        return types.SimpleNamespace(feed=feed, batch=int(feed["items"].shape[0]))

    def train_step(self, db, losses_out=None):
        self.step += 1
        self.calls["train"] += 1
        self.dense(W).add_(0.01)                               # "training": scores drift with the step count
        self.dense(W, "dense_m").add_(1.0)
        base = 1.0 / self.step
        return torch.tensor([4 * base, base, base, base, base])

    def train_step_async(self, db):
        losses = self.train_step(db).numpy()
        return types.SimpleNamespace(result=lambda: losses)

    def to_host_async(self, dev_tensor):
        value = dev_tensor.numpy().copy()
        return types.SimpleNamespace(result=lambda: value)

    def forward(self, db, training=False, want_pred=True):
        self.calls["forward"] += 1
        f = db.feed
        # a score that knows something about the label (so that metrics move) and depends on the weight
        x = (f["items"] % 7).astype(np.float32) / 7 + self.quality * np.asarray(f["labels_satisfied"]).reshape(-1) + float(self.dense(W)[0])
        return torch.sigmoid(torch.from_numpy(x.astype(np.float32)))

    quality = 0.0


@pytest.fixture()
def setup(tmp_path, monkeypatch, lib_built):
    monkeypatch.setattr(M, "Engine", StubEngine)
    d = synth.generate(str(tmp_path), "wechat", n_users=80, n_items=300, n_cates=12, mean_len=40, seed=3, n_neg=3)

    def make(**kw):
        base = dict(model_type="mmoe", dataset="wechat", bucket_num=10, method="classification", loss="cross_entropy_loss", optimizer="adam",
                    item_embedding_dim=16, cate_embedding_dim=4, user_embedding_dim=20, layer_sizes=[100, 64], expert_layer_sizes=[100, 64],
                    gate_layer_sizes=[64, 5], expert_num=5, activation=["relu", "relu"], enable_BN=True, dropout=[0.0, 0.0],
                    embedding_dropout=0.0, hidden_size=40, attention_size=40, att_fcn_layer_sizes=[20, 1], fuzhu_weight=0.5,
                    batch_size=60, max_seq_length=20, epochs=2, eval_step=10, EARLY_STOP=2, show_step=10 ** 9, train_num_ngs=0,
                    need_sample=False, metrics=["auc", "logloss"], pairwise_metrics=["mean_mrr", "ndcg@2", "hit@2", "group_auc"],
                    weighted_metrics=["wauc"], MODEL_DIR=str(tmp_path / "model") + "/", SUMMARIES_DIR=str(tmp_path / "sum") + "/",
                    noise_train_hist=0, noise_train_listwise=0, noise_only_predict=0,
                    user_vocab=os.path.join(d, "user_vocab.pkl"), item_vocab=os.path.join(d, "item_vocab.pkl"),
                    cate_vocab=os.path.join(d, "category_vocab.pkl"))
        base.update(kw)
        return M.PAMRECModel(DU.prepare_hparams(None, **base), SequentialIterator, seed=8)
    return types.SimpleNamespace(d=d, make=make, train=os.path.join(d, "train_data"), valid=os.path.join(d, "valid_data"),
                                 test=os.path.join(d, "test_data"), model_dir=str(tmp_path / "model"))


def test_fit_step_evaluates_checkpoints_and_stops_early(setup, capsys):
    m = setup.make(epochs=50)
    evals = []
    real = m.run_weighted_eval

    def scripted(filename, num_ngs, **kw):                       # metric improves twice, then stalls -> early stop after 2 more evals
        res = real(filename, num_ngs, **kw)
        res["group_auc"] = [0.5, 0.6, 0.55, 0.58, 0.7][len(evals)]
        evals.append(m.engine.step)
        return res
    m.run_weighted_eval = scripted
    assert m.fit_step(setup.train, setup.valid, valid_num_ngs=3, eval_metric="group_auc") is m
    assert evals == [10, 20, 30, 40]                             # SBM:316-349: stop when step - best_step >= EARLY_STOP * eval_step
    assert m.best_step == 20 and m.engine.calls["train"] == 40
    out = capsys.readouterr().out
    assert "early stop at epoch" in out and "best step: 20" in out
    # checkpoints only on improvement (SBM:322-333), newest last in the index file
    assert sorted(f for f in os.listdir(setup.model_dir)) == ["checkpoint", "step_10.safetensors", "step_20.safetensors"]
    assert M.latest_checkpoint(setup.model_dir) == os.path.join(setup.model_dir, "step_20")
    w_now = float(m.engine.dense(W)[0])
    m.load_model(M.latest_checkpoint(setup.model_dir))
    assert abs(float(m.engine.dense(W)[0]) - 0.2) < 1e-6 and abs(w_now - 0.4) < 1e-5
    assert float(m.engine.dense(W, "dense_m")[0]) == 40.0        # the reference's Saver holds no optimizer slots: untouched
    with pytest.raises(IOError, match="Failed to find any matching files for"):
        m.load_model(os.path.join(setup.model_dir, "step_30"))


def test_saver_rotation_formats_and_optimizer_state(setup):
    m = setup.make(epochs=2, checkpoint_format="tf", save_optimizer=True)        # max_to_keep = epochs (BM:62)
    feed = next(m.iterator.load_data_from_file(setup.train))
    for k in range(3):
        m.train(None, feed)
        m.saver.save(save_path=os.path.join(setup.model_dir, f"epoch_{k}"))
    names = sorted(os.listdir(setup.model_dir))
    assert names == ["checkpoint", "epoch_1.data-00000-of-00001", "epoch_1.index", "epoch_2.data-00000-of-00001", "epoch_2.index"]
    text = open(os.path.join(setup.model_dir, "checkpoint")).read().splitlines()
    assert text == ['model_checkpoint_path: "epoch_2"', 'all_model_checkpoint_paths: "epoch_1"', 'all_model_checkpoint_paths: "epoch_2"']
    variables, opt = CK.load(os.path.join(setup.model_dir, "epoch_1"))
    assert set(variables) == set(m.engine.variable_shapes()) and int(opt["step"]) == 2 and opt[W + "/Adam"][0] == 2.0
    # a fresh model resumes exactly where epoch_1 was written; a checkpoint of another graph is refused with the reference's error
    fresh = setup.make(epochs=2)
    fresh.load_model(os.path.join(setup.model_dir, "epoch_1"))
    assert fresh.engine.step == 2 and float(fresh.engine.dense(W, "dense_m")[0]) == 2.0
    for name, val in m.engine.get_variables().items():
        if name != W:
            assert np.array_equal(fresh.engine.get_variables()[name], val), name
    variables["sequential/not_in_this_graph"] = np.zeros(3, np.float32)
    CK.save(os.path.join(setup.model_dir, "alien"), variables)
    with pytest.raises(IOError):
        fresh.load_model(os.path.join(setup.model_dir, "alien"))
    variables.pop("sequential/not_in_this_graph")
    variables[W] = np.zeros(5, np.float32)
    CK.save(os.path.join(setup.model_dir, "misshapen"), variables)
    with pytest.raises(IOError):
        fresh.load_model(os.path.join(setup.model_dir, "misshapen"))
    with pytest.raises(ValueError, match="checkpoint_format"):
        setup.make(checkpoint_format="hdf5")


def test_run_weighted_eval_filters_like_the_reference(setup):
    m = setup.make()
    m.engine.quality = 1.5
    res = m.run_weighted_eval(setup.valid, num_ngs=3)
    assert set(res) == {"auc", "logloss", "mean_mrr", "ndcg@2", "hit@2", "group_auc", "wauc"}
    # recompute from the raw scores: groups of 1 + 3 lines, groups without a positive dropped for the pairwise metrics
    # (SBM:456-464), users whose labels are all 0 or all 1 dropped for the point and weighted metrics (SBM:466-486)
    users, preds, labels = [], [], []
    for feed in m.iterator.load_data_from_file(setup.valid):
        u, p, l = m.eval_with_user(None, feed)
        users += u.tolist(); preds += p.reshape(-1).tolist(); labels += l.reshape(-1).tolist()
    users, preds, labels = np.asarray(users), np.asarray(preds), np.asarray(labels)
    gp, gl = preds.reshape(-1, 4), labels.reshape(-1, 4)
    keep_g = gl.sum(1) != 0
    assert 0 < keep_g.sum() < len(gl), "the synthetic file must hold both kinds of groups"
    want = DU.cal_metric(list(gl[keep_g]), list(gp[keep_g]), ["mean_mrr", "ndcg@2", "hit@2", "group_auc"])
    mixed = np.asarray([0 < labels[users == u].sum() < (users == u).sum() for u in users])
    assert 0 < mixed.sum() <= len(mixed)
    want.update(DU.cal_metric(labels[mixed].tolist(), preds[mixed].tolist(), ["auc", "logloss"]))
    want.update(DU.cal_weighted_metric(users[mixed].tolist(), preds[mixed].tolist(), labels[mixed].tolist(), ["wauc"]))
    assert res == want
    assert res["group_auc"] > 0.9 and res["auc"] > 0.7          # the stub's scores know the label
    with pytest.raises(NotImplementedError):
        m.run_weighted_eval(setup.valid, num_ngs=3, calc_mean_alpha=True)
    plain = m.run_eval(setup.valid, num_ngs=3)                   # SBM:380-413: no filtering at all
    assert set(plain) == {"auc", "logloss", "mean_mrr", "ndcg@2", "hit@2", "group_auc"}
    assert plain["auc"] == DU.cal_metric(labels.tolist(), preds.tolist(), ["auc"])["auc"]


def test_predict_writes_one_score_per_line_and_fit_runs_epochs(setup, tmp_path, capsys):
    m = setup.make(epochs=3, EARLY_STOP=1)
    out = str(tmp_path / "scores.txt")
    assert m.predict(setup.test, out) is m
    lines = open(out).read().split("\n")
    n = sum(1 for _ in open(setup.test))
    assert len(lines) == n + 1 and lines[-1] == ""
    want = np.concatenate([m.infer(None, f)[0].reshape(-1) for f in m.iterator.load_data_from_file(setup.test)])
    assert [float(x) for x in lines[:-1]] == [float(str(v)) for v in want]
    with pytest.raises(ValueError, match="negative numbers for training"):
        m.fit(setup.train, setup.valid, valid_num_ngs=3)         # need_sample False and train_num_ngs 0 (SBM:149-152)
    m.need_sample = True
    with pytest.raises(ValueError, match="negative numbers for validation"):
        m.fit(setup.train, setup.valid, valid_num_ngs=0)
    random.seed(1)
    m2 = setup.make(epochs=3, EARLY_STOP=1, need_sample=True, train_num_ngs=0)
    seq = iter([0.6, 0.5, 0.9])
    real = m2.run_weighted_eval
    m2.run_weighted_eval = lambda f, n, **kw: {**real(f, n, **kw), "group_auc": next(seq)}
    m2.fit(setup.train, setup.valid, valid_num_ngs=3)
    text = capsys.readouterr().out
    assert m2.best_epoch == 1 and "early stop at epoch 2!" in text       # SBM:196-206
    assert m2.train_num_ngs == 1                                         # SBM:153-154 forces one negative when sampling is "needed"
