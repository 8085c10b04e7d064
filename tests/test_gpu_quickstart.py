"""The reference's driver flow (example/00_quick_start/sequential.py:435-522) end to end on the GPU, from text files:
fit_step (iterator -> train steps -> periodic run_weighted_eval -> checkpoint on improvement) -> latest_checkpoint ->
load_model -> run_weighted_eval -> predict, through the compat/ import paths the reference driver uses.  The scores of the
trained model are then re-computed by the fp64 oracle from the checkpointed variables."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def data_root(tmp_path_factory):
    from pamrec_b200 import synth
    root = tmp_path_factory.mktemp("quickstart")
    synth.generate(str(root), "wechat", n_users=400, n_items=3000, n_cates=40, mean_len=60, seed=7, eval_per_user=2)
    return str(root)


@pytest.mark.parametrize("train_num_ngs", [0, 4])
def test_driver_runs_as_subprocess(data_root, tmp_path, train_num_ngs):
    """The repo's driver with the reference's flag names (compat/example/00_quick_start/sequential.py); train_num_ngs = 4 is
    the 1 positive + 4 in-batch negatives layout (every row of a batch of 20 becomes a listwise group of 5)."""
    drv = os.path.join(ROOT, "compat", "example", "00_quick_start", "sequential.py")
    cmd = [sys.executable, drv, "--dataset", "wechat", "--data_path", data_root, "--epochs", "1",
           "--batch_size", "100" if train_num_ngs == 0 else "20", "--train_num_ngs", str(train_num_ngs),
           "--eval_step", "5", "--show_step", "5", "--save_path", str(tmp_path / "ranking"), "--write_prediction_to_file"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=os.path.dirname(drv))
    assert r.returncode == 0, r.stdout[-3000:]
    assert "Time cost for training" in r.stdout and "'auc'" in r.stdout and "wauc" in r.stdout, r.stdout[-2000:]
    out = os.path.join(data_root, "wechat", "output.txt")
    preds = np.loadtxt(out)
    n_test = sum(1 for _ in open(os.path.join(data_root, "wechat", "test_data")))
    assert preds.shape == (n_test,) and np.isfinite(preds).all() and (preds > 0).all() and (preds < 1).all()


def test_fit_checkpoint_eval_match_oracle(data_root, tmp_path):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    from reco_utils.recommender.deeprec.deeprec_utils import prepare_hparams
    from reco_utils.recommender.deeprec.io.sequential_iterator import SequentialIterator
    from reco_utils.recommender.deeprec.models.sequential.pamrec import PAMRECModel
    import tensorflow.compat.v1 as tf                      # compat shim: latest_checkpoint only
    from oracle import pamrec_oracle as O                   # checker
    d = os.path.join(data_root, "wechat")
    model_dir = str(tmp_path / "model") + "/"
    hp = prepare_hparams(os.path.join(ROOT, "pamrec_b200", "config", "mmoe.yaml"), dataset="wechat", bucket_num=10, add_feature=False,
                         embed_l2=1e-6, layer_l2=1e-6, discrepancy_loss_weight=0.1, learning_rate=0.001, epochs=1, EARLY_STOP=5,
                         is_clip_norm=1, batch_size=100, show_step=10 ** 9, MODEL_DIR=model_dir, SUMMARIES_DIR=str(tmp_path / "s") + "/",
                         user_vocab=os.path.join(d, "user_vocab.pkl"), item_vocab=os.path.join(d, "item_vocab.pkl"),
                         cate_vocab=os.path.join(d, "category_vocab.pkl"), train_num_ngs=0, max_seq_length=50, pairwise_metrics=[],
                         weighted_metrics=["wauc", "wmrr", "wndcg@2;4", "whit@2;4"], fuzhu_weight=0.5, fine_tune=False, eval_step=4,
                         noise_train_hist=0, noise_train_listwise=0, noise_only_predict=0, write_tfevents=False)
    model = PAMRECModel(hp, SequentialIterator, seed=8)
    valid, test = os.path.join(d, "valid_data"), os.path.join(d, "test_data")
    assert model.fit_step(os.path.join(d, "train_data"), valid, valid_num_ngs=0, eval_metric="auc") is model
    ckpt = tf.train.latest_checkpoint(model_dir)
    assert ckpt and os.path.exists(ckpt + ".safetensors")
    final = model.run_weighted_eval(test, num_ngs=0)
    for k in ("auc", "logloss", "wauc", "wmrr", "wndcg@2", "whit@4"):
        assert k in final and np.isfinite(final[k]), (k, final)
    # a fresh model restored from the checkpoint reproduces the checkpointed model's metrics exactly
    model.load_model(ckpt)
    at_ckpt = model.run_weighted_eval(test, num_ngs=0)
    fresh = PAMRECModel(hp, SequentialIterator, seed=123)
    fresh.load_model(ckpt)
    assert fresh.run_weighted_eval(test, num_ngs=0) == at_ckpt
    # the oracle, fed the checkpointed variables, scores the same impressions to 1e-4 (north_star: metrics within 1e-4)
    var = fresh.engine.get_variables()
    nu, ni, nc, T, _ = fresh.engine.dims
    om = O.OracleModel(nu, ni, nc, T, seed=1)
    for n in om.params:
        om.params[n] = torch.as_tensor(var[n], dtype=om.params[n].dtype).reshape(om.params[n].shape)
    for n in om.bn_state:
        om.bn_state[n] = torch.as_tensor(var[n], dtype=om.bn_state[n].dtype).reshape(om.bn_state[n].shape)
    got, want = [], []
    for feed in fresh.iterator.load_data_from_file(test, min_seq_length=fresh.min_seq_length, batch_num_ngs=0):
        if not feed:
            continue
        got.append(fresh.eval(None, feed)[0].reshape(-1))
        f2 = {k: np.asarray(v) for k, v in feed.items()}
        f2["mask"] = f2["mask"].astype(np.int32)
        f2["users"] = f2["users"].astype(np.int32)
        want.append(om.eval_forward(f2).t["pred"].numpy().reshape(-1))
    got, want = np.concatenate(got), np.concatenate(want)
    assert got.shape == want.shape and np.abs(got - want).max() <= 1e-4, float(np.abs(got - want).max())


def _worst_difference(va, vb):
    """Largest |difference| over the variables two runs of the same steps can be compared on, with the variable's name.
    A bias in front of a batch norm (the `b_nn_layer*` of every BN MLP) has an exactly-zero data gradient: what reaches Adam is
    the fp32 rounding noise of a sum whose order the atomics change from run to run, and Adam normalises it to steps of
    +-learning_rate - so those biases random-walk differently in ANY two runs and are left out, and with
    them the BN moving means, which track the batch mean of `h W + b` and so inherit the walk (moving variances do not)."""
    worst, where = 0.0, None
    for n in va:
        leaf = n.rsplit("/", 1)[-1]
        if leaf.startswith("b_nn_layer") or leaf == "moving_mean":
            continue
        d = float(np.abs(va[n] - vb[n]).max())
        if d > worst:
            worst, where = d, n
    return worst, where


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["safetensors", "tf"])
def test_resume_from_a_checkpoint_with_optimizer_state(data_root, tmp_path, fmt):
    """hparams.save_optimizer (an extension: the reference's Saver never holds Adam slots, BM:61-63) stores m, v and the step
    under `optimizer/`; a fresh model restored from it continues like the uninterrupted run, in either file format."""
    import random
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    from reco_utils.recommender.deeprec.deeprec_utils import prepare_hparams
    from reco_utils.recommender.deeprec.io.sequential_iterator import SequentialIterator
    from reco_utils.recommender.deeprec.models.sequential.pamrec import PAMRECModel
    d = os.path.join(data_root, "wechat")
    model_dir = str(tmp_path / "model") + "/"

    def make(seed, **kw):
        hp = prepare_hparams(os.path.join(ROOT, "pamrec_b200", "config", "mmoe.yaml"), dataset="wechat", bucket_num=10, add_feature=False,
                             embed_l2=1e-6, layer_l2=1e-6, discrepancy_loss_weight=0.1, learning_rate=0.001, epochs=1, is_clip_norm=1,
                             batch_size=100, show_step=10 ** 9, MODEL_DIR=model_dir, SUMMARIES_DIR=str(tmp_path / "s") + "/",
                             user_vocab=os.path.join(d, "user_vocab.pkl"), item_vocab=os.path.join(d, "item_vocab.pkl"),
                             cate_vocab=os.path.join(d, "category_vocab.pkl"), train_num_ngs=0, max_seq_length=50, pairwise_metrics=[],
                             weighted_metrics=["wauc"], fuzhu_weight=0.5, fine_tune=False, noise_train_hist=0, noise_train_listwise=0,
                             noise_only_predict=0, write_tfevents=False, **kw)
        return PAMRECModel(hp, SequentialIterator, seed=seed)
    a = make(8)
    random.seed(5)
    feeds = []
    for f in a.iterator.load_data_from_file(os.path.join(d, "train_data")):
        feeds.append(f)
        if len(feeds) == 6:
            break
    assert len(feeds) == 6
    for f in feeds:
        a.train(None, f)
    want = a.engine.get_variables()
    b = make(8, save_optimizer=True, checkpoint_format=fmt)
    for f in feeds[:3]:
        b.train(None, f)
    path = b.saver.save(save_path=model_dir + "mid")
    c = make(99)                                                  # other initial values, no optimizer history
    c.load_model(path)
    assert c.engine.step == 3
    for f in feeds[3:]:
        c.train(None, f)
    got = c.engine.get_variables()
    worst, where = _worst_difference(got, want)
    assert worst <= 5e-5, (worst, where)                          # fp32 atomics reorder sums between runs; a lost Adam state is ~1e-3
    # the same checkpoint without the optimizer key-space (what the reference's Saver writes) restarts Adam from zero
    from pamrec_b200 import checkpoint as CK
    variables, opt = CK.load(path)
    assert opt is not None and int(opt["step"]) == 3
    CK.save(model_dir + "weights_only", variables, fmt=fmt)
    e = make(99)
    e.load_model(model_dir + "weights_only")
    for f in feeds[3:]:
        e.train(None, f)
    cold = e.engine.get_variables()
    assert _worst_difference(cold, want)[0] > 1e-4


@pytest.mark.gpu
def test_train_async_one_step_ahead_equals_blocking_train(data_root):
    """PAMRECModel.train_async (what fit_step and the bench's e2e loop drive) against train() on a second model with the same
    initial values: same losses step by step (fp32 atomics reorder sums between runs, nothing more), same variables at the end."""
    import random
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    from reco_utils.recommender.deeprec.deeprec_utils import prepare_hparams
    from reco_utils.recommender.deeprec.io.sequential_iterator import SequentialIterator
    from reco_utils.recommender.deeprec.models.sequential.pamrec import PAMRECModel
    d = os.path.join(data_root, "wechat")
    hp = prepare_hparams(os.path.join(ROOT, "pamrec_b200", "config", "mmoe.yaml"), dataset="wechat", bucket_num=10, add_feature=False,
                         embed_l2=1e-6, layer_l2=1e-6, discrepancy_loss_weight=0.1, learning_rate=0.001, epochs=1, is_clip_norm=1,
                         batch_size=100, show_step=10 ** 9, save_model=False, user_vocab=os.path.join(d, "user_vocab.pkl"),
                         item_vocab=os.path.join(d, "item_vocab.pkl"), cate_vocab=os.path.join(d, "category_vocab.pkl"), train_num_ngs=0,
                         max_seq_length=50, pairwise_metrics=[], weighted_metrics=["wauc"], fuzhu_weight=0.5, fine_tune=False,
                         noise_train_hist=0, noise_train_listwise=0, noise_only_predict=0, write_tfevents=False)
    a, b = PAMRECModel(hp, SequentialIterator, seed=8), PAMRECModel(hp, SequentialIterator, seed=8)
    random.seed(5)
    feeds = []
    for f in a.iterator.load_data_from_file(os.path.join(d, "train_data")):
        feeds.append(f)
        if len(feeds) == 9:
            break
    blocking = [a.train(None, f) for f in feeds]
    handles, ahead = [], []
    for f in feeds:                                              # up to 9 steps queued before the first result is read: the
        handles.append(b.train_async(None, f))                   # 4-slot loss ring must hand every step its own losses
    ahead = [h.result() for h in handles]
    assert handles[0].result() is not None and len(ahead) == 9
    for r0, r1 in zip(blocking, ahead):
        assert r0[:2] == r1[:2] == [None, None]
        assert np.allclose(r0[2:7], r1[2:7], rtol=2e-6, atol=1e-7), (r0, r1)
    va, vb = a.engine.get_variables(), b.engine.get_variables()
    worst, where = _worst_difference(va, vb)
    # The two runs launch the same kernels; what differs is the order of the fp32 atomics in the weight-gradient sums (~1e-7 of the
    # gradient).  Adam divides by sqrt(v): on entries whose gradient is itself that small the noise decides the direction of a step of
    # size lr, so nine steps can open a gap of a few percent of lr = 1e-3 on single entries (observed 2e-5 .. 9e-5 over the runs of this
    # round); anything systematic - a step applied twice, a stale batch, a missed update - shows up at >= lr.
    assert worst <= 0.2 * 1e-3, (worst, where)
