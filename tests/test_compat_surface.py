"""The reference's import paths (what example/00_quick_start/sequential.py imports and calls) resolve
through `compat/` to this package, with the reference's names."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_paths_and_names(monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(ROOT, "compat"))
    for name in [m for m in sys.modules if m == "tensorflow" or m.startswith(("tensorflow.", "reco_utils"))]:
        monkeypatch.delitem(sys.modules, name)
    pam = importlib.import_module("reco_utils.recommender.deeprec.models.sequential.pamrec")
    it = importlib.import_module("reco_utils.recommender.deeprec.io.sequential_iterator")
    du = importlib.import_module("reco_utils.recommender.deeprec.deeprec_utils")
    tf = importlib.import_module("tensorflow.compat.v1")
    for method in ("fit_step", "fit", "run_weighted_eval", "run_eval", "predict", "load_model", "train", "eval", "eval_with_user",
                   "infer", "step_train", "batch_train"):
        assert callable(getattr(pam.PAMRECModel, method)), method
    for method in ("parser_one_line", "load_data_from_file", "_convert_data", "gen_feed_dict"):      # io/iterator.py:9-24
        assert callable(getattr(it.SequentialIterator, method)), method
    assert it.lisan(0.5, "wechat") == 3 and len(it.bar_border_list) == 10 and set(it.takatak_bar_border_list_dict) == {10, 8, 6}
    for fn in ("prepare_hparams", "load_dict", "cal_metric", "cal_weighted_metric", "mrr_score", "ndcg_score", "dcg_score", "hit_score"):
        assert callable(getattr(du, fn)), fn
    assert callable(tf.disable_v2_behavior) and callable(tf.train.latest_checkpoint) and isinstance(tf.__version__, str)
    assert tf.train.latest_checkpoint(os.path.join(ROOT, "tests")) is None        # no `checkpoint` state file there
    import pamrec_b200.models as M
    assert pam.PAMRECModel is M.PAMRECModel
