# coding=utf-8
"""Quick start for PAMREC on the B200-native engine.  Flag names and defaults follow the reference driver
(example/00_quick_start/sequential.py:49-91); run from this directory:

    python sequential.py --dataset wechat --data_path /path/to/data
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "..", ".."), os.path.join(HERE, "..", "..", "..")]

from absl import app, flags  # noqa: E402

import tensorflow.compat.v1 as tf  # noqa: E402  (shim)
from reco_utils.recommender.deeprec.deeprec_utils import prepare_hparams  # noqa: E402
from reco_utils.recommender.deeprec.io.sequential_iterator import SequentialIterator  # noqa: E402
from reco_utils.recommender.deeprec.models.sequential.pamrec import PAMRECModel  # noqa: E402
from reco_utils.recommender.deeprec.models.sequential.mmoe import MMoEModel_original  # noqa: E402
from reco_utils.recommender.deeprec.models.sequential.ple import PLEModel  # noqa: E402
from reco_utils.recommender.deeprec.models.sequential.sharebottom import ShareBottomModel  # noqa: E402
from reco_utils.recommender.deeprec.models.sequential.sasrec import SASRecModel  # noqa: E402

FLAGS = flags.FLAGS
flags.DEFINE_string("dataset", "wechat", "Dataset name.")
flags.DEFINE_string("eval_metric", "auc", "metric to eval")
flags.DEFINE_integer("val_num_ngs", 0, "negatives per positive in valid_data")
flags.DEFINE_integer("test_num_ngs", 0, "negatives per positive in test_data")
flags.DEFINE_integer("batch_size", 500, "Batch size.")
flags.DEFINE_string("save_path", "ranking", "Save path.")
flags.DEFINE_string("name", "try", "Experiment name.")
flags.DEFINE_string("model", "PAMREC", "Model name: PAMREC, MMOE_ORIGINAL, PLE, SHAREBOTTOM or SASREC.")
flags.DEFINE_boolean("only_test", False, "Only test and do not train.")
flags.DEFINE_boolean("write_prediction_to_file", False, "Whether to write prediction to file.")
flags.DEFINE_integer("is_clip_norm", 1, "Whether to clip gradient norm.")
flags.DEFINE_integer("epochs", 10, "Number of epochs.")
flags.DEFINE_integer("early_stop", 5, "Patience for early stop.")
flags.DEFINE_string("data_path", os.path.join("..", "..", "tests", "resources", "deeprec", "sequential"), "Data file path.")
flags.DEFINE_integer("train_num_ngs", 0, "negatives per positive for training")
flags.DEFINE_float("embed_l2", 1e-6, "L2 regulation for embeddings.")
flags.DEFINE_float("layer_l2", 1e-6, "L2 regulation for layers.")
flags.DEFINE_float("discrepancy_loss_weight", 0.1, "weight of the order (ApproxNDCG) loss")
flags.DEFINE_float("learning_rate", 0.001, "Learning rate.")
flags.DEFINE_integer("show_step", 500, "Step for showing metrics.")
flags.DEFINE_integer("bucket_num", 10, "number of play-ratio buckets")
flags.DEFINE_boolean("add_feature", False, "add time feature")
flags.DEFINE_float("fuzhu_weight", 0.5, "weight of the auxiliary loss")
flags.DEFINE_integer("eval_step", 2500, "Step for evaluation.")
flags.DEFINE_float("noise_train_hist", 0, "noise_train_hist")
flags.DEFINE_float("noise_train_listwise", 0, "noise_train_listwise")
flags.DEFINE_float("noise_only_predict", 0, "noise_only_predict")
flags.DEFINE_string("sparse_adam", "dense_exact", "dense_exact (tf.train.AdamOptimizer semantics) or lazy")
flags.DEFINE_string("checkpoint_format", "safetensors", "safetensors, tf (TensorFlow tensor bundle) or npz")
flags.DEFINE_boolean("save_optimizer", False, "also checkpoint the Adam state (resume exactly; the reference's Saver does not)")
flags.DEFINE_string("loss", "cross_entropy_loss", "cross_entropy_loss or softmax (over groups of train_num_ngs + 1 rows)")


def get_model(f, model_path, summary_path, user_vocab, item_vocab, cate_vocab):
    # model name -> (class, yaml) as in the reference driver (example/00_quick_start/sequential.py:113-296)
    models = {"PAMREC": (PAMRECModel, "mmoe.yaml"), "MMOE_ORIGINAL": (MMoEModel_original, "mmoe.yaml"), "PLE": (PLEModel, "ple.yaml"),
              "SHAREBOTTOM": (ShareBottomModel, "sharebottom.yaml"), "SASREC": (SASRecModel, "sasrec.yaml")}
    if f.model not in models:
        raise NotImplementedError("--model must be one of " + ", ".join(models))
    cls, yaml_name = models[f.model]
    weighted = {"wechat": ["wauc", "wmrr", "wndcg@2;4;6;8;10", "whit@2;4;6;8;10"],
                "takatak": ["wauc", "wmrr", "wndcg@10", "whit@10", "wmrr@10"]}[f.dataset]
    yaml_file = os.path.join(HERE, "..", "..", "reco_utils", "recommender", "deeprec", "config", yaml_name)
    hparams = prepare_hparams(
        yaml_file, dataset=f.dataset, bucket_num=f.bucket_num, add_feature=f.add_feature, embed_l2=f.embed_l2,
        layer_l2=f.layer_l2, discrepancy_loss_weight=f.discrepancy_loss_weight, learning_rate=f.learning_rate,
        epochs=f.epochs, EARLY_STOP=f.early_stop, is_clip_norm=f.is_clip_norm, batch_size=f.batch_size,
        show_step=f.show_step, MODEL_DIR=model_path, SUMMARIES_DIR=summary_path, user_vocab=user_vocab,
        item_vocab=item_vocab, cate_vocab=cate_vocab, train_num_ngs=f.train_num_ngs, max_seq_length=100,
        pairwise_metrics=[], weighted_metrics=weighted, fuzhu_weight=f.fuzhu_weight, fine_tune=False,
        eval_step=f.eval_step, noise_train_hist=f.noise_train_hist, noise_train_listwise=f.noise_train_listwise,
        noise_only_predict=f.noise_only_predict, sparse_adam=f.sparse_adam, checkpoint_format=f.checkpoint_format,
        save_optimizer=f.save_optimizer, loss=f.loss)
    return cls(hparams, SequentialIterator, seed=8)


def main(argv):
    f = FLAGS
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # torchrun --nproc-per-node N sequential.py ...: one process per GPU, every rank reads the same files and trains on
        # its listwise groups of every batch (pamrec_b200/dist.py); no counterpart in the reference (single device)
        import torch
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    chief = int(os.environ.get("RANK", "0")) == 0
    data_path = os.path.join(f.data_path, f.dataset)
    train_file, valid_file, test_file = (os.path.join(data_path, n) for n in ("train_data", "valid_data", "test_data"))
    vocabs = [os.path.join(data_path, n) for n in ("user_vocab.pkl", "item_vocab.pkl", "category_vocab.pkl")]
    save_path = os.path.join(f.save_path, f.model, f.name)
    model_path, summary_path = os.path.join(save_path, "model/"), os.path.join(save_path, "summary/")
    model = get_model(f, model_path, summary_path, *vocabs)
    if f.only_test:
        model.load_model(tf.train.latest_checkpoint(model_path))
        res = model.run_weighted_eval(test_file, num_ngs=f.test_num_ngs)
        if chief:
            print(res)
        return
    t0 = time.time()
    model = model.fit_step(train_file, valid_file, valid_num_ngs=f.val_num_ngs, eval_metric=f.eval_metric)
    if chief:
        print("Time cost for training is {0:.2f} mins".format((time.time() - t0) / 60.0))
    ckpt = tf.train.latest_checkpoint(model_path)
    if chief:
        print(ckpt)
    if ckpt:
        model.load_model(ckpt)
    res = model.run_weighted_eval(test_file, num_ngs=f.test_num_ngs)
    if chief:
        print(f.name)
        print("TEST_METRICS", {k: float(v) for k, v in res.items()})
    if f.write_prediction_to_file:
        model.predict(test_file, os.path.join(data_path, "output.txt"))


if __name__ == "__main__":
    app.run(main)
