"""BASELINE.json configs[0]: the quick-start path from TEXT FILES (synthetic WeChat-Channels-shaped data), end to end:
tokenise -> native batcher (prefetched) -> pinned staging -> H2D -> train step -> losses D2H, through PAMRECModel.fit_step's
own loop body.  Prints one JSON line: samples/s of the second epoch (the first also parses the file)."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from pamrec_b200 import synth  # noqa: E402
from pamrec_b200.deeprec_utils import prepare_hparams  # noqa: E402
from pamrec_b200.models import PAMRECModel  # noqa: E402
from pamrec_b200.prefetch import Prefetcher  # noqa: E402
from pamrec_b200.sequential_iterator import SequentialIterator  # noqa: E402

n_users = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
tmp = tempfile.mkdtemp(prefix="pamrec_file_e2e_")
d = synth.generate(tmp, "wechat", n_users=n_users, n_items=20000, n_cates=100, mean_len=150, seed=1)
hp = prepare_hparams(os.path.join(ROOT, "pamrec_b200", "config", "mmoe.yaml"), dataset="wechat", bucket_num=10, add_feature=False,
                     embed_l2=1e-6, layer_l2=1e-6, discrepancy_loss_weight=0.1, learning_rate=0.001, epochs=1, EARLY_STOP=5, is_clip_norm=1,
                     batch_size=500, show_step=10 ** 9, MODEL_DIR=os.path.join(tmp, "m/"), SUMMARIES_DIR=os.path.join(tmp, "s/"),
                     user_vocab=os.path.join(d, "user_vocab.pkl"), item_vocab=os.path.join(d, "item_vocab.pkl"),
                     cate_vocab=os.path.join(d, "category_vocab.pkl"), train_num_ngs=0, max_seq_length=100, pairwise_metrics=[],
                     weighted_metrics=["wauc"], fuzhu_weight=0.5, fine_tune=False, eval_step=10 ** 9, noise_train_hist=0,
                     noise_train_listwise=0, noise_only_predict=0, write_tfevents=False)
model = PAMRECModel(hp, SequentialIterator, seed=8)
train = os.path.join(d, "train_data")
res = []
for epoch in range(3):
    t0 = time.perf_counter()
    n = steps = 0
    pending = None
    for feed in Prefetcher(model.iterator.load_data_from_file(train, min_seq_length=1, batch_num_ngs=0)):
        queued = model.train_async(None, feed)                   # one step ahead, like fit_step
        if pending is not None:
            r = pending.result()
        pending = queued
        n += feed["items"].shape[0]
        steps += 1
    r = pending.result()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res.append(dict(epoch=epoch, samples=n, steps=steps, seconds=dt, samples_per_s=n / dt, ms_per_step=1e3 * dt / steps, loss=r[2]))
t0 = time.perf_counter()
ev = model.run_weighted_eval(os.path.join(d, "valid_data"), num_ngs=0)
dt_ev = time.perf_counter() - t0
n_ev = sum(1 for _ in open(os.path.join(d, "valid_data")))
print(json.dumps({"workload": "wechat quick start from text files, B=500, T=100", "n_users": n_users, "epochs": res,
                  "eval": {"impressions": n_ev, "seconds": dt_ev, "impressions_per_s": n_ev / dt_ev, "metrics": {k: float(v) for k, v in ev.items()}}}))
