"""A few full train steps at the batch of bench.py's sparse microbench (B = 65535, T = 50, 10 M items, Zipf ids): the command
profiled by ncu for the kernels of the sparse backward (kernels_sparse2.cu)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pamrec_b200 import synth  # noqa: E402
from pamrec_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="dense_exact")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=65535)
ap.add_argument("--min-len", type=int, default=50)
a = ap.parse_args()
ni, nc, T = 10_000_000, 100_000, 50
eng = Engine(50000, ni, nc, T, a.batch, sparse_adam=a.mode).allocate("cuda:0")
eng.pool["item_w"].normal_(0, 0.01)
eng.pool["cate_w"].normal_(0, 0.01)
eng.pool["dense_param"].normal_(0, 0.05)
db = eng.upload(synth.array_batch(77, a.batch, T, 50000, ni, nc, zipf_a=1.05, min_len=a.min_len))
eng.profile(True)
for _ in range(a.steps):
    eng.train_step(db)
tab = eng.profile_table()
for k in ("sparse_plan", "sparse_walk", "sparse_adam", "embed_fwd", "embed_bwd_reduce"):
    print(k, round(tab[k][0] / a.steps, 4), "ms")
print("unique", eng.ws("sp.nuniq").cpu().numpy()[:3])
