"""Runs a few train steps of one workload and nothing else: the command profiled by ncu (B200_PROFILING.md)."""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from pamrec_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="takatak_b1025_t50")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--eval", action="store_true")
a = ap.parse_args()
w = bench.WORKLOADS[a.workload]
model = bench.build_model(w, tempfile.mkdtemp())
eng = model.engine
feeds = [synth.array_batch(1000 + i, w["B"], w["T"], w["n_users"], w["n_items"], w["n_cates"]) for i in range(2)]
dbs = [eng.upload(f) for f in feeds]
for i in range(a.steps):
    if a.eval:
        eng.forward(dbs[i % 2], training=False)
    else:
        eng.train_step(dbs[i % 2])
torch.cuda.synchronize()
print("done", a.workload, a.steps)
