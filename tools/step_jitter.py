"""Per-step device times of the bench loop, with and without the nvidia-smi clock sampler running (diagnostic for outliers)."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from pamrec_b200 import synth  # noqa: E402

w = bench.WORKLOADS["takatak_b1025_t50"]
model = bench.build_model(w, tempfile.mkdtemp())
eng = model.engine
feeds = [synth.array_batch(1000 + 17 * i, w["B"], w["T"], w["n_users"], w["n_items"], w["n_cates"]) for i in range(8)]
res = [eng.upload(f) for f in feeds]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(n, tag):
    torch.cuda.synchronize()
    evs = []
    for i in range(n):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.train_step(res[i % 8])
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    print(tag, "mean %.3f median %.3f max %.3f" % (sum(ms) / n, sorted(ms)[n // 2], max(ms)), " ".join("%.2f" % x for x in ms[:40]), flush=True)


for i in range(16):
    eng.train_step(res[i % 8])
run(40, "no sampler a")
run(40, "no sampler b")
cs = bench.ClockSampler(0).start()
import time
time.sleep(0.5)
run(40, "sampler a  ")
run(40, "sampler b  ")
run(40, "sampler c  ")
cs.stop()
run(40, "stopped    ")
