"""Runs the stand-alone HBM kernels of bench.py's microbench once or twice and nothing else: the command profiled by ncu."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pamrec_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1 << 18)
ap.add_argument("--items", type=int, default=10_000_000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--zipf", type=float, default=0.0)
a = ap.parse_args()
ni, nc, T = a.items, 100_000, 50
dev = torch.device("cuda:0")
eng = Engine(1000, ni, nc, T, 64).allocate("cuda:0")
eng.pool["item_w"].normal_(0, 0.01)
g = torch.Generator(device="cpu").manual_seed(1)
if a.zipf > 0:
    import numpy as np
    r = np.random.default_rng(1).zipf(a.zipf, size=a.rows * T)
    ih = torch.from_numpy(((r - 1) % ni).astype("int32")).to(dev)
else:
    ih = torch.randint(0, ni, (a.rows * T,), generator=g, dtype=torch.int32).to(dev)
ch = torch.randint(0, nc, (a.rows * T,), generator=g, dtype=torch.int32).to(dev)
ti = torch.randint(0, ni, (a.rows,), generator=g, dtype=torch.int32).to(dev)
tc = torch.randint(0, nc, (a.rows,), generator=g, dtype=torch.int32).to(dev)
out = torch.empty(a.rows * T * 40, dtype=torch.float32, device=dev)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for i in range(a.reps):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    eng._check(eng.lib.pamrec_bench_gather(eng.handle, C.c_void_p(ih.data_ptr()), C.c_void_p(ch.data_ptr()), C.c_void_p(ti.data_ptr()),
                                           C.c_void_p(tc.data_ptr()), a.rows, T, C.c_void_p(out.data_ptr()), st))
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print(f"gather {a.rows * T} lookups {ms:.3f} ms  {248 * a.rows * T / ms / 1e6:.0f} GB/s algorithmic")
for i in range(a.reps):
    eng._check(eng.lib.pamrec_bench_table_adam(eng.handle, i + 1, st))
torch.cuda.synchronize()
print("done")
