"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one train step, per kernel and per launch."""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
data = [(x[ki], x[gi], x[bi], float(x[vi].replace(",", ""))) for x in r]
idx = [i for i, d in enumerate(data) if "k_embed_fwd" in d[0]]
# the bench runs its timed steps first and the large-batch HBM microbench afterwards: a step of the bench workload is a window
# between two gathers of the FIRST grid size seen (the 4th such window: past the warm-up), unless --last asks for the last one
same = [i for i in idx if data[i][1] == data[idx[0]][1]] if idx else []
if "--last" not in sys.argv and len(same) >= 2:
    k = min(4, len(same) - 2)
    s, e = same[k], same[k + 1]
elif len(idx) >= 2:
    s, e = idx[-2], idx[-1]
else:                                   # window holds one step start: take a full step's worth of launches around it
    fin = [i for i, d in enumerate(data) if "k_finish_losses" in d[0]]
    s = idx[0] if idx and any(i > idx[0] for i in fin) else (fin[0] + 1 if fin else 0)
    e = min([i for i in fin if i > s] or [len(data) - 1]) + 1
step = data[s:e]
tot = sum(d[3] for d in step)
agg = OrderedDict()
for name, g, b, v in step:
    n = name.split("(")[0].replace("pamrec::", "").replace("void ", "")[:46]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
print(f"one step: {len(step)} launches, {tot / 1000:.1f} us (serialised, cold cache)")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n:46s} x{c:<3d} {v / 1000:8.1f} us  {100 * v / tot:5.1f} %")
if "--all" in sys.argv:
    for name, g, b, v in step:
        print(f"{name[:60]:60s} {g:>16s} {b:>14s} {v / 1000:8.1f}")
