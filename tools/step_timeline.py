"""Timeline of one replayed step graph (CUPTI kernel records through torch.profiler): idle gaps between consecutive kernels.
Tells whether the difference between the step time and the sum of kernel durations is spread over every kernel boundary or
sits in a few places (stream joins, copies).   python tools/step_timeline.py [--model small --batch 256]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import ofb_b200  # noqa: F401
from ofb_b200.engine import SearchStepEngine

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="small")
ap.add_argument("--batch", type=int, default=256)
a = ap.parse_args()
D, H = {"tiny": (192, 3), "small": (384, 6), "base": (768, 12)}[a.model]
eng = SearchStepEngine(D, H, 12, a.batch, drop_path_rate=0.1)
eng.init_params(seed=0)
eng.set_schedule(0.0)
img = torch.randn(a.batch, 3, 224, 224, device="cuda")
lab = torch.randint(0, 1000, (a.batch,), device="cuda")
for _ in range(4):
    eng.step_graphed(img, lab)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        eng.step_graphed(img, lab)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if getattr(e, "device_type", None) is not None and "CUDA" in str(e.device_type)]
ev = sorted(ev, key=lambda e: e.time_range.start)
# the middle replay: kernels between the 2nd and 3rd hyper upload (ofb::copy_f32_kernel)
starts = [i for i, e in enumerate(ev) if "copy_f32" in e.name]
lo, hi = starts[1], starts[2]
step = ev[lo:hi]
t0, t1 = step[0].time_range.start, ev[hi].time_range.start
busy_end = step[0].time_range.start
gaps = []
for e in step:
    if e.time_range.start > busy_end:
        gaps.append((e.time_range.start - busy_end, prev, e.name[:60]))
    if e.time_range.end > busy_end:
        busy_end, prev = e.time_range.end, e.name[:60]
total_gap = sum(g[0] for g in gaps)
print(f"step {t1 - t0:.0f} us, {len(step)} device activities, idle {total_gap:.0f} us in {len(gaps)} gaps "
      f"(median {sorted(g[0] for g in gaps)[len(gaps) // 2]:.2f} us)")
hist = {}
for g in gaps:
    k = "<1" if g[0] < 1 else "1-2" if g[0] < 2 else "2-4" if g[0] < 4 else "4-8" if g[0] < 8 else ">8"
    hist[k] = hist.get(k, 0) + 1
print("gap histogram (us):", hist)
for g in sorted(gaps, reverse=True)[:12]:
    print(f"  {g[0]:7.2f} us  after {g[1]}  ->  {g[2]}")
# gap by kind of boundary
by = {}
for g in gaps:
    key = (g[1].split("<")[0].split("(")[0][-28:], g[2].split("<")[0].split("(")[0][-28:])
    v = by.setdefault(key, [0.0, 0])
    v[0] += g[0]; v[1] += 1
for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:12]:
    print(f"  {v[0]:7.1f} us over {v[1]:3d} boundaries ({v[0] / v[1]:.2f} each): {k[0]} -> {k[1]}")
