"""In-situ, warm, per-kernel breakdown of the search step: CUDA-event pairs around every library launch of host-launched
steps (no profiler, real clocks, activations resident as in the real step).

  python tools/step_breakdown.py [--model small] [--batch 256] [--depth 12] [--steps 3] [--set bn_nD=128] [--out file]
"""
import argparse
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ofb_b200  # noqa: F401
from ofb_b200 import ops
from ofb_b200.engine import SearchStepEngine

MODELS = {"tiny": (192, 3), "small": (384, 6), "base": (768, 12)}
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="small")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--depth", type=int, default=12)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--set", action="append", default=[], help="engine attribute override, e.g. bn_nD=128")
ap.add_argument("--out", default=None)
ap.add_argument("--workload", default="search", choices=["search", "finetune"],
                help="finetune: FinetuneStepEngine on the pruned subnet of bench.py FT_SUBNET (BASELINE configs[4])")
args = ap.parse_args()

D, H = MODELS[args.model]
if args.workload == "finetune":
    from bench import FT_SUBNET
    from ofb_b200.finetune_engine import FinetuneStepEngine
    eng = FinetuneStepEngine(batch=args.batch, lr=2.5e-4, **FT_SUBNET)
    args.model = "pruned-subnet"
else:
    eng = SearchStepEngine(D, H, args.depth, args.batch, drop_path_rate=0.1, lr=2.5e-4)
for kv in args.set:
    k, v = kv.split("=")
    setattr(eng, k, int(v))
eng.init_params(seed=0)
if args.workload == "search":
    eng.set_schedule(0.0)
g = torch.Generator(device="cpu").manual_seed(1)
img = torch.randn(args.batch, 3, 224, 224, generator=g).cuda()
lab = torch.randint(0, 1000, (args.batch,), generator=g).cuda()
for _ in range(3):
    eng.step(img, lab)
torch.cuda.synchronize()
ops.OP_TIMING = []
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(args.steps):
    eng.step(img, lab)
t1.record()
torch.cuda.synchronize()
rec, ops.OP_TIMING = ops.OP_TIMING, None
agg = defaultdict(lambda: [0, 0.0])
for name, tag, e0, e1 in rec:
    k = name.replace("ofb_", "") + (" " + tag if tag else "")
    agg[k][0] += 1
    agg[k][1] += e0.elapsed_time(e1) * 1e3
tot = sum(v[1] for v in agg.values()) / args.steps
lines = [f"{args.model} B{args.batch} depth{args.depth} {' '.join(args.set)}: sum of kernels {tot / 1e3:.3f} ms/step; "
         f"host-launched wall {t0.elapsed_time(t1) / args.steps:.3f} ms/step"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{v[1] / args.steps / 1e3:8.3f} ms {100 * v[1] / args.steps / tot:5.1f}%  n={v[0] // args.steps:3d}  avg {v[1] / v[0]:8.1f} us  {k}")
print("\n".join(lines))
if args.out:
    open(args.out, "w").write("\n".join(lines) + "\n")
