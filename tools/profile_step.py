"""Profiling harness: one warm search step, then `--steps` steps bracketed by cudaProfilerStart/Stop so that
`ncu --profile-from-start off` sees exactly the kernels of the hot path (B200_PROFILING.md recipe).

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py --model small --batch 256
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ofb_b200  # noqa: F401
from ofb_b200.engine import SearchStepEngine

MODELS = {"tiny": (192, 3), "small": (384, 6), "base": (768, 12)}

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="small")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--depth", type=int, default=12)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()

D, H = MODELS[args.model]
eng = SearchStepEngine(D, H, args.depth, args.batch, drop_path_rate=0.1, lr=2.5e-4)
eng.init_params(seed=0)
eng.set_schedule(0.0)
g = torch.Generator(device="cpu").manual_seed(1)
img = torch.randn(args.batch, 3, 224, 224, generator=g).cuda()
lab = torch.randint(0, 1000, (args.batch,), generator=g).cuda()
for _ in range(args.warmup):
    eng.step(img, lab)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(args.steps):
    eng.step(img, lab)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("losses", eng.scal.cpu().tolist()[:4])
