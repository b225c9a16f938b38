"""norm_targets (PMIM target normalisation, 47 x 47 window) timed alone at the benchmark shape: microseconds per launch at two
masked fractions."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ofb_b200  # noqa: F401
from ofb_b200 import ops

B, L = 256, 196
img = torch.randn(B, 3, 224, 224, device="cuda")
tgt = torch.zeros(B * L, 768, device="cuda")
for frac in (0.25, 0.75):
    mask = (torch.rand(B, L, device="cuda") < frac).float()
    for _ in range(3):
        ops.norm_targets(img, mask, tgt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.norm_targets(img, mask, tgt)
    e1.record()
    torch.cuda.synchronize()
    print(f"norm_targets, {frac:.0%} of the patches masked: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
