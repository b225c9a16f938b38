"""Isolated timing of the attention kernels at the DeiT-S batch-256 shape (B = 256, H = 6, T = 197, d = 64): microseconds per
launch over 30 launches (CUDA events), for quick A/B of kernel changes. OFB_B200_LIB selects the library."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ofb_b200  # noqa: F401
from ofb_b200 import ops

B, H, T, d = int(os.environ.get("B", 256)), int(os.environ.get("H", 6)), int(os.environ.get("T", 197)), 64
D = H * d
torch.manual_seed(0)
qkv = torch.randn(B, T, 3, H, d, device="cuda").to(torch.bfloat16)
gate = torch.rand(D, device="cuda") * 0.5 + 0.5
ds = torch.ones(B, device="cuda")
o = torch.empty(B, T, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device="cuda")
dO = torch.randn(B, T, D, device="cuda").to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
pg, pb = torch.empty(B, D, device="cuda"), torch.empty(B, 3 * D, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()                       # the step never finds its operands in L2 either
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


f = timeit(lambda: ops.attention_fwd(qkv, o, lse, ds, B, T, H, d ** -0.5))
b = timeit(lambda: ops.attention_bwd(qkv, o, dO, lse, gate, ds, dqkv, pg, pb, B, T, H, d ** -0.5))
print(f"attention B={B} H={H} T={T}: fwd {f:.1f} us  bwd {b:.1f} us   (lib {os.environ.get('OFB_B200_LIB', 'product')})")
