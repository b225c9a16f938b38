"""LayerNorm kernels at pruned embedding widths, in isolation: CUDA-event time per launch and algorithmic GB/s (fwd 2 M D 2 B,
bwd (3 + dres) M D 2 B) at M = 256 x 197 rows. Run once per setting of OFB_LN_WORDS (0 = the 16-byte-chunk kernels, 1 = the
word-granular packed kernels; the switch is read once per process)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ofb_b200  # noqa: E402,F401
from ofb_b200 import ops  # noqa: E402

M = 256 * 197
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print(f"OFB_LN_WORDS={os.environ.get('OFB_LN_WORDS', '1')}  M={M}")
for D, Dv in ((288, 288), (336, 336), (240, 240), (256, 252), (384, 384), (192, 192)):
    x = (torch.randn(M, D, device="cuda") * 2).to(torch.bfloat16)
    x[:, Dv:] = 0
    dy, dres = torch.randn(M, D, device="cuda").to(torch.bfloat16), torch.randn(M, D, device="cuda").to(torch.bfloat16)
    gamma, beta = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    y, dx = torch.empty_like(x), torch.empty_like(x)
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    R = ops.layernorm_bwd_parts(M)
    pg, pb, pd = (torch.zeros(R, D, device="cuda") for _ in range(3))
    rs = torch.ones(256, device="cuda")
    cases = {"fwd": (lambda: ops.layernorm_fwd(x, gamma, beta, y, mean, rstd, 1e-6, d_valid=Dv), 2),
             "bwd": (lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, pg, pb, pd, rs, 197, d_valid=Dv), 3),
             "bwd+res": (lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, pg, pb, pd, rs, 197, dres=dres, d_valid=Dv), 4)}
    for name, (fn, nrw) in cases.items():
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()                      # L2 flush between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print(f"  D={D} Dv={Dv} {name:8s} {t * 1e3:7.1f} us  {nrw * M * D * 2 / (t * 1e-3) / 1e9:7.0f} GB/s")
