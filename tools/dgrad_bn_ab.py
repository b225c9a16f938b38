"""N-tile A/B of the N = embed_dim GEMMs at the DeiT-S batch-256 shapes (microseconds per launch, L2 flushed between launches):
the library's choice (bn = 0) against explicit 128 / 192 / 256-wide tiles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ofb_b200  # noqa: F401
from ofb_b200 import ops

dev, bf = "cuda", torch.bfloat16
M, D, HID, T, B = 50432, 384, 1536, 197, 256
torch.manual_seed(0)
res = torch.randn(M, D, device=dev).to(bf)
y = torch.empty(M, D, device=dev, dtype=bf)
x = torch.randn(M, D, device=dev).to(bf)
dqkv = torch.randn(M, 3 * D, device=dev).to(bf)
Wqkv = (torch.randn(3 * D, D, device=dev) * .04).to(bf)
Wp = (torch.randn(D, D, device=dev) * .04).to(bf)
W1 = (torch.randn(HID, D, device=dev) * .04).to(bf)
W2 = (torch.randn(D, HID, device=dev) * .04).to(bf)
ldT = (M + 7) // 8 * 8
du = torch.randn(HID, ldT, device=dev).to(bf)
dp = torch.ones(B, device=dev)
bias = torch.zeros(D, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


cases = {
    "fc1 dgrad (K 1536, a_mn b_mn, res)": lambda bn: ops.gemm(ops.EPI_STORE, du, W1, M=M, N=D, K=HID, out0=y, a_mn=True, b_mn=True, res=res, bn=bn),
    "qkv dgrad (K 1152, b_mn, res)": lambda bn: ops.gemm(ops.EPI_STORE, dqkv, Wqkv, M=M, N=D, K=3 * D, out0=y, b_mn=True, res=res, bn=bn),
    "proj dgrad (K 384, b_mn)": lambda bn: ops.gemm(ops.EPI_STORE, x, Wp, M=M, N=D, K=D, out0=y, b_mn=True, rowscale=dp, rows_per_scale=T, bn=bn),
    "fc2 fwd (K 1536, a_mn, res)": lambda bn: ops.gemm(ops.EPI_STORE, du, W2, M=M, N=D, K=HID, out0=y, bias=bias, rowscale=dp, rows_per_scale=T, bias_rowscaled=True, res=res, a_mn=True, bn=bn),
    "proj fwd (K 384, res)": lambda bn: ops.gemm(ops.EPI_STORE, x, Wp, M=M, N=D, K=D, out0=y, bias=bias, rowscale=dp, rows_per_scale=T, bias_rowscaled=True, res=res, bn=bn),
}
bns = [int(b) for b in os.environ.get("OFB_BNS", "0,128,192,256").split(",")]
print(f"{'GEMM':38s} " + " ".join(f"bn={b:<5d}" for b in bns))
for name, f in cases.items():
    print(f"{name:38s} " + " ".join(f"{timeit(lambda: f(b)):8.1f}" for b in bns))
