"""A/B of the data-gradient GEMM operand layout at the DeiT-S batch-256 shapes: the nn.Linear weight [N_out, K_in] read directly
as an MN-major B operand (what the engine does) against a K-major transposed bf16 copy (tile 192, CTA pairs, no column padding).
Microseconds per launch, L2 flushed between launches."""
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
dqkv = torch.randn(M, 3 * D, device=dev).to(bf)
g2 = torch.randn(M, D, device=dev).to(bf)
Wqkv = (torch.randn(3 * D, D, device=dev) * .04).to(bf)
Wp = (torch.randn(D, D, device=dev) * .04).to(bf)
W1 = (torch.randn(HID, D, device=dev) * .04).to(bf)
WqkvT, WpT, W1T = Wqkv.t().contiguous(), Wp.t().contiguous(), W1.t().contiguous()
ldT = (M + 7) // 8 * 8
du = torch.randn(HID, ldT, device=dev).to(bf)
dp = torch.ones(B, device=dev)
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
    "fc1 dgrad (K 1536, a_mn)": (
        lambda: ops.gemm(ops.EPI_STORE, du, W1, M=M, N=D, K=HID, out0=y, a_mn=True, b_mn=True, res=res),
        lambda: ops.gemm(ops.EPI_STORE, du, W1T, M=M, N=D, K=HID, out0=y, a_mn=True, res=res)),
    "qkv dgrad (K 1152)": (
        lambda: ops.gemm(ops.EPI_STORE, dqkv, Wqkv, M=M, N=D, K=3 * D, out0=y, b_mn=True, res=res),
        lambda: ops.gemm(ops.EPI_STORE, dqkv, WqkvT, M=M, N=D, K=3 * D, out0=y, res=res)),
    "proj dgrad (K 384)": (
        lambda: ops.gemm(ops.EPI_STORE, g2, Wp, M=M, N=D, K=D, out0=y, b_mn=True, rowscale=dp, rows_per_scale=T),
        lambda: ops.gemm(ops.EPI_STORE, g2, WpT, M=M, N=D, K=D, out0=y, rowscale=dp, rows_per_scale=T)),
}
ref = {}
for name, (f_mn, f_k) in cases.items():
    f_mn(); a = y.float().clone(); f_k(); b = y.float().clone()
    err = ((a - b).norm() / a.norm()).item()
    print(f"{name:28s} weight as stored (MN-major B) {timeit(f_mn):7.1f} us   transposed copy (K-major B) {timeit(f_k):7.1f} us   rel diff {err:.1e}")
x = torch.randn(12 * (3 * D * D + D * D + HID * D), device=dev).to(bf)
print(f"transposing the 36 weight matrices of a step with torch (upper bound for a batched kernel): "
      f"{timeit(lambda: [Wqkv.t().contiguous(), Wp.t().contiguous(), W1.t().contiguous()]) * 12:.1f} us")
