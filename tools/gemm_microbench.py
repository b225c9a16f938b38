"""Mainloop-rate microbenchmark of the tcgen05 GEMM per N tile / operand layout against cuBLAS (torch.matmul) on the same
shapes. Run under gpurun:  python tools/gemm_microbench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ofb_b200  # noqa: F401
from ofb_b200 import ops

dev = "cuda"


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def run(M, N, K, bns=(128, 192, 256), a_mn=False, b_mn=False):
    A = (torch.randn(K, M, device=dev) if a_mn else torch.randn(M, K, device=dev)).to(torch.bfloat16)
    B = (torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    Am = A.t() if a_mn else A
    Bm = B if b_mn else B.t()
    t = timeit(lambda: torch.matmul(Am, Bm, out=out))
    line = f"M{M} N{N} K{K} a_mn={int(a_mn)} b_mn={int(b_mn)}: cublas {fl / t / 1e12:7.1f} TF/s |"
    for bn in bns:
        t = timeit(lambda: ops.gemm(ops.EPI_STORE, A, B, M=M, N=N, K=K, out0=out, bn=bn, a_mn=a_mn, b_mn=b_mn,
                                    lda=A.stride(0), ldb=B.stride(0)))
        line += f" bn{bn} {fl / t / 1e12:7.1f}"
    print(line, flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "big"):
        run(8192, 8192, 8192)
        run(8192, 8192, 8192, a_mn=True, b_mn=True)
    if which in ("all", "step"):
        M = 50432
        run(M, 1152, 384)                 # qkv forward
        run(M, 384, 384)                  # proj forward
        run(M, 384, 1536, a_mn=True)      # fc2 forward (A = h^T)
        run(M, 384, 1152, b_mn=True)      # qkv dgrad
        run(M, 384, 1536, a_mn=True, b_mn=True)   # fc1 dgrad
        run(M, 1536, 384)                 # fc1-like plain store
