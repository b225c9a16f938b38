"""Bring-up check of the CTA-pair (cta_group::2) GEMM path against torch on small and ragged shapes. Run under gpurun with a
short timeout:  timeout 120 python tools/gemm_pair_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ofb_b200  # noqa: F401
from ofb_b200 import ops

dev = "cuda"
torch.manual_seed(0)
ok_all = True


def rnd(*shape, s=1.0):
    return (torch.randn(*shape, device=dev) * s).to(torch.bfloat16)


def check(name, got, ref, tol=2e-2):
    global ok_all
    torch.cuda.synchronize()
    err = (got.float() - ref.float()).abs().max().item() / (ref.float().abs().max().item() + 1e-9)
    ok = err < tol
    ok_all &= ok
    print(("OK  " if ok else "FAIL"), name, f"rel={err:.3e}", flush=True)


for (M, N, K, bn, a_mn, b_mn) in [(256, 128, 64, 128, 0, 0), (512, 256, 128, 256, 0, 0), (1576, 1152, 384, 192, 0, 0),
                                  (1571, 384, 384, 192, 0, 0), (1576, 384, 1536, 128, 1, 1), (1576, 384, 1152, 256, 0, 1),
                                  (1536, 1576, 384, 256, 0, 0), (50432, 384, 384, 192, 0, 0)]:
    A = rnd(K, M) if a_mn else rnd(M, K)
    B = rnd(K, N, s=0.05) if b_mn else rnd(N, K, s=0.05)
    Am = A.t() if a_mn else A
    Bm = B if b_mn else B.t()
    ref = Am.float() @ Bm.float()
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(ops.EPI_STORE, A, B, M=M, N=N, K=K, out0=out, bn=bn, a_mn=bool(a_mn), b_mn=bool(b_mn), lda=A.stride(0), ldb=B.stride(0))
    check(f"store M{M} N{N} K{K} bn{bn} a{a_mn}b{b_mn}", out, ref)
    res = rnd(M, N)
    bias = torch.randn(N, device=dev)
    ops.gemm(ops.EPI_STORE, A, B, M=M, N=N, K=K, out0=out, bn=bn, a_mn=bool(a_mn), b_mn=bool(b_mn), lda=A.stride(0), ldb=B.stride(0),
             res=res, bias=bias)
    check(f"store+res M{M} N{N} K{K} bn{bn} a{a_mn}b{b_mn}", out, ref + bias + res.float())
# split-K weight gradient
for (R, NO, KI, bn) in [(1576, 1536, 384, 128), (1576, 384, 1536, 256), (50432, 1536, 384, 128)]:
    dy, x = rnd(R, NO), rnd(R, KI)
    dW = torch.zeros(NO, KI, device=dev)
    ops.gemm(ops.EPI_WGRAD, dy, x, M=NO, N=KI, K=R, out0=dW, a_mn=True, b_mn=True, bn=bn)
    check(f"wgrad R{R} NO{NO} KI{KI} bn{bn}", dW, dy.float().t() @ x.float(), tol=2e-3)
print("ALL OK" if ok_all else "SOME FAILED")
