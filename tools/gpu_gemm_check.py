"""GPU bring-up check of the tcgen05 GEMM against torch (fp32 math on the same bf16 inputs). Run under gpurun."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ofb_b200  # noqa: F401
from ofb_b200 import ops

torch.manual_seed(0)
dev = "cuda"
ok_all = True


def report(name, got, ref, tol=2e-2):
    global ok_all
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item() + 1e-9
    ok = err / scale < tol
    ok_all &= ok
    print(f"{'OK  ' if ok else 'FAIL'} {name}: max_abs_err={err:.4e} ref_max={scale:.4e} rel={err / scale:.3e}", flush=True)
    return ok


def rnd(*shape, s=1.0):
    return (torch.randn(*shape, device=dev) * s).to(torch.bfloat16)


def test_store(M, N, K, bn):
    A, B = rnd(M, K), rnd(N, K, s=0.05)
    bias = torch.randn(N, device=dev)
    gate = torch.rand(N, device=dev) + 0.5
    res = rnd(M, N)
    rs = torch.rand((M + 196) // 197, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(ops.EPI_STORE, A, B, M=M, N=N, K=K, out0=out, bn=bn)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    report(f"store plain M{M} N{N} K{K} bn{bn}", out, ref)
    ops.gemm(ops.EPI_STORE, A, B, M=M, N=N, K=K, out0=out, bias=bias, colscale=gate, rowscale=rs, rows_per_scale=197,
             res=res, bn=bn)
    torch.cuda.synchronize()
    rows = torch.arange(M, device=dev) // 197
    ref2 = rs[rows, None] * (ref + bias) * gate + res.float()
    report(f"store full  M{M} N{N} K{K} bn{bn}", out, ref2)
    outf = torch.empty(M, N, device=dev, dtype=torch.float32)
    ops.gemm(ops.EPI_STORE, A, B, M=M, N=N, K=K, out0=outf, out_fp32=True, bias=bias, bn=bn)
    torch.cuda.synchronize()
    report(f"store fp32  M{M} N{N} K{K} bn{bn}", outf, ref + bias, tol=1e-3)


def test_fc1(M, N, K, bn):
    A, B = rnd(M, K), rnd(N, K, s=0.05)
    bias = torch.randn(N, device=dev) * 0.1
    gate = torch.rand(N, device=dev) + 0.3
    u = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    h = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(ops.EPI_FC1, A, B, M=M, N=N, K=K, out0=u, out1=h, bias=bias, colscale=gate, bn=bn)
    torch.cuda.synchronize()
    ru = A.float() @ B.float().t() + bias
    report(f"fc1 u M{M} N{N} K{K} bn{bn}", u, ru)
    report(f"fc1 h M{M} N{N} K{K} bn{bn}", h, torch.nn.functional.gelu(ru * gate))


def test_fc2_dgrad(M, N, K, bn):
    # dy [M,K] x W2t [N=hidden, K=embed]  -> dh [M, hidden]
    dy, Wt = rnd(M, K), rnd(N, K, s=0.05)
    u = rnd(M, N)
    gate = torch.rand(N, device=dev) + 0.3
    rs = torch.rand((M + 196) // 197, device=dev)
    du = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    mt = (M + 127) // 128
    p0 = torch.zeros(mt, N, device=dev)
    p1 = torch.zeros(mt, N, device=dev)
    ops.gemm(ops.EPI_FC2_DGRAD, dy, Wt, M=M, N=N, K=K, out0=du, aux=u, colscale=gate, rowscale=rs, rows_per_scale=197,
             colpart0=p0, colpart1=p1, bn=bn)
    torch.cuda.synchronize()
    rows = torch.arange(M, device=dev) // 197
    dh = rs[rows, None] * (dy.float() @ Wt.float().t())
    uf = u.float().requires_grad_(True)
    gf = gate.clone().requires_grad_(True)
    hh = torch.nn.functional.gelu(uf * gf)
    hh.backward(dh)
    report(f"fc2dgrad du    M{M} N{N} K{K} bn{bn}", du, uf.grad)
    report(f"fc2dgrad dgate M{M} N{N} K{K} bn{bn}", p0.sum(0), gf.grad)
    report(f"fc2dgrad dbias M{M} N{N} K{K} bn{bn}", p1.sum(0), uf.grad.sum(0))


def test_wgrad(R, NO, KI, bn, splits=0):
    # dW[NO, KI] = dY[R, NO]^T X[R, KI]
    dY, X = rnd(R, NO), rnd(R, KI)
    dW = torch.zeros(NO, KI, device=dev)
    ops.gemm(ops.EPI_WGRAD, dY, X, M=NO, N=KI, K=R, out0=dW, a_mn=True, b_mn=True, bn=bn, k_splits=splits)
    torch.cuda.synchronize()
    ref = dY.float().t() @ X.float()
    report(f"wgrad R{R} NO{NO} KI{KI} bn{bn} s{splits}", dW, ref, tol=2e-3)


def bench(M, N, K, bn, iters=20):
    A, B = rnd(M, K), rnd(N, K, s=0.05)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(ops.EPI_STORE, A, B, M=M, N=N, K=K, out0=out, bn=bn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(ops.EPI_STORE, A, B, M=M, N=N, K=K, out0=out, bn=bn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    e0.record()
    for _ in range(iters):
        torch.matmul(A, B.t(), out=out)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"bench M{M} N{N} K{K} bn{bn}: ofb {ms * 1e3:.1f} us {tf:.0f} TF/s | cublas {ms2 * 1e3:.1f} us "
          f"{2.0 * M * N * K / ms2 / 1e9:.0f} TF/s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), flush=True)
    t0 = time.time()
    if which in ("all", "store"):
        test_store(256, 128, 64, 128)     # single tile, single k-block
        test_store(256, 128, 384, 128)
        test_store(1576, 384, 384, 128)   # ragged M
        test_store(1576, 1152, 384, 192)
        test_store(1576, 1536, 384, 256)
        test_store(1576, 384, 1536, 64)
        test_store(256, 1000, 384, 0)     # ragged N, auto tile
        test_store(50432, 1152, 384, 0)
    if which in ("all", "fc1"):
        test_fc1(1576, 1536, 384, 256)
        test_fc1(1576, 768, 192, 128)
    if which in ("all", "dgrad"):
        test_fc2_dgrad(1576, 1536, 384, 256)
        test_fc2_dgrad(1576, 768, 192, 192)
    if which in ("all", "wgrad"):
        test_wgrad(128, 128, 128, 128, 1)
        test_wgrad(1576, 1536, 384, 192)
        test_wgrad(1576, 384, 1536, 256)
        test_wgrad(1576, 1000, 384, 128)
        test_wgrad(50432, 1152, 384, 0)
    if which in ("all", "bench"):
        bench(50432, 1152, 384, 192)
        bench(50432, 384, 384, 128)
        bench(50432, 1536, 384, 256)
        bench(50432, 384, 1536, 128)
        bench(8192, 8192, 8192, 256)
    print(f"ALL_OK={ok_all} ({time.time() - t0:.1f}s)", flush=True)
    sys.exit(0 if ok_all else 1)
