"""What bounds each GEMM of the step? A -DOFB_GEMM_DEBUG build (never the product library) can switch off, at run time,
(1) the L2 -> SM operand feed (the producer stops issuing TMA loads once the ring is full; the MMAs run on stale tiles),
(2) the epilogue's output stores, (4) the epilogue's residual / saved-activation panel loads. Timing every step GEMM with each
combination tells which resource its duration follows.

  build (CPU box):  python tools/gemm_bound.py --build      -> tools/micro/libofb_b200_gemmdbg.so
  run (GPU box):    OFB_B200_LIB=tools/micro/libofb_b200_gemmdbg.so python tools/gemm_bound.py
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "micro", "libofb_b200_gemmdbg.so")

if "--build" in sys.argv:
    csrc = os.path.join(ROOT, "once-for-both_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith(".cu")]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "-DOFB_GEMM_DEBUG", "-shared", "-o", OUT] + srcs
    subprocess.check_call(cmd)
    print(OUT)
    sys.exit(0)

import torch  # noqa: E402

import ofb_b200  # noqa: E402,F401
from ofb_b200 import _lib, ops  # noqa: E402

lib = _lib.lib()
lib.ofb_debug_gemm_flags.argtypes = [C.c_int]
dev = "cuda"
bf = torch.bfloat16
M, D, HID, T, B = 50432, 384, 1536, 197, 256
torch.manual_seed(0)
x = torch.randn(M, D, device=dev).to(bf)
res = torch.randn(M, D, device=dev).to(bf)
y = torch.empty(M, D, device=dev, dtype=bf)
qkv = torch.empty(M, 3 * D, device=dev, dtype=bf)
dqkv = torch.randn(M, 3 * D, device=dev).to(bf)
Wqkv = (torch.randn(3 * D, D, device=dev) * .04).to(bf)
Wp = (torch.randn(D, D, device=dev) * .04).to(bf)
W1 = (torch.randn(HID, D, device=dev) * .04).to(bf)
W2 = (torch.randn(D, HID, device=dev) * .04).to(bf)
ldT = (M + 7) // 8 * 8
u = torch.randn(HID, ldT, device=dev).to(bf)
h = torch.randn(HID, ldT, device=dev).to(bf)
du = torch.empty(HID, ldT, device=dev, dtype=bf)
bias_d = torch.zeros(3 * D, device=dev)
bias_h = torch.zeros(HID, device=dev)
gate_d = torch.rand(3 * D, device=dev) + .5
gate_h = torch.rand(HID, device=dev) + .5
dp = torch.ones(B, device=dev)
parts = ops.gemm_mlp_partial_rows(M, 256)
cp0, cp1 = torch.empty(parts, HID, device=dev), torch.empty(parts, HID, device=dev)
gW = torch.zeros(HID, D, device=dev)

CASES = {
    "qkv fwd   (store, K384 N1152)": lambda: ops.gemm(ops.EPI_STORE, x, Wqkv, M=M, N=3 * D, K=D, out0=qkv, bias=bias_d, colscale=gate_d[:D].contiguous(), colscale_period=D),
    "proj fwd  (store+res, K384 N384)": lambda: ops.gemm(ops.EPI_STORE, x, Wp, M=M, N=D, K=D, out0=y, bias=bias_d[:D].contiguous(), rowscale=dp, rows_per_scale=T, bias_rowscaled=True, res=res),
    "fc1       (FC1, K384)": lambda: ops.gemm(ops.EPI_FC1, W1, x, M=HID, N=M, K=D, out0=u, out1=h, bias=bias_h, colscale=gate_h, rowscale=dp, rows_per_scale=T, bn=256),
    "fc2 fwd   (store+res, K1536 a_mn)": lambda: ops.gemm(ops.EPI_STORE, h, W2, M=M, N=D, K=HID, out0=y, bias=bias_d[:D].contiguous(), rowscale=dp, rows_per_scale=T, bias_rowscaled=True, res=res, a_mn=True),
    "fc2 dgrad (FC2_DGRAD, K384)": lambda: ops.gemm(ops.EPI_FC2_DGRAD, W2, x, M=HID, N=M, K=D, out0=du, aux=u, colscale=gate_h, rowscale=dp, rows_per_scale=T, colpart0=cp0, colpart1=cp1, a_mn=True, bn=256),
    "fc1 dgrad (store+res, K1536 a_mn b_mn)": lambda: ops.gemm(ops.EPI_STORE, du, W1, M=M, N=D, K=HID, out0=y, a_mn=True, b_mn=True, res=res),
    "qkv dgrad (store+res, K1152 b_mn)": lambda: ops.gemm(ops.EPI_STORE, dqkv, Wqkv, M=M, N=D, K=3 * D, out0=y, b_mn=True, res=res),
    "proj dgrad(store, K384 b_mn)": lambda: ops.gemm(ops.EPI_STORE, x, Wp, M=M, N=D, K=D, out0=y, b_mn=True, rowscale=dp, rows_per_scale=T),
    "fc1 wgrad (WGRAD, K50432)": lambda: ops.gemm(ops.EPI_WGRAD, du, x, M=HID, N=D, K=M, out0=gW, b_mn=True),
}


def timeit(fn, iters=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


import subprocess, threading, time
clk = []
def sample():
    while not stop:
        try:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout
            clk.append(int(o.strip().splitlines()[0]))
        except Exception:
            pass
        time.sleep(0.05)
stop = False
th = threading.Thread(target=sample, daemon=True); th.start()
print(f"{'GEMM':42s} {'normal':>8s} {'no feed':>8s} {'no store':>9s} {'no aux':>8s} {'none':>8s} {'mainloop':>9s} {'ml+feed':>8s}   (us)   [SM MHz during the row]")
only = os.environ.get("OFB_BOUND_CASES")          # comma-separated substrings of the case names
flag_sets = [int(f) for f in os.environ.get("OFB_BOUND_FLAGS", "0,1,2,4,7,15,14").split(",")]
if only or "OFB_BOUND_FLAGS" in os.environ:
    print("flags per column:", flag_sets)
for name, fn in CASES.items():
    if only and not any(o in name for o in only.split(",")):
        continue
    row = []
    c0 = len(clk)
    for flags in flag_sets:
        lib.ofb_debug_gemm_flags(flags)
        row.append(timeit(fn, iters=40))
    lib.ofb_debug_gemm_flags(0)
    cs = sorted(clk[c0:]) or [0]
    print(f"{name:42s} " + " ".join(f"{t:8.1f}" for t in row) + f"   [{cs[0]}..{cs[len(cs)//2]}..{cs[-1]}]")
stop = True
