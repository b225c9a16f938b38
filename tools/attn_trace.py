"""Phase timeline of attn_bwd_kernel (CTA 0, items 2..5): SM-clock stamps written by compute warps 0 / 15 and the MMA warp of a
-DOFB_ATTN_TRACE debug build (never the product library).

  build (CPU box):  python tools/attn_trace.py --build      -> tools/micro/libofb_b200_trace.so
  run (GPU box):    OFB_B200_LIB=tools/micro/libofb_b200_trace.so python tools/attn_trace.py [--heads 6]
"""
import argparse
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "micro", "libofb_b200_trace.so")

ap = argparse.ArgumentParser()
ap.add_argument("--build", action="store_true")
ap.add_argument("--heads", type=int, default=6)
ap.add_argument("--batch", type=int, default=256)
args = ap.parse_args()

if args.build:
    csrc = os.path.join(ROOT, "once-for-both_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith(".cu")]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "-DOFB_ATTN_TRACE", "-shared", "-o", OUT] + srcs
    subprocess.check_call(cmd)
    print(OUT)
    sys.exit(0)

import torch  # noqa: E402

import ofb_b200  # noqa: E402,F401
from ofb_b200 import _lib, ops  # noqa: E402

B, H, T, d = args.batch, args.heads, 197, 64
D = H * d
torch.manual_seed(0)
qkv = (torch.randn(B, T, 3, H, d, device="cuda")).to(torch.bfloat16)
gate = torch.rand(D, device="cuda") * 0.5 + 0.5
ds = torch.ones(B, device="cuda")
o = torch.empty(B, T, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device="cuda")
ops.attention_fwd(qkv, o, lse, ds, B, T, H, d ** -0.5)
dO = torch.randn(B, T, D, device="cuda").to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
pg, pb = torch.empty(B, D, device="cuda"), torch.empty(B, 3 * D, device="cuda")
lib = _lib.lib()
lib.ofb_debug_attn_trace.argtypes = [C.c_void_p, C.c_int, C.c_int]
for _ in range(3):
    ops.attention_bwd(qkv, o, dO, lse, gate, ds, dqkv, pg, pb, B, T, H, d ** -0.5)
lib.ofb_debug_attn_trace(None, 0, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.attention_bwd(qkv, o, dO, lse, gate, ds, dqkv, pg, pb, B, T, H, d ** -0.5)
e1.record()
torch.cuda.synchronize()
buf = (C.c_longlong * (3 * 4 * 32))()
lib.ofb_debug_attn_trace(buf, 3 * 4 * 32, 0)
tr = [[[buf[(r * 4 + it) * 32 + s] for s in range(32)] for it in range(4)] for r in range(3)]
print(f"B={B} H={H}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us, items per CTA {B * H / 148:.2f}")
names = ["start", "dO/V landed", "stats done", "P00", "P10", "dS00", "dS10", "P01", "epi_kv0", "P11", "dS01", "dS11", "epi_kv1",
         "dQ epi + gate", "item barrier"]
for r, who in ((0, "compute warp 0"), (1, "compute warp 15")):
    print(who)
    for it in range(4):
        t = tr[r][it]
        base = t[0]
        nxt = tr[r][it + 1][0] if it < 3 else None
        steps = " ".join(f"{names[k]}+{t[k] - t[k - 1]}" for k in range(1, 15))
        print(f"  item {it + 2}: total {t[14] - base} clk ({'next start +' + str(nxt - t[14]) if nxt else ''}) | {steps}")
        print(f"           before the dO/V wait (row statistics loads issued): +{t[15] - t[0]}")
        print(f"           waits: S_FULL {t[20]} PB_FREE {t[21]} DP_FULL {t[22]} ACC_FULL {t[23]} DQ_FULL {t[24]}")
print("MMA warp")
for it in range(4):
    t = tr[2][it]
    print(f"  item {it + 2}: wait QK_FULL {t[1] - t[0]} | issue span {t[2] - t[1]} | waits: P_FULL {t[8]} ACC_EMPTY {t[9]} DS_FULL {t[10]} "
          f"DQ_EMPTY {t[11]}" + (f" | next item start +{tr[2][it + 1][0] - t[2]}" if it < 3 else ""))
