#!/bin/bash
# round 2, third session: group-panel store epilogue (gemm.cu Cfg::GP) - parity of the GEMM tests, per-GEMM bound table of the old
# and the new debug builds, same-box step A/B against the previous product build
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm" 2>&1 | tail -3
export OFB_BOUND_CASES="qkv fwd,proj fwd,fc2 fwd,proj dgrad"
for l in tools/micro/libofb_b200_gemmdbg_base.so tools/micro/libofb_b200_gemmdbg.so; do
  echo "== $l"; OFB_B200_LIB=$l python tools/gemm_bound.py 2>&1 | tail -6
done
unset OFB_BOUND_CASES
LIBS="tools/micro/libofb_b200_base.so product" REPS=2 bash tools/r02b_ab_step.sh
