#!/bin/bash
# FC1 / FC2-dgrad epilogue iteration: parity of the transposed-hidden GEMMs, then their bound-finding table (debug library)
# flags: 1 no feed, 2 no store, 4 no aux, 8 no epilogue
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" 2>&1 | tail -3
export OFB_B200_LIB=tools/micro/libofb_b200_gemmdbg.so
OFB_BOUND_CASES="fc1  ,fc2 dgrad" OFB_BOUND_FLAGS=${FLAGS:-0,1,2,4,7,15,14} timeout 300 python tools/gemm_bound.py 2>&1 | grep -v Warn
