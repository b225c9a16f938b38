#!/bin/bash
# second-session A/B bundle: attention parity + isolated timing of library variants
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -2
for rep in 1 2; do
python tools/attn_bench.py
for l in ${LIBS:-base}; do OFB_B200_LIB=tools/micro/libofb_b200_$l.so python tools/attn_bench.py; done
done
