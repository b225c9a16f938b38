#!/bin/bash
# round-2 GPU call 1: state of parity at the fixed tolerance (whole GPU suite, no -x), emulated rounding study at scale,
# bench with CUPTI breakdown / eager baseline / extra configs
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 2400 python -m pytest tests -m gpu -q -s -rA > gpurun_out/pytest_call1.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_call1.log
timeout 300 python tests/precision_study.py --dim 384 --heads 6 --depth 12 --batch 64 --device cuda > gpurun_out/precision_s64.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_call1.json 2> gpurun_out/bench_call1.err
echo "bench exit $?" >> gpurun_out/bench_call1.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref_call1.json 2> gpurun_out/bench_ref_call1.err
tail -5 gpurun_out/pytest_call1.log
cat gpurun_out/bench_call1.json | head -c 3000
