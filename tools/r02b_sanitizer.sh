#!/bin/bash
# compute-sanitizer over the GPU parity tests (SURVEY section 5 asks for a sanitizer pass); summary lines only
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/r02b_${tool}.log \
      python -m pytest tests/test_kernels_gpu.py tests/test_step_gpu.py -m gpu -q -x -k "not batch256 and not benchmark" 2>&1 | tail -1
  echo "$tool: $(grep -E 'ERROR SUMMARY' gpurun_out/r02b_${tool}.log | tail -1)"
  grep -E "Invalid|Error:|hazard|Barrier error" gpurun_out/r02b_${tool}.log | sort | uniq -c | head -10
done
