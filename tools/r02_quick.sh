#!/bin/bash
# quick GPU iteration: kernel + step parity, then a bench without the baselines; prints the essentials
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_step_gpu.py -m gpu -q -x 2>&1 | tail -3
for pdl in ${PDLS:-0}; do
OFB_PDL=$pdl timeout 600 python bench.py --no-eager --no-extra --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err || tail -5 gpurun_out/bench_q.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_q.json"))
print("PDL=$pdl", round(d["value"]), "img/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), " gemm frac", round(d["roofline"]["frac"],3), " kernel sum", round(d["kernel_sum_ms"],3), d["clocks"])
for k in d["top_kernels"][:${NTOP:-6}]: print("   ", round(k["ms_per_step"],3), k["launches_per_step"], k["name"][:64])
print("   ", {k: round(v["ms_per_step"],3) for k,v in d["kernel_shares"].items()})
PY
done
