#!/bin/bash
# same-box A/B of the whole step between library builds: LIBS="a.so b.so" (a path or "product"), alternating, REPS rounds
mkdir -p gpurun_out
for rep in $(seq 1 ${REPS:-2}); do
for l in ${LIBS:-product}; do
  if [ "$l" = "product" ]; then unset OFB_B200_LIB; else export OFB_B200_LIB=$l; fi
  timeout 600 python bench.py --no-eager --no-extra --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err || tail -5 gpurun_out/bench_ab.err
  python - "$l" <<PY
import json, sys
d = json.load(open("gpurun_out/bench_ab.json"))
print(f"{sys.argv[1]:44s} {d['value']:8.0f} img/s {d['ms_per_step']:7.3f} ms  kernel sum {d['kernel_sum_ms']:7.3f}  sm {d['clocks']['sm_mhz']}",
      {k: round(v["ms_per_step"], 3) for k, v in d["kernel_shares"].items()})
PY
done
done
