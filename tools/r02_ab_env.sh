# A/B of an environment switch: VAR=name VALS="0 1 0 1"
for v in ${VALS:-0 1 0 1}; do
env $VAR=$v timeout 600 python bench.py --no-eager --no-extra --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err || tail -3 gpurun_out/bench_q.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_q.json"))
print("$VAR=$v", round(d["value"]), "img/s", round(d["ms_per_step"],3), "ms  kernel sum", round(d["kernel_sum_ms"],3), d["clocks"]["sm_mhz"], {k: round(v["ms_per_step"],3) for k,v in d["kernel_shares"].items()})
PY
done
