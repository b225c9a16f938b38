#!/bin/bash
# usage (under gpurun): tools/round_profile.sh <tag>
# Produces every profile artefact of a round under gpurun_out/<tag>_*: in-situ breakdown, ncu launch list, ncu --set full of
# the GEMM / attention / LayerNorm kernels of one block (raw + SASS CSV), and the extra-config bench lines.
tag=$1
o=gpurun_out/${tag}
python tools/step_breakdown.py --out ${o}_breakdown_small.txt > /dev/null 2>&1
python tools/step_breakdown.py --model tiny --batch 1024 --out ${o}_breakdown_tiny1024.txt > /dev/null 2>&1
python tools/step_breakdown.py --model base --batch 256 --out ${o}_breakdown_base256.txt > /dev/null 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${o}_launches.csv \
    python tools/profile_step.py --model small --batch 256 > ${o}_launches.log 2>&1
python tools/summarize_launches.py ${o}_launches.csv > ${o}_launches.txt 2>&1
for k in gemm_kernel attn_ ln_; do
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:${k}" -o ${o}_full_${k} -f python tools/profile_step.py --model small --batch 256 --depth 1 > ${o}_full_${k}.log 2>&1
  ncu -i ${o}_full_${k}.ncu-rep --page raw --csv > ${o}_full_${k}_raw.csv 2>/dev/null
  python tools/ncu_summary.py ${o}_full_${k}_raw.csv > ${o}_full_${k}.txt 2>&1
  if [ "$k" != "gemm_kernel" ]; then
    ncu -i ${o}_full_${k}.ncu-rep --page source --csv --print-source sass > ${o}_full_${k}_sass.csv 2>/dev/null
  fi
  rm -f ${o}_full_${k}.ncu-rep
done
python bench.py --model tiny --batch 1024 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > ${o}_bench_tiny1024.json
python bench.py --model base --batch 256 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > ${o}_bench_base256.json
ls -la gpurun_out | tail -30
