#!/bin/bash
# usage: tools/ncu_capture.sh <tag> <kernel regex> [profile_step args...]
# ncu --set full capture of the kernels matching <regex> in one profiled search step; exports the raw / source pages as CSV
# (gpurun_out/ is capped at 64 MiB, so the .ncu-rep itself is kept only when small).
tag=$1; regex=$2; shift 2
out=gpurun_out/${tag}
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k "regex:${regex}" -o ${out} -f python tools/profile_step.py "$@" > ${out}.log 2>&1
ncu -i ${out}.ncu-rep --page raw --csv > ${out}_raw.csv 2>/dev/null
ncu -i ${out}.ncu-rep --page source --csv --print-source sass > ${out}_sass.csv 2>/dev/null
sz=$(stat -c %s ${out}.ncu-rep)
if [ "$sz" -gt 30000000 ]; then rm -f ${out}.ncu-rep; fi
ls -la gpurun_out/ | tail -8
