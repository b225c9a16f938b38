#!/bin/bash
# usage (under gpurun): tools/round_profile_attn.sh <tag>   - reduced round profile after an attention / LayerNorm change:
# ncu launch list of one step + ncu --set full of the attention kernels of one block (raw summary + SASS hot spots)
tag=$1
o=gpurun_out/${tag}
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${o}_launches.csv \
    python tools/profile_step.py --model small --batch 256 > ${o}_launches.log 2>&1
python tools/summarize_launches.py ${o}_launches.csv > ${o}_launches.txt 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k "regex:attn_" -o ${o}_full_attn_ -f python tools/profile_step.py --model small --batch 256 --depth 1 > ${o}_full_attn_.log 2>&1
ncu -i ${o}_full_attn_.ncu-rep --page raw --csv > ${o}_full_attn__raw.csv 2>/dev/null
python tools/ncu_summary.py ${o}_full_attn__raw.csv > ${o}_full_attn_.txt 2>&1
rm -f ${o}_full_attn_.ncu-rep
tail -25 ${o}_launches.txt; cat ${o}_full_attn_.txt
