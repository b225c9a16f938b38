#!/bin/bash
# round-2 profile artefacts (under gpurun): ncu launch list of one DeiT-S step, ncu --set full of the GEMM / attention / LayerNorm
# kernels of one block, the GEMM bound-finding experiment, step timeline, determinism A/B, the full bench line
tag=${1:-r02a}
o=gpurun_out/${tag}
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${o}_launches.csv \
    python tools/profile_step.py --model small --batch 256 > ${o}_launches.log 2>&1
python tools/summarize_launches.py ${o}_launches.csv > ${o}_launches.txt 2>&1
for k in gemm_kernel attn_ ln_; do
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:${k}" -o ${o}_full_${k} -f python tools/profile_step.py --model small --batch 256 --depth 1 > ${o}_full_${k}.log 2>&1
  ncu -i ${o}_full_${k}.ncu-rep --page raw --csv > ${o}_full_${k}_raw.csv 2>/dev/null
  python tools/ncu_summary.py ${o}_full_${k}_raw.csv > ${o}_full_${k}.txt 2>&1
  rm -f ${o}_full_${k}.ncu-rep ${o}_full_${k}_raw.csv
done
OFB_B200_LIB=tools/micro/libofb_b200_gemmdbg.so timeout 300 python tools/gemm_bound.py > ${o}_gemm_bound.txt 2>&1
timeout 300 python tools/step_timeline.py 2>&1 | grep -v -i warn > ${o}_step_timeline.txt
bash tools/r02_ab_det.sh > ${o}_determinism_ab.txt 2>&1
timeout 900 python bench.py > ${o}_bench_small.json 2> ${o}_bench_small.err
timeout 300 python bench.py --impl reference > ${o}_bench_reference.json 2>/dev/null
ls -la gpurun_out | tail -20
tail -30 ${o}_launches.txt
