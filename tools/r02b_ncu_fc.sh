#!/bin/bash
# source-level ncu capture of the FC1 / FC2-dgrad GEMMs of one block (product library)
bash tools/ncu_capture.sh r02b_fc 'gemm_kernel<\(int\)256, \(int\)[01], \(int\)0, \(int\)[12]' --model small --batch 256 --depth 1 > /dev/null
python tools/ncu_split_sass.py gpurun_out/r02b_fc_sass.csv gpurun_out/r02b_fc_k
python tools/ncu_summary.py gpurun_out/r02b_fc_raw.csv > gpurun_out/r02b_fc_summary.txt 2>&1
rm -f gpurun_out/r02b_fc_sass.csv gpurun_out/r02b_fc.ncu-rep
tail -5 gpurun_out/r02b_fc.log
