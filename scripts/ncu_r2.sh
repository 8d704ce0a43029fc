#!/bin/bash
# Round-2 `ncu --set full` capture of every hot kernel family (one launch each, of scripts/prof_all.py, filtered by kernel name), reduced
# on the box to the details page and the raw CSV; scripts/ncu_traffic.py turns the raw page into profiles/r2_ncu_summary.* here.
mkdir -p gpurun_out
PROF_PASSES=1 timeout 1700 ncu --set full --import-source on --clock-control none -k "regex:conv_line_tma|wgrad_line|gemm_tma|wgrad_gemm|wgrad_reduce_batch|gate_fuse|bn_act2|dice_multi|ln_metapool|breg_|fp_" -o /tmp/r2_all python scripts/prof_all.py > gpurun_out/r2_ncu_all.log 2>&1
ncu -i /tmp/r2_all.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw.csv 2>/dev/null
ncu -i /tmp/r2_all.ncu-rep --page details > gpurun_out/r2_ncu_details.txt 2>/dev/null
tail -2 gpurun_out/r2_ncu_all.log; wc -l gpurun_out/r2_ncu_raw.csv
