#!/bin/bash
# One `ncu --set full` capture of a hot kernel at the stage-0 shape of workload K2, reduced on the box to the details page and
# the raw CSV (the .ncu-rep itself is too big to travel).   usage: scripts/ncu_full.sh <tag> <kernel regex> <prof_kernels target> [skip] [count]
tag=$1; regex=$2; target=$3; skip=${4:-2}; count=${5:-1}
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none $NCU_EXTRA -k regex:$regex -s $skip -c $count -o /tmp/full_$tag python scripts/prof_kernels.py $target > gpurun_out/full_$tag.log 2>&1
ncu -i /tmp/full_$tag.ncu-rep --page details > gpurun_out/full_${tag}_details.txt 2>/dev/null
ncu -i /tmp/full_$tag.ncu-rep --page raw --csv > gpurun_out/full_${tag}_raw.csv 2>/dev/null
grep -E "dram__bytes_(read|write).sum |gpu__time_duration.sum" gpurun_out/full_${tag}_raw.csv | head -5
tail -1 gpurun_out/full_$tag.log
