#!/bin/bash
# ncu capture of scripts/prof_ops.py (second call of every op), reduced to CSV on the box so that the result fits
# gpurun's 64 MiB return limit.   usage: scripts/ncu_ops.sh <tag> [groups...]
tag=$1; shift
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --clock-control none \
  --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section InstructionStats --metrics dram__bytes_read.sum,dram__bytes_write.sum \
  -k regex:'^(?!.*(at::|elementwise|pack_weights)).*' -c 120 -o /tmp/prof_ops_$tag python scripts/prof_ops.py "$@" > gpurun_out/prof_ops_$tag.log 2>&1
ncu -i /tmp/prof_ops_$tag.ncu-rep --page raw --csv > gpurun_out/prof_ops_$tag.csv 2>/dev/null
ls -la /tmp/prof_ops_$tag.ncu-rep gpurun_out/prof_ops_$tag.csv
tail -2 gpurun_out/prof_ops_$tag.log
