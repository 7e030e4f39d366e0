#!/bin/bash
# on the GPU box (via gpurun): ncu captures of the kernels of the named sections, summarised there -- only text comes back
# (gpurun_out <= 64 MiB).   scripts/gpu_ncu_all.sh <tag> <only=section,...> [launchlist]
TAG=${1:-all}; ONLY=${2:-}
mkdir -p gpurun_out
if [ -n "$3" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_${TAG}_launches.csv python scripts/ncu_all_kernels.py $ONLY > /dev/null 2>&1 < /dev/null
fi
timeout 1000 ncu --set full --clock-control none --import-source on -c 1500 -o /tmp/prof_r2_$TAG python scripts/ncu_all_kernels.py $ONLY > gpurun_out/ncu_$TAG.log 2>&1 < /dev/null
tail -2 gpurun_out/ncu_$TAG.log
timeout 200 python scripts/ncu_by_kernel.py /tmp/prof_r2_$TAG.ncu-rep > gpurun_out/r2_ncu_${TAG}_by_kernel.txt 2>&1 < /dev/null
timeout 300 python scripts/ncu_hot.py /tmp/prof_r2_$TAG.ncu-rep 14 > gpurun_out/r2_ncu_${TAG}_hot_lines.txt 2>&1 < /dev/null
wc -l gpurun_out/r2_ncu_${TAG}_by_kernel.txt gpurun_out/r2_ncu_${TAG}_hot_lines.txt
