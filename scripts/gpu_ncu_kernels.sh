#!/bin/bash
# usage (GPU box): scripts/gpu_ncu_kernels.sh <tag> <only=sections> <kernel regex> [launch-skip] [launch-count]
# `ncu --set full` of the kernels matching the regex in the named sections of scripts/ncu_all_kernels.py; text summaries only.
TAG=$1; ONLY=$2; RE=$3; SKIP=${4:-0}; CNT=${5:-40}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $CNT -o /tmp/prof_r2_$TAG python scripts/ncu_all_kernels.py $ONLY > gpurun_out/ncu_$TAG.log 2>&1 < /dev/null
tail -2 gpurun_out/ncu_$TAG.log
timeout 200 python scripts/ncu_by_kernel.py /tmp/prof_r2_$TAG.ncu-rep > gpurun_out/r2_ncu_${TAG}_by_kernel.txt 2>&1 < /dev/null
timeout 300 python scripts/ncu_hot.py /tmp/prof_r2_$TAG.ncu-rep ${HOTN:-24} > gpurun_out/r2_ncu_${TAG}_hot_lines.txt 2>&1 < /dev/null
wc -l gpurun_out/r2_ncu_${TAG}_by_kernel.txt gpurun_out/r2_ncu_${TAG}_hot_lines.txt
