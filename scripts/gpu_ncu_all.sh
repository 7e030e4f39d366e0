#!/bin/bash
# on the GPU box (via gpurun): ncu captures of EVERY kernel, summarised there -- only text comes back (gpurun_out <= 64 MiB)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_all_launches.csv python scripts/ncu_all_kernels.py > /dev/null 2>&1 < /dev/null
timeout 900 ncu --set full --clock-control none --import-source on -c 900 -o /tmp/prof_r2_all python scripts/ncu_all_kernels.py > gpurun_out/ncu_all.log 2>&1 < /dev/null
tail -2 gpurun_out/ncu_all.log
timeout 200 python scripts/ncu_by_kernel.py /tmp/prof_r2_all.ncu-rep > gpurun_out/r2_ncu_all_by_kernel.txt 2>&1 < /dev/null
timeout 300 python scripts/ncu_hot.py /tmp/prof_r2_all.ncu-rep 14 > gpurun_out/r2_ncu_all_hot_lines.txt 2>&1 < /dev/null
wc -l gpurun_out/r2_all_launches.csv gpurun_out/r2_ncu_all_by_kernel.txt gpurun_out/r2_ncu_all_hot_lines.txt
