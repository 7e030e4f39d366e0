#!/bin/bash
# usage (GPU box): scripts/gpu_final_evidence.sh <tag> -- the round's closing record on one GPU: full GPU test suite, the
# default bench line, the ncu launch list of the same command and an `ncu --set full` capture of the headline path's kernels
tag=$1
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest_gpu.txt
( time timeout 600 python bench.py ) > gpurun_out/${tag}_default_bench.json 2> gpurun_out/${tag}_default_bench.err < /dev/null
grep real gpurun_out/${tag}_default_bench.err
python scripts/bench_brief.py default < gpurun_out/${tag}_default_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_computers.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --secondary 0 > /dev/null 2>&1 < /dev/null
wc -l gpurun_out/${tag}_launches_computers.csv
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:filtration_table|vorder_kernel|sweep_kernel|vicinity_light|pimg_kernel" -s 5 -c 10 -o /tmp/prof_$tag python bench.py --steps 1 --warmup 1 --no-cpu-baseline --secondary 0 > gpurun_out/ncu_$tag.log 2>&1 < /dev/null
tail -1 gpurun_out/ncu_$tag.log
timeout 200 python scripts/ncu_by_kernel.py /tmp/prof_$tag.ncu-rep > gpurun_out/${tag}_ncu_headline_by_kernel.txt 2>&1 < /dev/null
timeout 300 python scripts/ncu_hot.py /tmp/prof_$tag.ncu-rep 30 > gpurun_out/${tag}_ncu_headline_hot_lines.txt 2>&1 < /dev/null
wc -l gpurun_out/${tag}_ncu_headline_by_kernel.txt gpurun_out/${tag}_ncu_headline_hot_lines.txt
