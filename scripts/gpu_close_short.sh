#!/bin/bash
# usage (GPU box): scripts/gpu_close_short.sh <tag> -- full GPU test suite and the default bench line
tag=$1
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest_gpu.txt
( time timeout 600 python bench.py ) > gpurun_out/${tag}_default_bench.json 2> gpurun_out/${tag}_default_bench.err < /dev/null
grep real gpurun_out/${tag}_default_bench.err
python scripts/bench_brief.py default < gpurun_out/${tag}_default_bench.json
