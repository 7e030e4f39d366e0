#!/bin/bash
# usage (GPU box): scripts/gpu_close.sh <tag> -- full GPU test suite, smoke(), the default bench line and the reference arm
tag=$1
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 600 python bench.py ) > gpurun_out/${tag}_default_bench.json 2> gpurun_out/${tag}_default_bench.err < /dev/null
grep real gpurun_out/${tag}_default_bench.err
python scripts/bench_brief.py default < gpurun_out/${tag}_default_bench.json
( time timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${tag}_reference_arm.json 2> gpurun_out/${tag}_reference_arm.err < /dev/null
grep real gpurun_out/${tag}_reference_arm.err; cut -c1-200 gpurun_out/${tag}_reference_arm.json
