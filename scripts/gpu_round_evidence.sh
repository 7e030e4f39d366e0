#!/bin/bash
# usage (GPU box): scripts/gpu_round_evidence.sh <tag>  -- the driver's own commands (default bench, reference arm, smoke),
# the N=1 baseline of the strong-scaling run, and compute-sanitizer logs; bounded, text only.
tag=$1
mkdir -p gpurun_out
( time timeout 600 python bench.py ) > gpurun_out/${tag}_default_bench.json 2> gpurun_out/${tag}_default_bench.err < /dev/null
tail -4 gpurun_out/${tag}_default_bench.err | grep real
python scripts/bench_brief.py default < gpurun_out/${tag}_default_bench.json
( time timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${tag}_reference_arm.json 2> gpurun_out/${tag}_reference_arm.err < /dev/null
tail -4 gpurun_out/${tag}_reference_arm.err | grep real; cut -c1-400 gpurun_out/${tag}_reference_arm.json
timeout 300 python bench.py --steps 4 --warmup 3 --secondary 0 --no-cpu-baseline --scaling strong --batch 32768 > gpurun_out/${tag}_n1_strong.json 2> gpurun_out/${tag}_n1_strong.err < /dev/null
python scripts/bench_brief.py n1_strong < gpurun_out/${tag}_n1_strong.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/${tag}_memcheck.log python scripts/ncu_all_kernels.py small > gpurun_out/${tag}_memcheck.out 2>&1 < /dev/null
tail -3 gpurun_out/${tag}_memcheck.log
