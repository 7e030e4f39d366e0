#!/bin/bash
# usage (on the GPU box, via gpurun --gpus N): scripts/gpu_multi_set.sh <tag> <N> [weak strong collab nccl tests]
# torchrun launches of bench.py on N GPUs, one JSON line each into gpurun_out/<tag>_n<N>_<what>.json; every run bounded.
tag=$1; N=$2; shift; shift
mkdir -p gpurun_out
port=29510
run() {  # name, bench args...
  local name=$1; shift
  port=$((port + 1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --steps 10 --warmup 3 --secondary 0 --no-cpu-baseline "$@" \
    > gpurun_out/${tag}_n${N}_${name}.json 2> gpurun_out/${tag}_n${N}_${name}.err < /dev/null
  grep -h '"metric"' gpurun_out/${tag}_n${N}_${name}.json | python scripts/bench_brief.py "n${N}_$name" 2>/dev/null || { echo "n${N}_$name FAILED"; tail -5 gpurun_out/${tag}_n${N}_${name}.err; }
}
for what in "$@"; do
  case $what in
    weak) run weak ;;
    nccl) run weak_nccl --exchange nccl ;;
    strong) run strong --scaling strong --batch 32768 --steps 4 --warmup 3 ;;
    collab) run collab_strong --workload collab --negatives 1 --batch 16384 --scaling strong ;;
    tests) timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 ;;
    *) echo "unknown $what" ;;
  esac
done
