#!/bin/bash
# usage (on the GPU box, via gpurun): scripts/gpu_bench_set.sh <tag> [configs...]
# runs bench.py on the named configurations, one JSON line each into gpurun_out/<tag>_<config>.json, and prints a digest.
# every run is bounded by `timeout`; nothing here reads stdin.
tag=$1; shift
mkdir -p gpurun_out
run() {  # name, bench args...
  local name=$1; shift
  timeout 240 python bench.py --steps 10 --warmup 3 --secondary 0 "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err < /dev/null
  python scripts/bench_brief.py "$name" < gpurun_out/${tag}_${name}.json 2>/dev/null || { echo "$name FAILED"; tail -3 gpurun_out/${tag}_${name}.err; }
}
for cfg in "$@"; do
  case $cfg in
    computers) run computers --no-cpu-baseline ;;
    computers_cpu) run computers_cpu ;;
    computers_ext) run computers_ext --extended 1 --batch 256 --no-cpu-baseline ;;
    hop1_ext) run hop1_ext --hop 1 --extended 1 --batch 8192 --no-cpu-baseline ;;
    hop1) run hop1 --hop 1 --batch 8192 --no-cpu-baseline ;;
    cora) run cora --workload cora --batch 5278 --no-cpu-baseline ;;
    cora_ext) run cora_ext --workload cora --batch 5278 --extended 1 --no-cpu-baseline ;;
    pubmed) run pubmed --workload pubmed --batch 8192 --no-cpu-baseline ;;
    pubmed_ext) run pubmed_ext --workload pubmed --batch 8192 --extended 1 --no-cpu-baseline ;;
    ppi) run ppi --workload ppi --mode node --extended 1 --batch 4096 --no-cpu-baseline ;;
    collab) run collab --workload collab --negatives 1 --batch 16384 --no-cpu-baseline ;;
    *) echo "unknown config $cfg" ;;
  esac
done
