#!/bin/bash
# One gpurun call that establishes the state of the tree: GPU tests, smoke, bench lines.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests.log
tail -5 gpurun_out/gputests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -6 gpurun_out/smoke.log
for wl in C2 C3 C4 C5 C1; do
  extra=""; [ $wl != C2 ] && extra="--no-cpu-baseline --no-e2e"
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-strong $extra > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  echo "$wl rc=$?"; head -c 1500 gpurun_out/bench_$wl.json; echo
done
timeout 300 python tools/nlist_bench.py > gpurun_out/nlist_bench.log 2>&1; tail -12 gpurun_out/nlist_bench.log
