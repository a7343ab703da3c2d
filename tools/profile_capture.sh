#!/bin/bash
# Run ON THE GPU BOX (under gpurun): one `ncu --set full` capture of the force kernels of a
# workload at BASELINE size plus the source hash the capture belongs to.
#   tools/profile_capture.sh C4 r02_c4 [extra bench.py flags]
# -> gpurun_out/<tag>.ncu-rep, gpurun_out/<tag>.srchash, gpurun_out/<tag>_raw.csv
# The kernels of the timed step are the LAST launches of the process: warm-up 3, 1 step, no CPU
# baseline, no strong-scaling leg, no autotuning sweep (--no-tune uses the library default shape
# unless --block/--tpp are given); -k filters on the kernel name, --launch-skip drops the warm-up.
set -e
WL=$1; TAG=$2; shift 2
mkdir -p gpurun_out
python tools/srchash.py > gpurun_out/${TAG}.srchash
ncu --set full --clock-control none --import-source on -k regex:row_kernel \
    --launch-skip ${SKIP:-6} --launch-count ${COUNT:-2} -f -o gpurun_out/${TAG} \
    python bench.py --workload ${WL} --steps 1 --warmup 3 --no-cpu-baseline --no-strong --no-e2e "$@" \
    > gpurun_out/${TAG}.log 2>&1 || { tail -20 gpurun_out/${TAG}.log; exit 1; }
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv
