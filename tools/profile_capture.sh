#!/bin/bash
# Run ON THE GPU BOX (under gpurun): one `ncu --set full` capture of the force kernels of a
# workload at BASELINE size -- exactly the launches of bench.py's timed region (autotuned launch
# shape; the region is bracketed with cudaProfilerStart/Stop) -- plus the hash of the kernel
# sources the capture belongs to.
#   tools/profile_capture.sh C4 r02_c4 [extra bench.py flags]
# -> gpurun_out/<tag>.ncu-rep, gpurun_out/<tag>.srchash, gpurun_out/<tag>.sasshash, gpurun_out/<tag>_raw.csv
set -e
WL=$1; TAG=$2; shift 2
mkdir -p gpurun_out
python tools/srchash.py > gpurun_out/${TAG}.srchash
python tools/srchash.py --sass ${WL} > gpurun_out/${TAG}.sasshash
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:row_kernel -f -o gpurun_out/${TAG} \
    python bench.py --workload ${WL} --steps 1 --warmup 3 --no-cpu-baseline --no-strong --no-e2e \
    --cuda-profiler "$@" > gpurun_out/${TAG}.log 2>&1 || { tail -20 gpurun_out/${TAG}.log; exit 1; }
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv
