#!/bin/bash
# After a gpurun call that ran tools/profile_capture.sh <WL> <TAG>_<WL> for the workloads given:
#   tools/register_captures.sh r02b C2 C3 C4 C5
# condenses gpurun_out/<TAG>_<WL>_raw.csv into profiles/<TAG>_<WL>_ncu_full_summary.csv and
# registers DRAM bytes / warp instructions (with the kernel-source hash of the capture) in
# profiles/ncu_constants.json, which bench.py reads at run time.
set -e
TAG=$1; shift
declare -A N=( [C1]=32000 [C2]=1000000 [C3]=3999766 [C4]=8000000 [C5]=16000000 )
for WL in "$@"; do
  python tools/ncu_summary.py gpurun_out/${TAG}_${WL}_raw.csv --register $WL --n ${N[$WL]} \
      --summary profiles/${TAG}_${WL}_ncu_full_summary.csv --hash $(cat gpurun_out/${TAG}_${WL}.srchash) \
      --sass-hash "$(cat gpurun_out/${TAG}_${WL}.sasshash 2>/dev/null)"
done
