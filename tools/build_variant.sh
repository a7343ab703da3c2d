#!/bin/bash
# Build a library variant for A/B timing: tools/build_variant.sh <name> <nvcc -D flags...>
# -> scratch/libazp_<name>.so (scratch/ is git-ignored; it travels to the GPU box)
set -e
NAME=$1; shift
ROOT=$(cd $(dirname $0)/.. && pwd)
mkdir -p $ROOT/scratch
make -s -C $ROOT/azplugins_b200/csrc -j8 OUT=$ROOT/scratch/libazp_$NAME.so BUILD=$ROOT/scratch/build_$NAME EXTRA_NVCCFLAGS="$*" > $ROOT/scratch/build_$NAME.log 2>&1 || { tail -20 $ROOT/scratch/build_$NAME.log; exit 1; }
echo "built scratch/libazp_$NAME.so ($*)"
