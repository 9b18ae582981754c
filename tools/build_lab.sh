#!/bin/bash
# builds the sweep_lab variants (development tool); usage: tools/build_lab.sh
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
rm -f tools/bin/lab_*
NV="nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Ipalabos_b200/csrc tools/sweep_lab.cu"
build() { local nm=$1; shift; $NV "$@" -o tools/bin/lab_$nm & }
for p in 0 1 2 3 4 5; do build pol$p -DLAB_POLICY=$p; done
build pol0pf0 -DPLB_SWEEP_PF_ODD=0
build pol5pf0 -DLAB_POLICY=5 -DPLB_SWEEP_PF_ODD=0
wait
ls tools/bin
