#!/bin/bash
# builds the sweep_lab variants (development tool); usage: tools/build_lab.sh
set -e
cd "$(dirname "$0")/.."
rm -f tools/bin/lab_*
NV="nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Ipalabos_b200/csrc tools/sweep_lab.cu"
build() { name=$1; shift; $NV "$@" -o tools/bin/lab_$name & }
build default
build e512pf0_o512pf1 -DPLB_SWEEP_THREADS_ODD=512
build e256pf0_o128pf1 -DPLB_SWEEP_THREADS_EVEN=256 -DPLB_SWEEP_THREADS_ODD=128
build e512pf1_o256pf2 -DPLB_SWEEP_PF_EVEN=1 -DPLB_SWEEP_PF_ODD=2
build f32 -DLAB_T=float
build q27 -DLAB_Q=27
build q27f32 -DLAB_Q=27 -DLAB_T=float
build trt -DLAB_DYN=DYN_TRT
wait
ls tools/bin
