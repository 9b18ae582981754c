#!/bin/bash
# builds the sweep_lab variants (development tool); usage: tools/build_lab.sh
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
rm -f tools/bin/lab_*
NV="nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -Ipalabos_b200/csrc tools/sweep_lab.cu"
build() { local nm=$1; shift; $NV "$@" -o tools/bin/lab_$nm & }
build f32q27 -DLAB_T=float -DLAB_Q=27 -DLAB_DYN=DYN_SMAGO
build f32q19 -DLAB_T=float -DLAB_Q=19 -DLAB_DYN=DYN_BGK
build f64q27 -DLAB_T=double -DLAB_Q=27 -DLAB_DYN=DYN_SMAGO
build f64q19 -DLAB_T=double -DLAB_Q=19 -DLAB_DYN=DYN_BGK
wait
ls tools/bin
