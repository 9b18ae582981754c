#!/bin/bash
# builds the sweep_lab variants (development tool); usage: tools/build_lab.sh
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
rm -f tools/bin/lab_*
NV="nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Ipalabos_b200/csrc tools/sweep_lab.cu"
build() { local nm=$1; shift; $NV "$@" -o tools/bin/lab_$nm & }
for cfg in "q27 -DLAB_Q=27" "q27f32 -DLAB_Q=27 -DLAB_T=float" "f32 -DLAB_T=float" "q27smago -DLAB_Q=27 -DLAB_DYN=DYN_SMAGO"; do
  set -- $cfg; name=$1; shift
  build ${name}_t512r128 "$@"
  build ${name}_t384r168 "$@" -DPLB_SWEEP_THREADS_EVEN=384 -DPLB_SWEEP_THREADS_ODD=384 -DPLB_SWEEP_REGS=168
  build ${name}_t256r255 "$@" -DPLB_SWEEP_THREADS_EVEN=256 -DPLB_SWEEP_THREADS_ODD=256 -DPLB_SWEEP_REGS=255
  build ${name}_t256r128 "$@" -DPLB_SWEEP_THREADS_EVEN=256 -DPLB_SWEEP_THREADS_ODD=256
  wait
done
ls tools/bin
