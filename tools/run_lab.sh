#!/bin/bash
# runs the built sweep_lab variants (on the GPU box); output to gpurun_out/lab.log
mkdir -p gpurun_out
{
for b in tools/bin/lab_*; do for g in 64 128; do echo -n "$(basename $b) "; timeout 120 $b ${1:-512} 5 64 $g; done; done
} 2>&1 | tee gpurun_out/lab.log
