#!/bin/bash
# runs the built sweep_lab variants over grid factors (on the GPU box); output to gpurun_out/lab.log
mkdir -p gpurun_out
{
for b in tools/bin/lab_*; do for g in ${2:-8 16 32 64 128 256}; do echo -n "$(basename $b) gf $g: "; timeout 120 $b ${1:-512} 5 $g $g | sed 's/.*| even \([0-9.]*\) ms \([0-9]*\) GB.s | odd \([0-9.]*\) ms \([0-9]*\) GB.s | avg \([0-9]*\) GB.s \([0-9]*\) MLUPS/even \1 ms \2 GB\/s  odd \3 ms \4 GB\/s  avg \5 GB\/s  \6 MLUPS/'; done; done
} 2>&1 | tee gpurun_out/lab.log
