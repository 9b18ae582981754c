#!/bin/bash
# runs the built sweep_lab variants (on the GPU box); output to gpurun_out/lab.log
mkdir -p gpurun_out
{
for ge in 64 128 256; do for go in 12 16 20; do timeout 120 tools/bin/lab_default ${1:-512} 6 $ge $go; done; done
for b in e512pf0_o512pf1 e256pf0_o128pf1 e512pf1_o256pf2; do for go in 8 16 32; do timeout 120 tools/bin/lab_$b ${1:-512} 6 64 $go; done; done
for b in f32 q27 q27f32 trt; do timeout 120 tools/bin/lab_$b ${1:-512} 6 64 16;  timeout 120 tools/bin/lab_$b ${1:-512} 6 16 8; done
} 2>&1 | tee gpurun_out/lab.log
