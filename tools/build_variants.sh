#!/bin/bash
# development tool: builds copies of the library with different -D flags under tools/bin/<name>/ (f64/D3Q19 only;
# the other three instantiations are stubbed) so that one gpurun call can compare them with bench.py.
# usage: tools/build_variants.sh name1 "flags1" name2 "flags2" ...
set -e
cd "$(dirname "$0")/.."
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  d=tools/bin/$name
  rm -rf $d; mkdir -p $d/palabos_b200 $d/include
  cp -r palabos_b200/csrc palabos_b200/__init__.py $d/palabos_b200/
  rm -rf $d/palabos_b200/csrc/build
  cp -r include/* $d/include/
  cp bench.py MEASURED_PEAKS.json $d/ 2>/dev/null || true
  cp -r tests $d/tests; mkdir -p $d/oracle; cp oracle/liboracle.so oracle/Makefile oracle/plb_oracle* $d/oracle/
  for v in f64_q27 f32_q19 f32_q27; do
    printf '#include "engine.h"\nnamespace plb200 { LatticeBase* make_lattice_%s(const CreateDesc&) { throw PlbError("not in this variant build"); } }\n' $v > $d/palabos_b200/csrc/inst_$v.cu
  done
  ( make -C $d/palabos_b200/csrc -j2 EXTRA="$flags" > $d/build.log 2>&1 && echo "built $name" || echo "FAILED $name" ) &
done
wait
