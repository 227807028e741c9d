#!/bin/bash
# ncu --set full on one steady-state launch of the v2 kernel via the C-ABI harness (no graph, chain of 4)
prof() { name=$1; shift; bits=$1; K=$2; N=$3; sp=$4; shift 4
  env "$@" ncu --set full --clock-control none --import-source on -k regex:lutgemv -s 6 -c 1 -f -o gpurun_out/$name ./tests/perf/th_v2 $bits $K $N 4 0 $sp > gpurun_out/$name.log 2>&1
  tail -1 gpurun_out/$name.log; }
for spec in "$@"; do
  set -- $spec
  prof $1 $2 $3 $4 $5 $6
done
