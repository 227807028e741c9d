#!/bin/bash
# A/B: entry L2 prefetch of the CTA's chunk (SQLLM_L2PF=1) vs none; every launch has its own weights (>= 400 MB rotation)
run() { bin=$1; bits=$2; sh=$3; sp=$4; shift 4; out=$(env "$@" timeout 120 ./tests/perf/$bin $bits $sh 16 1 $sp 2>&1 | head -1); echo "$bin [sparse=$sp $*] $out"; }
for bits in 4 3; do
for sh in "4096 4096" "4096 12288" "4096 22016" "11008 4096"; do
  for pf in 0 1; do
    run th_v2 $bits "$sh" 2 SQLLM_L2PF=$pf
    run th_v2 $bits "$sh" 2 SQLLM_L2PF=$pf SQLLM_LUT_MODE=fp16
  done
done
done
