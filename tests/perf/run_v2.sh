#!/bin/bash
# A/B on the GPU box: v1 vs v2 (exact / fp16 pair table), chains of 16 launches in a CUDA graph through the C ABI
run() { desc=$1; bits=$2; sh=$3; sp=$4; shift 4; out=$(env "$@" timeout 120 ./tests/perf/th_v2 $bits $sh 16 1 $sp 2>&1 | head -1); echo "$desc [sparse=$sp $*] $out"; }
for bits in ${BITS:-4 3}; do
for sp in 0 2; do
for sh in "4096 4096" "4096 12288" "4096 22016" "11008 4096"; do
  [ -z "$SKIP_V1" ] && run v1 $bits "$sh" $sp SQLLM_KERNEL=v1 SQLLM_CTAS_PER_SM=3
  run v2-exact $bits "$sh" $sp SQLLM_X=1
  [ -n "$PFD" ] && run v2-exact-pfd $bits "$sh" $sp SQLLM_PFD=$PFD
  if [ $sp = 2 ]; then run v2-fp16 $bits "$sh" $sp SQLLM_LUT_MODE=fp16; [ -n "$PFD" ] && run v2-fp16-pfd $bits "$sh" $sp SQLLM_LUT_MODE=fp16 SQLLM_PFD=$PFD; fi
done
done
done
