#!/bin/bash
# same-box A/B of harness binaries (v1, previous build, current build); every launch has its own weights, LUT and outlier arrays
run() { bin=$1; bits=$2; sh=$3; sp=$4; shift 4; out=$(env "$@" timeout 120 ./tests/perf/$bin $bits $sh 16 1 $sp 2>&1 | head -1); echo "$bin [sparse=$sp $*] $out"; }
for sh in "4096 4096" "4096 12288" "4096 22016" "11008 4096"; do
  run th_v2 4 "$sh" 2 SQLLM_KERNEL=v1 SQLLM_CTAS_PER_SM=3
  for bin in th_v2_prev th_v2; do
    run $bin 4 "$sh" 2 SQLLM_X=1
    run $bin 4 "$sh" 2 SQLLM_LUT_MODE=fp16
  done
done
