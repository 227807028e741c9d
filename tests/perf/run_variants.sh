#!/bin/bash
# run from the repo root on the GPU box: every th_* variant x 4 shapes x {dense accumulate, fused fp16 + outliers}
run() { desc=$1; bin=$2; sh=$3; sp=$4; shift 4; out=$(env "$@" timeout 120 ./tests/perf/$bin 4 $sh 16 1 $sp 2>&1 | head -1); echo "$desc [sparse=$sp $*] $out"; }
for sp in 0 2; do
for sh in "4096 4096" "4096 12288" "4096 22016" "11008 4096"; do
  run base th_base "$sh" $sp SQLLM_CTAS_PER_SM=3
  run tma14x2 th_tma14 "$sh" $sp SQLLM_CTAS_PER_SM=2 SQLLM_SMEM_BUDGET_KB=110
  run tma14x1 th_tma14 "$sh" $sp SQLLM_CTAS_PER_SM=1 SQLLM_SMEM_BUDGET_KB=110
  run cpa14x2 th_cpa14 "$sh" $sp SQLLM_CTAS_PER_SM=2
  run cpa14x1 th_cpa14 "$sh" $sp SQLLM_CTAS_PER_SM=1
  run ldg14x2 th_ldg14 "$sh" $sp SQLLM_CTAS_PER_SM=2
  run cpa16x1 th_cpa16 "$sh" $sp SQLLM_CTAS_PER_SM=1
  run tma22x1 th_tma22 "$sh" $sp SQLLM_CTAS_PER_SM=1 SQLLM_SMEM_BUDGET_KB=200
  run ldg22x1 th_ldg22 "$sh" $sp SQLLM_CTAS_PER_SM=1
done
done
