run() { desc=$1; bin=$2; sh=$3; sp=$4; shift 4; out=$(env "$@" timeout 60 ./tests/perf/$bin 4 $sh 16 1 $sp 2>&1); echo "$desc [$sh sparse=$sp $*] $(echo "$out" | head -1)"; }
for sh in "4096 4096" "4096 12288" "4096 22016" "11008 4096"; do
  run nfin1024/16 th_cpa8 "$sh" 2 SQLLM_CTAS_PER_SM=3
  run nfin256/16 th_cpa8 "$sh" 2 SQLLM_CTAS_PER_SM=3 SQLLM_NFIN_COLS=256
  run nfin256/32 th_cpa8 "$sh" 2 SQLLM_CTAS_PER_SM=3 SQLLM_NFIN_COLS=256 SQLLM_NFIN_MAX=32
  run nfin512/48 th_cpa8 "$sh" 2 SQLLM_CTAS_PER_SM=3 SQLLM_NFIN_COLS=512 SQLLM_NFIN_MAX=48
  run pf6 th_cpa6 "$sh" 2 SQLLM_CTAS_PER_SM=3
  run pf12 th_cpa12 "$sh" 2 SQLLM_CTAS_PER_SM=3
done
