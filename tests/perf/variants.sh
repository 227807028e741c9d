run() { desc=$1; bin=$2; sh=$3; ch=$4; gr=$5; shift 5; out=$(env "$@" timeout 60 ./tests/perf/$bin 4 $sh $ch $gr 2>&1); echo "$desc [$*] $(echo "$out" | head -1)"; }
for b in cpa4 cpa8 cpa12; do for c in 2 3 4; do
run "$b" th_$b "4096 4096" 16 1 SQLLM_CTAS_PER_SM=$c
run "$b" th_$b "4096 11008" 16 1 SQLLM_CTAS_PER_SM=$c
done; done
run "cpa8 flags1" th_cpa8 "4096 4096" 16 1 SQLLM_DEBUG_FLAGS=1
run "cpa8 flags1" th_cpa8 "4096 11008" 16 1 SQLLM_DEBUG_FLAGS=1
run "cpa8 nopdl" th_cpa8 "4096 4096" 16 1 SQLLM_NO_PDL=1
run "cpa8 huge" th_cpa8 "4096 65536" 2 0 SQLLM_NO_PDL=1
run "cpa8 huge flags1" th_cpa8 "4096 65536" 2 0 SQLLM_NO_PDL=1 SQLLM_DEBUG_FLAGS=1
run "cpa8 huge cps4" th_cpa8 "4096 65536" 2 0 SQLLM_NO_PDL=1 SQLLM_CTAS_PER_SM=4
run "cpa12 huge flags1" th_cpa12 "4096 65536" 2 0 SQLLM_NO_PDL=1 SQLLM_DEBUG_FLAGS=1
