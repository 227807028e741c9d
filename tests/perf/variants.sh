run() { desc=$1; bin=$2; sh=$3; ch=$4; gr=$5; shift 5; out=$(env "$@" timeout 60 ./tests/perf/$bin 4 $sh $ch $gr 2>&1); echo "$desc [$*] $(echo "$out" | head -1)"; }
for b in cpa8 cpa8_mb3; do for c in 2 3 4; do
run "$b" th_$b "4096 4096" 16 1 SQLLM_CTAS_PER_SM=$c
run "$b" th_$b "4096 11008" 16 1 SQLLM_CTAS_PER_SM=$c
done; done
run "dbg flags1" th_cpa8_dbg "4096 4096" 16 1 SQLLM_DEBUG_FLAGS=1
run "dbg flags1" th_cpa8_dbg "4096 11008" 16 1 SQLLM_DEBUG_FLAGS=1
