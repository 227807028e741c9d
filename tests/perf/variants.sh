run() { desc=$1; bin=$2; sh=$3; ch=$4; gr=$5; shift 5; out=$(env "$@" timeout 60 ./tests/perf/$bin 4 $sh $ch $gr 2>&1); echo "$desc [$*] $(echo "$out" | head -1)"; }
for b in th_spin th_nw8_mb4; do
for f in 1 0; do
run "$b huge flags=$f" $b "4096 65536" 2 0 SQLLM_DEBUG_FLAGS=$f SQLLM_CTAS_PER_SM=2 SQLLM_NO_PDL=1
run "$b chain flags=$f" $b "4096 4096" 16 1 SQLLM_DEBUG_FLAGS=$f
run "$b chain flags=$f" $b "4096 11008" 16 1 SQLLM_DEBUG_FLAGS=$f
done; done
