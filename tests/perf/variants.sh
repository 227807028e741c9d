run() { desc=$1; bin=$2; sh=$3; sp=$4; shift 4; out=$(env "$@" timeout 60 ./tests/perf/$bin 4 $sh 16 1 $sp 2>&1); echo "$desc [$sh sparse=$sp $*] $(echo "$out" | head -1)"; }
for b in cpa8 cpa6 ldg6 ldg8 ldg4_mb4; do
  for sp in 0 2; do
    for sh in "4096 4096" "4096 11008" "11008 4096"; do
      run "$b" th_$b "$sh" $sp SQLLM_CTAS_PER_SM=3
    done
  done
done
for b in ldg4_mb4 ldg6; do for sh in "4096 4096" "4096 11008"; do run "$b" th_$b "$sh" 2 SQLLM_CTAS_PER_SM=4; done; done
