run() { desc=$1; bin=$2; sh=$3; sp=$4; shift 4; out=$(env "$@" timeout 60 ./tests/perf/$bin 4 $sh 16 1 $sp 2>&1); echo "$desc [$sh sparse=$sp $*] $(echo "$out" | head -1)"; }
for sp in 2; do
  for sh in "4096 4096" "4096 12288" "4096 22016" "11008 4096"; do
    run cpa8 th_cpa8 "$sh" $sp SQLLM_CTAS_PER_SM=3
  done
done
