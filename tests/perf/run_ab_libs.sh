#!/bin/bash
# same-box A/B of two builds: harness binaries $BIN_A / $BIN_B and libraries tests/perf/lib_$LIB_A.so / lib_$LIB_B.so (bench per shape)
set -u
mkdir -p gpurun_out
A=${BIN_A:-th_v2_nobox}; B=${BIN_B:-th_v2}; LA=${LIB_A:-pad}; LB=${LIB_B:-box}
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_zz_exchange_gpu.py tests/test_lut_fp16.py -m gpu -x -q --tb=short 2>&1 | tail -8 > gpurun_out/r02_ab_${LB}_pytest.txt
cat gpurun_out/r02_ab_${LB}_pytest.txt
run() { bin=$1; bits=$2; sh=$3; sp=$4; shift 4; out=$(env "$@" timeout 120 ./tests/perf/$bin $bits $sh 16 1 $sp 2>&1 | head -1); echo "$bin [sparse=$sp $*] $out"; }
for sh in "4096 4096" "4096 12288" "4096 22016" "11008 4096"; do
  for bin in $A $B; do
    run $bin 4 "$sh" 2 SQLLM_X=1
    run $bin 4 "$sh" 2 SQLLM_LUT_MODE=fp16
  done
done
for bin in $A $B; do run $bin 3 "8192 8192" 2 SQLLM_X=1; run $bin 3 "8192 22016" 2 SQLLM_X=1; done
cp squeezellm_b200/libsqllm_b200.so /tmp/lib_orig.so
for v in $LA $LB $LA $LB; do
  cp tests/perf/lib_$v.so squeezellm_b200/libsqllm_b200.so
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --per-shape > gpurun_out/r02_ab_${v}_bench.json 2> gpurun_out/r02_ab_${v}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_ab_${v}_bench.json").read().strip().splitlines()[-1])
ps=d.get("per_shape",{})
print("${v}", round(d["value"],1), "fp16", round(d.get("lut_fp16",{}).get("value",0),1), {m:{k:round(v["us_per_launch"],2) for k,v in ps[m].items()} for m in ps}, d.get("parity_check",{}).get("ok"))
PY
done
cp /tmp/lib_orig.so squeezellm_b200/libsqllm_b200.so
TRACE_ALL=1 timeout 120 ./tests/perf/th_v2_trace 4 4096 4096 16 1 2 > gpurun_out/r02_trace_box_o.txt 2>&1
TRACE_ALL=1 timeout 120 ./tests/perf/th_v2_trace 4 4096 12288 16 1 2 > gpurun_out/r02_trace_box_qkv.txt 2>&1
