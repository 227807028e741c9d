#!/bin/bash
# same-box A/B of the two outlier schedules (SQLLM_CSR_LOCAL=0 balanced / 1 owner-local): parity tests, then bench per shape with each library
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_exchange_gpu.py tests/test_lut_fp16.py -m gpu -x -q --tb=short 2>&1 | tail -5 > gpurun_out/r02_ab_csr_pytest.txt
cat gpurun_out/r02_ab_csr_pytest.txt
cp squeezellm_b200/libsqllm_b200.so /tmp/lib_orig.so
for v in 0 1 0 1; do
  cp tests/perf/lib_csr$v.so squeezellm_b200/libsqllm_b200.so
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --per-shape > gpurun_out/r02_ab_csr${v}_bench.json 2> gpurun_out/r02_ab_csr${v}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_ab_csr${v}_bench.json").read().strip().splitlines()[-1])
ps=d.get("per_shape",{})
print("csr_local=${v}", round(d["value"],1), "fp16", round(d.get("lut_fp16",{}).get("value",0),1), {m:{k:round(v["us_per_launch"],2) for k,v in ps[m].items()} for m in ps})
PY
done
cp /tmp/lib_orig.so squeezellm_b200/libsqllm_b200.so
