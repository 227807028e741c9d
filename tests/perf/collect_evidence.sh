#!/bin/bash
# Round-2 evidence pass on ONE GPU (run from the repo root under gpurun): every number quoted in DESIGN.md / README.md / profiles/README.md.
#   1. bench.py for the five BASELINE workloads (per-shape timing, parity_check, fp16 object); the headline with the CPU baseline
#   2. ncu launch list of the headline command (duration + DRAM bytes per launch)  -> profiles/r02_launches_bench.csv, traffic.json
#   3. ncu --set full of one decoder layer's launches: 7B w4-s45 (exact, fp16), 7B w3-s45, 13B w4-s5, 65B w3-s45
#   4. ours vs the reference's own kernels per shape (batch 1) and the batched symbols
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 5 --per-shape > $O/r02_bench_llama7b_w4_s45.json 2> $O/r02_bench_llama7b_w4_s45.err
for w in llama7b-w4-s0 llama7b-w3-s45 llama13b-w4-s5 llama65b-w3-s45; do
  python bench.py --workload $w --steps 10 --warmup 3 --per-shape --no-cpu-baseline > $O/r02_bench_${w//-/_}.json 2> $O/r02_bench_${w//-/_}.err
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:lutgemv -c 400 --csv \
    --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --lut exact > $O/r02_launches_bench.log 2>&1
prof() { name=$1; shift; ncu --set full --clock-control none --import-source on -k regex:lutgemv2 -s 48 -c 4 -f -o $O/$name python bench.py --layers 4 --steps 2 --warmup 3 --no-cpu-baseline "$@" > $O/$name.log 2>&1; tail -1 $O/$name.log; }
prof r02_ncu_layer_7b_w4_s45_exact --lut exact
prof r02_ncu_layer_7b_w4_s45_fp16 --lut fp16
prof r02_ncu_layer_7b_w3_s45_exact --lut exact --workload llama7b-w3-s45
prof r02_ncu_layer_13b_w4_s5_exact --lut exact --workload llama13b-w4-s5
prof r02_ncu_layer_65b_w3_s45_exact --lut exact --workload llama65b-w3-s45
python oracle/ref_gpu_timing.py > $O/r02_per_shape_vs_reference_kernel.jsonl 2> $O/r02_per_shape_vs_reference_kernel.err
python oracle/ref_gpu_timing.py --batched > $O/r02_batched_vs_reference_kernel.jsonl 2>> $O/r02_per_shape_vs_reference_kernel.err
ls -la $O | tail -30
