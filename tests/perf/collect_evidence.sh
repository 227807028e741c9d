#!/bin/bash
# Round-2 evidence pass on ONE GPU (run from the repo root under gpurun): every number quoted in DESIGN.md / README.md / profiles/README.md.
# .ncu-rep files are summarised ON THE BOX (raw page -> csv -> tests/perf/ncu_summary.py) and deleted: gpurun_out/ is capped at 64 MiB.
set -x
O=gpurun_out
T="timeout 240"
$T python bench.py --per-shape > $O/r02_bench_llama7b_w4_s45.json 2> $O/r02_bench_llama7b_w4_s45.err
$T python bench.py --launch seq --no-cpu-baseline > $O/r02_bench_llama7b_w4_s45_seq.json 2> $O/r02_bench_llama7b_w4_s45_seq.err
for w in llama7b-w4-s0 llama7b-w3-s45 llama13b-w4-s5 llama65b-w3-s45; do
  timeout 420 python bench.py --workload $w --steps 10 --warmup 3 --per-shape --no-cpu-baseline > $O/r02_bench_${w//-/_}.json 2> $O/r02_bench_${w//-/_}.err
done
$T ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:lutgemv -c 400 --csv \
    --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --lut exact --blocks 1 > $O/r02_launches_bench.log 2>&1
prof() { name=$1; kern=$2; skip=$3; cnt=$4; shift 4
  $T ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c $cnt -f -o $O/$name python bench.py --layers 4 --steps 2 --warmup 3 --blocks 1 --no-cpu-baseline "$@" > $O/$name.log 2>&1
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
  python tests/perf/ncu_summary.py $O/$name.raw.csv > $O/$name.txt 2>&1
  rm -f $O/$name.ncu-rep $O/$name.raw.csv; tail -1 $O/$name.log; }
prof r02_ncu_layer_7b_w4_s45_exact lutgemv2 48 4 --lut exact --launch graph
prof r02_ncu_layer_7b_w4_s45_fp16 lutgemv2 48 4 --lut fp16 --launch graph
prof r02_ncu_layer_7b_w3_s45_exact lutgemv2 48 4 --lut exact --launch graph --workload llama7b-w3-s45
prof r02_ncu_layer_13b_w4_s5_exact lutgemv2 48 4 --lut exact --launch graph --workload llama13b-w4-s5
prof r02_ncu_layer_65b_w3_s45_exact lutgemv2 48 4 --lut exact --launch graph --workload llama65b-w3-s45
prof r02_ncu_seq_7b_w4_s45_exact lutgemv_seq 4 1 --lut exact --launch seq
$T python oracle/ref_gpu_timing.py > $O/r02_per_shape_vs_reference_kernel.jsonl 2> $O/r02_per_shape_vs_reference_kernel.err
$T python oracle/ref_gpu_timing.py --batched > $O/r02_batched_vs_reference_kernel.jsonl 2>> $O/r02_per_shape_vs_reference_kernel.err
ls -la $O | tail -40
