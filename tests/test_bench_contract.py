"""bench.py contract checks that need no GPU: the reference arm (CPU restatement timed on host cores) prints exactly one
JSON line with the keys the driver reads, and both arms build metric / config from the same helpers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["workload"] == "llama7b-w4-s45" and d["config"]["matvecs_per_step"] == 224
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_both_arms_share_metric_and_config():
    sys.path.insert(0, ROOT)
    import bench

    class A:
        workload, layers = "llama7b-w4-s45", 0
    cfg = dict(bench.WORKLOADS["llama7b-w4-s45"])
    assert bench.metric_name("llama7b-w4-s45").startswith("LLaMA-7B w4-s45 decode tokens/s at batch=1")
    c = bench.base_config(A, cfg)
    assert c["matvecs_per_step"] == 7 * cfg["layers"] == 224 and c["bits"] == 4 and c["layers_overridden"] is False
    # algorithmic bytes of the headline workload: 3.62 GB per token
    total = 0
    for name, kin, kout in bench.MATS:
        K, N = cfg[kin], cfg[kout]
        nnz = int(round(cfg["sparsity"] * K * N))
        total += bench.alg_bytes(cfg["bits"], K, N, nnz, cfg["topX"])
    assert abs(total * cfg["layers"] - 3.619e9) < 5e6
