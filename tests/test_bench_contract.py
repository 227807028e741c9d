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


def test_sequence_wiring_matches_the_launch_by_launch_step():
    """bench.seq_record_layer must wire the 7 matvecs of a decoder layer exactly as make_step's run_layer does (x -> q,k,v ; v -> o ;
    o -> gate,up ; gate -> down), with and without sibling stacking, on one and on several ranks (full-length vectors)."""
    import bench
    from squeezellm_b200.runtime import SeqVec

    class FakeLayer:
        def __init__(self, K, N, group=None):
            self.infeatures, self.outfeatures, self._sibling_group = K, N, group

    class FakeSeq:
        def __init__(self, world):
            self.world, self.items = world, []

        def matvec(self, layer, x, members=1):
            assert len(x) == layer.infeatures
            n = layer.outfeatures // members * self.world * members
            self.items.append((layer, x.item, x.offset, members))
            return SeqVec(len(self.items) - 1, 0, n)

    H, F = 4096, 11008
    for world in (1, 4):
        w = lambda n: n // world
        for stacked in (False, True):
            L = {n: FakeLayer(H, w(H)) for n in ("q_proj", "k_proj", "v_proj", "o_proj")}
            L.update(gate_proj=FakeLayer(H, w(F)), up_proj=FakeLayer(H, w(F)), down_proj=FakeLayer(F, w(H)))
            if stacked:
                class G:  # what fusion.SiblingGroup exposes
                    pass
                gq, gg = G(), G()
                gq.layer, gg.layer = FakeLayer(H, 3 * w(H)), FakeLayer(H, 2 * w(F))
                for n in ("q_proj", "k_proj", "v_proj"):
                    L[n]._sibling_group = (gq, 0)
                for n in ("gate_proj", "up_proj"):
                    L[n]._sibling_group = (gg, 0)
            seq = FakeSeq(world)
            rec = bench.seq_record_layer(seq, L, SeqVec(-1, 0, H), world)
            names = [r[0] for r in rec]
            assert names == ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]
            by = {r[0]: r for r in rec}
            assert all(len(by[n][2]) == H for n in ("q_proj", "k_proj", "v_proj", "o_proj", "down_proj"))
            assert len(by["gate_proj"][2]) == F and len(by["up_proj"][2]) == F
            # o reads v, gate/up read o, down reads gate - as (item, offset) of full-length vectors
            v, o, g = by["v_proj"][2], by["o_proj"][2], by["gate_proj"][2]
            assert (by["o_proj"][1].item, by["o_proj"][1].offset) == (v.item, v.offset)
            assert (by["gate_proj"][1].item, by["gate_proj"][1].offset) == (o.item, o.offset) == (by["up_proj"][1].item, by["up_proj"][1].offset)
            assert (by["down_proj"][1].item, by["down_proj"][1].offset) == (g.item, g.offset)
            assert len(seq.items) == (4 if stacked else 7)
            if stacked:
                assert v.offset == 2 * H and by["up_proj"][2].offset == F
