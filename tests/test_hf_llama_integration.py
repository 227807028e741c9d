"""Drop-in check against the model code llama.py actually drives: a (tiny, random-init) transformers LlamaForCausalLM gets its
Linears replaced from a reference-format checkpoint (`checkpoint.load_quantized`), siblings stacked (`fusion.fuse_siblings`),
and its own forward() - unmodified HF code calling q_proj(x), k_proj(x), v_proj(x), gate_proj(x), up_proj(x) - must give the
same logits with and without stacking, with ONE stacked launch per sibling set.  CPU only: the two `quant_cuda` entry points
the module calls are replaced by the fp64 oracle, everything above them (QuantLinearLUT.forward, the sibling cache, reshapes)
is the product code."""
import types

import numpy as np
import pytest
import torch

from util import orc
from squeezellm_b200 import checkpoint as ck, fusion, quant as Q

transformers = pytest.importorskip("transformers")


def _layer_of(qweight, lut, bits, bias=None, rows=None, cols=None, vals=None, fr=None, fri=None):
    d = dict(bits=bits, infeatures=qweight.shape[0] // bits * 32, outfeatures=qweight.shape[1], qweight=qweight.contiguous().numpy(),
             lookup_table=lut.contiguous().numpy(), bias=None)
    for k, v in (("rows", rows), ("cols", cols), ("vals", vals), ("full_rows", fr), ("full_row_indices", fri)):
        d[k] = v.contiguous().numpy() if v is not None else None
    return d


class FakeQuantCuda(types.SimpleNamespace):
    """Oracle-backed stand-ins for the two symbols QuantLinearLUT.forward reaches in this test."""
    calls = 0

    def lutgemv_fused(self, x, qweight, lut, bits, bias, rows, cols, vals, fr, fri):
        FakeQuantCuda.calls += 1
        y = orc.forward_f64(_layer_of(qweight, lut, bits, None, rows, cols, vals, fr, fri), x.float().numpy().reshape(1, -1))[0]
        y = torch.from_numpy(y.astype(np.float32))
        return (y + bias if bias is not None else y).to(x.dtype)

    def vecquant4matmul_nuq_perchannel_batched(self, x, qweight, y, lut):
        FakeQuantCuda.calls += 1
        y += torch.from_numpy(orc.forward_f64(_layer_of(qweight, lut, 4), x.numpy()).astype(np.float32))


@pytest.fixture()
def tiny_llama(monkeypatch):
    cfg = transformers.LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=4, vocab_size=97, max_position_embeddings=32)
    torch.manual_seed(0)
    model = transformers.LlamaForCausalLM(cfg).eval()
    state = {k: v.clone() for k, v in model.state_dict().items()}
    n = 0
    for name, lin in ck.find_linear_layers(model).items():
        if name == "lm_head":
            continue
        L = orc.make_layer(4, lin.in_features, lin.out_features, seed=100 + n)
        n += 1
        del state[name + ".weight"]
        state[name + ".qweight"] = torch.from_numpy(L["qweight"])
        state[name + ".lookup_table"] = torch.from_numpy(L["lookup_table"] * 5.0)   # keep activations O(1) through the blocks
    res = ck.load_quantized(model, state, 4)
    assert not res.missing_keys and not res.unexpected_keys and n == 14
    monkeypatch.setattr(Q, "quant_cuda", FakeQuantCuda())
    return model


@pytest.mark.parametrize("seq", [1, 3], ids=["decode-shaped", "prefill-shaped"])
def test_hf_llama_forward_is_unchanged_by_sibling_stacking(tiny_llama, seq):
    model = tiny_llama
    ids = torch.tensor([[5, 17, 42][:seq]])
    with torch.no_grad():
        FakeQuantCuda.calls = 0
        before = model(ids).logits
        calls_unfused = FakeQuantCuda.calls
        groups = fusion.fuse_siblings(model)
        FakeQuantCuda.calls = 0
        after = model(ids).logits
    assert torch.isfinite(before).all() and before.abs().max() > 0
    assert len(groups) == 4 and sorted(len(g.members) for g in groups) == [2, 2, 3, 3]
    assert calls_unfused == 14 and FakeQuantCuda.calls == 8, "7 launches per block become 4"
    assert all(g.launches == 1 for g in groups)
    assert torch.allclose(before, after, rtol=1e-5, atol=1e-6)


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker_sharded_llama(rank, world, port, q):
    import os
    import torch.distributed as dist
    from squeezellm_b200.sharding import shard_model
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = transformers.LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                       num_key_value_heads=4, vocab_size=97, max_position_embeddings=32)
        torch.manual_seed(0)
        model = transformers.LlamaForCausalLM(cfg).eval()
        state = {k: v.clone() for k, v in model.state_dict().items()}
        for n, (name, lin) in enumerate((k, v) for k, v in ck.find_linear_layers(model).items() if k != "lm_head"):
            L = orc.make_layer(4, lin.in_features, lin.out_features, seed=100 + n)
            del state[name + ".weight"]
            state[name + ".qweight"] = torch.from_numpy(L["qweight"])
            state[name + ".lookup_table"] = torch.from_numpy(L["lookup_table"] * 5.0)
        ck.load_quantized(model, state, 4)
        Q.quant_cuda = FakeQuantCuda()
        out = []
        with torch.no_grad():
            for ids in (torch.tensor([[5]]), torch.tensor([[5, 17, 42]])):
                out.append(model(ids).logits)
            groups = shard_model(model, rank, world)
            for i, ids in enumerate((torch.tensor([[5]]), torch.tensor([[5, 17, 42]]))):
                out.append(model(ids).logits)
        same = all(torch.allclose(out[i], out[i + 2], rtol=1e-5, atol=1e-6) for i in range(2))
        q.put((rank, same, len(groups), sorted({g.w for g in groups}), model.model.layers[0].self_attn.q_proj.qweight.shape[1]))
    finally:
        dist.destroy_process_group()


def test_hf_llama_forward_is_unchanged_by_shard_model_world2_gloo():
    """shard_model on the real HF module tree: every rank keeps half of every QuantLinearLUT's columns, runs stacked shards and one
    exchange per launch (gloo all-reduce here), and the unmodified forward still produces the unsharded logits on every rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded_llama, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, ngroups, widths, qcols in res:
        assert same, f"rank {rank}: sharded logits differ"
        assert ngroups == 8 and widths == [32, 64] and qcols == 32   # per block: q/k/v, o, gate/up, down; hidden 64 -> 32, ffn 128 -> 64 per rank
