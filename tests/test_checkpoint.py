"""Checkpoint glue (squeezellm_b200/checkpoint.py): a reference-format state dict (buffers + sparse_threshold.* ints) loads into
a model whose Linears are replaced on the fly, and saving gives the same dict back.  CPU only."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from util import orc
from squeezellm_b200 import checkpoint as ck
from squeezellm_b200.quant import QuantLinearLUT


class Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.q_proj, self.k_proj = nn.Linear(64, 64, bias=False), nn.Linear(64, 64, bias=False)
        self.down_proj = nn.Linear(128, 64, bias=True)


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.layers = nn.ModuleList([Block(), Block()])
        self.norm = nn.LayerNorm(64)
        self.lm_head = nn.Linear(64, 32, bias=False)


def _reference_style_state(bits, sparse, topx):
    """What pack.py would have saved for Net: per quantized layer the buffer set, thresholds, and the untouched tensors."""
    net = Net()
    state = {k: v.clone() for k, v in net.state_dict().items() if "proj" not in k}
    expect = {}
    for name, lin in ck.find_linear_layers(net).items():
        if name == "lm_head":
            continue
        K, N = lin.in_features, lin.out_features
        L = orc.make_layer(bits, K, N, sparsity=0.03 if sparse else 0.0, topX=topx if sparse else 0, seed=len(expect), nonzero_full_rows=True,
                           bias=lin.bias is not None)
        for key in ("qweight", "lookup_table", "bias", "rows", "cols", "vals", "full_rows", "full_row_indices"):
            if L.get(key) is not None:
                state[f"{name}.{key}"] = torch.from_numpy(L[key])
        if sparse:
            state[ck.PREFIX + name] = len(L["vals"])
        expect[name] = L
    return state, expect


@pytest.mark.parametrize("bits,sparse,topx", [(4, False, 0), (3, True, 0), (4, True, 5)], ids=["w4-dense", "w3-sparse", "w4-hybrid"])
def test_load_reference_style_checkpoint_and_save_it_back(bits, sparse, topx):
    state, expect = _reference_style_state(bits, sparse, topx)
    before = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in state.items()}
    net = Net()
    res = ck.load_quantized(net, state, bits, include_sparse=sparse, topX=topx)
    assert not res.unexpected_keys and not res.missing_keys
    assert set(state) == set(before), "the caller's dict is left alone (llama.py deletes the thresholds from its own copy)"
    assert isinstance(net.lm_head, nn.Linear) and isinstance(net.norm, nn.LayerNorm)
    for name, L in expect.items():
        m = net.get_submodule(name)
        assert isinstance(m, QuantLinearLUT) and (m.bits, m.infeatures, m.outfeatures) == (bits, L["infeatures"], L["outfeatures"])
        for key in ("qweight", "lookup_table", "bias", "rows", "cols", "vals", "full_rows", "full_row_indices"):
            if L.get(key) is not None:
                assert np.array_equal(getattr(m, key).numpy(), L[key]), (name, key)
        assert m.topX == (topx if sparse else 0)
    saved = ck.quantized_state_dict(net)
    assert set(saved) == set(before)
    for k, v in before.items():
        assert torch.equal(saved[k], v) if torch.is_tensor(v) else saved[k] == v, k


def test_sparse_checkpoint_without_thresholds_is_rejected_and_missing_dense_rows_stay_zero():
    state, _ = _reference_style_state(4, True, 0)
    clean, numvals = ck.split_sparse_thresholds(state)
    assert all(isinstance(v, int) for v in numvals.values()) and len(numvals) == 6 and not any(k.startswith(ck.PREFIX) for k in clean)
    with pytest.raises(KeyError, match="sparse_threshold"):
        ck.load_quantized(Net(), clean, 4, include_sparse=True, topX=0)
    # the reference's released checkpoints carry no full_rows: topX buffers exist (llama.py passes topX=10) and stay zero
    net = Net()
    res = ck.load_quantized(net, state, 4, include_sparse=True, topX=10)
    assert sorted({k.rsplit(".", 1)[1] for k in res.missing_keys}) == ["full_row_indices", "full_rows"]
    q = net.layers[0].q_proj
    assert q.full_rows.shape == (64, 10) and not q.full_rows.any() and not q.full_row_indices.any()


@pytest.mark.parametrize("ext", ["pt", "safetensors"])
@pytest.mark.parametrize("bits,sparse,topx", [(4, False, 0), (3, True, 5)], ids=["w4-dense", "w3-hybrid"])
def test_checkpoint_files_roundtrip_with_quant_config_sidecar(tmp_path, ext, bits, sparse, topx):
    """File level: the reference's torch.save pickle and the .safetensors variant, both with the quant_config.json sidecar
    (quantization/pack.py:184-190); wbits / include_sparse are recovered from the files when not given."""
    import json
    state, expect = _reference_style_state(bits, sparse, topx)
    net = Net()
    ck.load_quantized(net, state, bits, include_sparse=sparse, topX=topx)
    path = str(tmp_path / f"model.{ext}")
    ck.save_checkpoint(net, path, bits)
    assert json.load(open(tmp_path / "quant_config.json")) == {"wbits": bits}
    assert ck.read_quant_config(path) == {"wbits": bits} and ck.read_quant_config(str(tmp_path)) == {"wbits": bits}
    net2 = Net()
    res = ck.load_checkpoint(net2, path, topX=topx)  # wbits from the sidecar, include_sparse from the thresholds in the file
    assert not res.unexpected_keys and not res.missing_keys
    a, b = ck.quantized_state_dict(net), ck.quantized_state_dict(net2)
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k], b[k]) if torch.is_tensor(a[k]) else a[k] == b[k], k
    if ext == "pt":  # the pickle is exactly what pack.py writes: a plain dict with int thresholds
        raw = torch.load(path, weights_only=False)
        assert all(isinstance(raw[k], int) for k in raw if k.startswith(ck.PREFIX)) and (len([k for k in raw if k.startswith(ck.PREFIX)]) == (6 if sparse else 0))
    os.remove(tmp_path / "quant_config.json")
    if ext == "safetensors":   # metadata carries wbits too
        ck.load_checkpoint(Net(), path, topX=topx)
    else:
        with pytest.raises(ValueError, match="wbits"):
            ck.load_checkpoint(Net(), path, topX=topx)
