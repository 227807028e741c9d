"""Loading / saving the reference's packed checkpoints into QuantLinearLUT modules.

The on-disk format IS the module's buffer set: a torch state dict whose keys are
`<layer>.qweight`, `<layer>.lookup_table`, optionally `<layer>.bias`, `<layer>.rows/.cols/.vals`,
`<layer>.full_rows/.full_row_indices`, next to the untouched fp16 tensors of the model, plus one python int per sparse
layer under `sparse_threshold.<layer>` = its CSR nnz (written by quantization/pack.py:173-178, read and deleted by
llama.py:158-167 because the CSR buffers must be allocated with the right length before `load_state_dict`).

This module is the glue the reference keeps inline in llama.py:136-185 (`load_quant`), minus model construction:
    state = torch.load(path)                       # reference checkpoint, unchanged
    load_quantized(model, state, wbits=4, include_sparse=True, topX=10)
    model.cuda(); fuse_siblings(model)             # optional: squeezellm_b200.fusion
`save_checkpoint` / `load_checkpoint` add the file level: the reference's `torch.save` pickle + `quant_config.json` sidecar
({"wbits": N}, quantization/pack.py:184-190) or - same tensors, same names - a `.safetensors` file whose metadata carries the
sidecar's content and the `sparse_threshold.*` integers (safetensors stores tensors only).
Nothing here touches the GPU.
"""
import json
import os

import torch
import torch.nn as nn

from .quant import QuantLinearLUT, make_quant_lut

__all__ = ["find_linear_layers", "split_sparse_thresholds", "merge_sparse_thresholds", "load_quantized", "quantized_state_dict",
           "write_quant_config", "read_quant_config", "save_checkpoint", "load_checkpoint"]

PREFIX = "sparse_threshold."


def find_linear_layers(module, kinds=(nn.Linear,), prefix=""):
    """{qualified name: module} of every `kinds` instance below `module` (what llama.py gets from modelutils.find_layers)."""
    found = {}
    for name, child in module.named_children():
        full = f"{prefix}.{name}" if prefix else name
        if isinstance(child, kinds):
            found[full] = child
        else:
            found.update(find_linear_layers(child, kinds, full))
    return found


def split_sparse_thresholds(state):
    """-> (state without the `sparse_threshold.*` entries, {layer name: nnz}).  The input dict is not modified."""
    clean, numvals = {}, {}
    for k, v in state.items():
        if k.startswith(PREFIX):
            numvals[k[len(PREFIX):]] = int(v)
        else:
            clean[k] = v
    return clean, numvals


def merge_sparse_thresholds(state, numvals):
    """Inverse of split_sparse_thresholds: what pack.py writes next to the buffers."""
    out = dict(state)
    for name, n in numvals.items():
        out[PREFIX + name] = int(n)
    return out


def load_quantized(model, state, wbits, include_sparse=False, topX=0, skip=("lm_head",)):
    """Replace every nn.Linear of `model` (except `skip`) by a QuantLinearLUT sized from the checkpoint and load it.

    Mirrors llama.py:156-182: thresholds -> numvals, make_quant_lut, load_state_dict(strict=False).  A sparse checkpoint
    that lacks `full_rows` (the reference's own ones do) leaves them zero, exactly as there.  Returns the
    (missing_keys, unexpected_keys) of the load so callers can check what the checkpoint did not cover."""
    clean, numvals = split_sparse_thresholds(state)
    layers = find_linear_layers(model)
    for name in skip:
        layers.pop(name, None)
    if include_sparse:
        lacking = [n for n in layers if n not in numvals]
        if lacking:
            raise KeyError(f"include_sparse=True but the checkpoint has no sparse_threshold for: {lacking[:4]}{' ...' if len(lacking) > 4 else ''}")
    make_quant_lut(model, set(layers), wbits, include_sparse=include_sparse, numvals=numvals if include_sparse else None, topX=topX)
    return model.load_state_dict(clean, strict=False)


def quantized_state_dict(model):
    """state_dict() of a model holding QuantLinearLUT modules + the `sparse_threshold.*` entries pack.py:173-178 adds."""
    numvals = {name: int(m.vals.numel()) for name, m in model.named_modules()
               if isinstance(m, QuantLinearLUT) and m.include_sparse and hasattr(m, "vals")}
    return merge_sparse_thresholds(model.state_dict(), numvals)


# ---- files ----------------------------------------------------------------------------------------------------------------
CONFIG_NAME = "quant_config.json"


def write_quant_config(directory, wbits, **extra):
    """The sidecar pack.py:184-190 writes next to the checkpoint: {"wbits": N} (extra keys are allowed, the reference ignores them)."""
    data = {"wbits": int(wbits)}
    data.update(extra)
    path = os.path.join(directory or ".", CONFIG_NAME)
    with open(path, "w") as f:
        json.dump(data, f, indent=4)
    return path


def read_quant_config(path_or_dir):
    """-> dict of the sidecar that sits next to a checkpoint file (or in a directory); {} if there is none."""
    d = path_or_dir if os.path.isdir(path_or_dir) else os.path.dirname(path_or_dir)
    path = os.path.join(d or ".", CONFIG_NAME)
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        return json.load(f)


def save_checkpoint(model, path, wbits, write_config=True):
    """Write the model's packed checkpoint.  `*.safetensors`: tensors by safetensors, thresholds + wbits in the metadata;
    anything else: exactly what quantization/pack.py writes (torch.save of the state dict + sparse_threshold.* ints)."""
    state = quantized_state_dict(model)
    if path.endswith(".safetensors"):
        from safetensors.torch import save_file
        clean, numvals = split_sparse_thresholds(state)
        meta = {"format": "pt", "wbits": str(int(wbits)), "sparse_threshold": json.dumps(numvals)}
        save_file({k: v.detach().cpu().contiguous().clone() for k, v in clean.items()}, path, metadata=meta)  # clone: stacked siblings share storage
    else:
        torch.save({k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in state.items()}, path)
    if write_config:
        write_quant_config(os.path.dirname(path), wbits)
    return path


def load_checkpoint(model, path, wbits=None, include_sparse=None, topX=0, skip=("lm_head",)):
    """Read a packed checkpoint file (reference pickle or .safetensors) into `model` (see load_quantized).  `wbits` defaults to the
    sidecar's / metadata's value, `include_sparse` to "the file has sparse thresholds".  Returns load_quantized's result."""
    cfg = read_quant_config(path)
    if path.endswith(".safetensors"):
        from safetensors import safe_open
        state, meta = {}, {}
        with safe_open(path, framework="pt", device="cpu") as f:
            meta = f.metadata() or {}
            for k in f.keys():
                state[k] = f.get_tensor(k)
        numvals = json.loads(meta.get("sparse_threshold", "{}"))
        state = merge_sparse_thresholds(state, numvals)
        if "wbits" in meta:
            cfg.setdefault("wbits", int(meta["wbits"]))
    else:
        state = torch.load(path, map_location="cpu", weights_only=False)
    if wbits is None:
        if "wbits" not in cfg:
            raise ValueError(f"wbits not given and no {CONFIG_NAME} / metadata next to {path}")
        wbits = int(cfg["wbits"])
    if include_sparse is None:
        include_sparse = any(k.startswith(PREFIX) for k in state)
    return load_quantized(model, state, wbits, include_sparse=include_sparse, topX=topX, skip=skip)
