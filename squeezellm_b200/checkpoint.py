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
Nothing here touches the GPU.
"""
import torch.nn as nn

from .quant import QuantLinearLUT, make_quant_lut

__all__ = ["find_linear_layers", "split_sparse_thresholds", "merge_sparse_thresholds", "load_quantized", "quantized_state_dict"]

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
