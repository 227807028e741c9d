#!/usr/bin/env python3
"""Generate tests/golden/pack2_*.npz by running the REFERENCE's own packer on CPU.

Run in the build container (needs /root/reference):   python tests/golden/make_golden_pack2.py

The reference module squeezellm/quant.py imports the CUDA extension `quant_cuda` at import time
(quant.py:5).  `pack2` (quant.py:97-208) never calls it, so an empty stub module is registered under
that name; nothing else of the reference is altered.  For each case we feed pack2 random per-channel
(centroids, indices) and (for sparse cases) an outlier matrix, and store its inputs and the buffers it
produced: qweight, lookup_table, rows, cols, vals.  These fixtures pin the packed-int32 layout (4-bit
and the 3-bit 32-in-3-words layout), the fp32 LUT layout and the "outlier minus zero-centroid" CSR
convention (quant.py:117-131) for the oracle and for our own packer.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SQLLM_REFERENCE_DIR", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference_quant():
    sys.modules.setdefault("quant_cuda", types.ModuleType("quant_cuda"))  # stub: pack2 never calls it
    sys.path.insert(0, REF)
    from squeezellm.quant import QuantLinearLUT  # noqa: E402
    return QuantLinearLUT


def one_case(QuantLinearLUT, bits, K, N, sparse_frac, seed):
    rng = np.random.default_rng(seed)
    centroids = np.sort(rng.standard_normal((N, 2**bits)).astype(np.float32) * 0.02, axis=1)
    indices = rng.integers(0, 2**bits, size=(N, K)).astype(np.int64)
    lut = [[(centroids[c], indices[c])] for c in range(N)]
    include_sparse = sparse_frac > 0
    outl_dense = np.zeros((N, K), dtype=np.float32)
    outliers = None
    if include_sparse:
        mask = rng.random((N, K)) < sparse_frac
        outl_dense[mask] = (rng.standard_normal(int(mask.sum())) * 0.3).astype(np.float32)
        outliers = torch.from_numpy(outl_dense.copy()).to_sparse()
    lin = torch.nn.Linear(K, N, bias=False)
    q = QuantLinearLUT(bits, K, N, False, include_sparse=include_sparse, numvals=0, topX=0)
    q.pack2(lin, (lut, outliers), include_sparse)
    out = dict(bits=bits, K=K, N=N, centroids=centroids, indices=indices.astype(np.uint8),
               outliers_dense=outl_dense, qweight=q.qweight.numpy(), lookup_table=q.lookup_table.numpy())
    if include_sparse:
        out.update(rows=q.rows.numpy(), cols=q.cols.numpy(), vals=q.vals.numpy())
    return out


def main():
    Q = load_reference_quant()
    cases = [  # (name, bits, K, N, sparse_frac, seed)
        ("w4_dense", 4, 128, 128, 0.0, 1),
        ("w3_dense", 3, 128, 128, 0.0, 2),
        ("w4_sparse", 4, 256, 128, 0.01, 3),
        ("w3_sparse", 3, 256, 128, 0.01, 4),
        ("w3_tall", 3, 512, 128, 0.0, 5),
        ("w4_wide", 4, 128, 256, 0.005, 6),
    ]
    for name, bits, K, N, sf, seed in cases:
        d = one_case(Q, bits, K, N, sf, seed)
        path = os.path.join(OUT, f"pack2_{name}.npz")
        np.savez_compressed(path, **d)
        print(path, {k: getattr(v, "shape", v) for k, v in d.items()})


if __name__ == "__main__":
    main()
