#!/usr/bin/env python3
"""Generate tests/golden/refkernel_*.npz on a B200: outputs of the REFERENCE's own CUDA kernels
(oracle/_ref/quant_cuda_ref*.so, built by oracle/build_ref.py from /root/reference with the 8-site dtype patch).

    gpurun -- 'python tests/golden/make_golden_gpu.py'      # writes gpurun_out/golden/*.npz; copy them to tests/golden/

Each file holds the inputs (reference buffer format), the initial `mul`, and `mul_out` after calling the
reference symbol that squeezellm/quant.py would pick.  They pin the CPU oracle's arithmetic
(tests/test_oracle_golden.py::test_oracle_matches_reference_kernel_outputs).  Shapes respect the reference's
unchecked preconditions (in % 128 == 0, out % 128 == 0).  Kept small so the fixtures stay a few hundred KB.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import build_ref  # noqa: E402
import oracle as orc  # noqa: E402

CASES = [  # name, bits, K, N, sparsity, topX, nonzero_full_rows, batch
    ("w4_dense", 4, 256, 128, 0.0, 0, False, 1),
    ("w3_dense", 3, 256, 128, 0.0, 0, False, 1),
    ("w4_spmv", 4, 256, 256, 0.01, 0, False, 1),
    ("w3_spmv", 3, 384, 128, 0.01, 0, False, 1),
    ("w4_hybrid", 4, 256, 128, 0.01, 10, True, 1),
    ("w3_hybrid", 3, 256, 128, 0.01, 10, True, 1),
    ("w4_dense_batched", 4, 128, 128, 0.0, 0, False, 3),
    ("w3_hybrid_batched", 3, 128, 128, 0.02, 4, True, 3),
]


def main():
    ref = build_ref.load()
    out = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    for name, bits, K, N, sp, topx, nz, batch in CASES:
        L = orc.make_layer(bits, K, N, sparsity=sp, topX=topx, seed=100 + CASES.index((name, bits, K, N, sp, topx, nz, batch)), nonzero_full_rows=nz)
        x = orc.make_vec(K, batch=batch, seed=17)
        init = (np.random.default_rng(5).standard_normal((batch, N)) * 0.5).astype(np.float32)
        T = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in L.items()}
        xv = torch.from_numpy(x).cuda() if batch > 1 else torch.from_numpy(x).cuda().reshape(-1)
        mul = torch.from_numpy(init.copy()).cuda() if batch > 1 else torch.from_numpy(init.copy()).cuda().reshape(-1)
        sfx = "_batched" if batch > 1 else ""
        if L["rows"] is not None and L["full_rows"] is not None:
            getattr(ref, f"vecquant{bits}matmul_spmv_hybrid_nuq_perchannel{sfx}")(
                T["rows"], T["cols"], T["vals"], xv, T["full_rows"], T["full_row_indices"], mul, N, T["qweight"], T["lookup_table"])
        elif L["rows"] is not None:
            getattr(ref, f"vecquant{bits}matmul_spmv_nuq_perchannel{sfx}")(T["rows"], T["cols"], T["vals"], xv, mul, N, T["qweight"], T["lookup_table"])
        else:
            getattr(ref, f"vecquant{bits}matmul_nuq_perchannel{sfx}")(xv, T["qweight"], mul, T["lookup_table"])
        torch.cuda.synchronize()
        d = dict(bits=bits, K=K, N=N, vec=x, mul_init=init, mul_out=mul.cpu().numpy().reshape(batch, N))
        for k in ("qweight", "lookup_table", "rows", "cols", "vals", "full_rows", "full_row_indices"):
            if L[k] is not None:
                d[k] = L[k]
        np.savez_compressed(os.path.join(out, f"refkernel_{name}.npz"), **d)
        e = np.abs(d["mul_out"] - orc.forward_f64(L, x, mul_init=init)).max()
        print(f"{name}: max |ref_kernel - oracle_f64| = {e:.3e}")


if __name__ == "__main__":
    main()
