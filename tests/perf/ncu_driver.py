#!/usr/bin/env python3
"""Tiny driver for ncu captures: a few GEMV launches of one shape, rotating weight copies (cold HBM reads).
    python tests/perf/ncu_driver.py BITS K N SPARSITY TOPX MODE(acc|fused) LAUNCHES"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from squeezellm_b200.quant import quant_cuda as qc  # noqa: E402


def main():
    bits, K, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    sp, topx, mode, n = float(sys.argv[4]), int(sys.argv[5]), sys.argv[6], int(sys.argv[7])
    g = torch.Generator(device="cuda").manual_seed(1)
    copies = []
    nnz = int(round(sp * K * N))
    for _ in range(min(n, 8)):
        L = dict(q=torch.randint(-2**31, 2**31 - 1, (K // 32 * bits, N), dtype=torch.int64, device="cuda", generator=g).to(torch.int32),
                 lut=torch.sort(torch.randn((N, 2**bits), device="cuda", generator=g) * 0.02, dim=1).values.contiguous())
        if nnz:
            counts = torch.bincount(torch.randint(0, N, (nnz,), device="cuda", generator=g), minlength=N)
            rows = torch.zeros(N + 1, dtype=torch.int32, device="cuda")
            rows[1:] = torch.cumsum(counts, 0).to(torch.int32)
            L.update(rows=rows, cols=torch.randint(0, K, (nnz,), device="cuda", generator=g).to(torch.int32),
                     vals=torch.randn(nnz, device="cuda", generator=g) * 0.1)
        if topx:
            L.update(fr=torch.zeros((K, topx), device="cuda"), fri=torch.zeros(topx, dtype=torch.int32, device="cuda"))
        copies.append(L)
    x32 = torch.randn(K, device="cuda").half().float()
    x16 = x32.half()
    y = torch.zeros(N, device="cuda")
    torch.cuda.synchronize()
    for i in range(n):
        L = copies[i % len(copies)]
        if mode == "fused":
            qc.lutgemv_fused(x16, L["q"], L["lut"], bits, None, L.get("rows"), L.get("cols"), L.get("vals"), L.get("fr"), L.get("fri"))
        elif "rows" in L and "fr" in L:
            getattr(qc, f"vecquant{bits}matmul_spmv_hybrid_nuq_perchannel")(L["rows"], L["cols"], L["vals"], x32, L["fr"], L["fri"], y, N, L["q"], L["lut"])
        elif "rows" in L:
            getattr(qc, f"vecquant{bits}matmul_spmv_nuq_perchannel")(L["rows"], L["cols"], L["vals"], x32, y, N, L["q"], L["lut"])
        else:
            getattr(qc, f"vecquant{bits}matmul_nuq_perchannel")(x32, L["q"], y, L["lut"])
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
