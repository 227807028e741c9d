"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tests/perf/sanitize_small.py
Shapes are tiny so that the instrumented run finishes in seconds; every flush mode, both bit widths, ragged strips,
CSR + dense rows, a misaligned CSR start (16-byte staging with clipped tails) and sibling stacking are covered."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import orc, rel_err, to_torch  # noqa: E402
from squeezellm_b200 import quant as Q  # noqa: E402

qc = Q.quant_cuda
worst = 0.0
for det in (False, True):
    qc.set_deterministic(det)
    for bits, K, N, sp, topx in [(4, 256, 192, 0.02, 3), (3, 512, 200, 0.03, 5), (4, 1024, 1028, 0.0, 0), (3, 256, 64, 0.05, 33), (4, 2048, 640, 0.01, 10)]:
        L = orc.make_layer(bits, K, N, sparsity=sp, topX=topx, seed=K + N, nonzero_full_rows=True, bias=True)
        T = to_torch(L)
        x = orc.make_vec(K, seed=1)
        for dtype in (torch.float16, torch.float32):
            xt = torch.from_numpy(x).cuda().reshape(-1).to(dtype)
            y = qc.lutgemv_fused(xt, T["qweight"], T["lookup_table"], bits, T["bias"], T.get("rows"), T.get("cols"), T.get("vals"),
                                 T.get("full_rows"), T.get("full_row_indices"))
            torch.cuda.synchronize()
            want = orc.forward_f64(L, x, mul_init=L["bias"][None, :])
            e = rel_err(y.float().cpu().numpy(), want)
            worst = max(worst, e)
            assert e < 1e-3, (bits, K, N, det, dtype, e)
        # accumulate path through one of the 12 symbols
        mul = torch.zeros(N, device="cuda")
        xv = torch.from_numpy(x).cuda().reshape(-1).float()
        if sp and topx:
            getattr(qc, f"vecquant{bits}matmul_spmv_hybrid_nuq_perchannel")(T["rows"], T["cols"], T["vals"], xv, T["full_rows"], T["full_row_indices"], mul, N, T["qweight"], T["lookup_table"])
        else:
            getattr(qc, f"vecquant{bits}matmul_nuq_perchannel")(xv, T["qweight"], mul, T["lookup_table"])
        torch.cuda.synchronize()
        e = rel_err(mul.cpu().numpy(), orc.forward_f64(L, x))
        worst = max(worst, e)
        assert e < 1e-4, (bits, K, N, e)
qc.set_deterministic(False)
print("sanitize_small: all cases passed, worst rel err %.2e" % worst)
