"""The sequence kernel across TWO real GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_zz_sequence_2gpu.py -m gpu`; skipped with
fewer than two GPUs).  Every rank runs the same chain over its own column shards (sharding.shard_state); owners store their slice of
each result into every rank's arena over NVLink (sharding.PeerArena) and the next matvec's input poll is the exchange.  Every rank must
end up with the full vectors of the fp64 oracle chain of the UNSHARDED layers (fp16 hand-overs), also when one rank starts late."""
import os
import sys
import time

import numpy as np
import pytest
import torch

from util import ROOT, REL_TOL, orc

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, results):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    from squeezellm_b200.quant import QuantLinearLUT
    from squeezellm_b200.runtime import DecodeSequence
    from squeezellm_b200.sharding import PeerArena, shard_bounds, shard_state
    try:
        worst = 0.0
        for bits, dims, mode in ((4, [(4096, 4096, 0), (2048, 11008, 1024), (11008, 4096, 0), (4096, 8192, 0)], "exact"),
                                 (3, [(512, 1024, 0), (512, 768, 256), (768, 512, 0)], "fp16")):
            Ls = []
            for i, (K, N, _) in enumerate(dims):
                L = orc.make_layer(bits, K, N, sparsity=0.0045 if K > 1000 else 0.02, topX=6, seed=300 + 7 * i + bits, nonzero_full_rows=True)
                L["lookup_table"] = (L["lookup_table"] * (50.0 / np.sqrt(K))).astype(np.float32)
                Ls.append(L)
            peer = PeerArena(rank, world, dev)
            seq = DecodeSequence(dims[0][0], dev, lut_mode=mode, peer=peer)
            prev, vecs, keep = seq.input, [], []
            for L, (K, N, off) in zip(Ls, dims):
                c0, c1 = shard_bounds(N, world)[rank]
                st = shard_state({k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else None) for k, v in L.items()
                                  if k in ("qweight", "lookup_table", "rows", "cols", "vals", "full_rows", "full_row_indices")}, c0, c1)
                m = QuantLinearLUT(bits, K, c1 - c0, False, include_sparse=True, numvals=int(st["vals"].numel()), topX=6)
                m.load_state_dict(st, strict=False)
                m = m.to(dev)
                keep.append(m)
                prev = seq.matvec(m, prev[off:off + K])
                vecs.append(prev)
            seq.compile(outputs=vecs)
            for rep in range(3):
                x = orc.make_vec(dims[0][0], seed=10 + rep).reshape(-1).astype(np.float16)
                seq.x.copy_(torch.from_numpy(x).to(dev))
                if rep == 1 and rank == 1:
                    torch.cuda.synchronize()
                    time.sleep(0.3)      # a late rank: the others wait inside the kernel (bounded), nothing is lost
                outs = [o.clone() for o in seq.replay()]
                torch.cuda.synchronize()
                assert not seq.error()
                want, cur = [], None
                for L, (K, N, off) in zip(Ls, dims):
                    xin = x.astype(np.float64) if cur is None else cur[off:off + K]
                    LL = dict(L)
                    if mode == "fp16":
                        LL["lookup_table"] = L["lookup_table"].astype(np.float16).astype(np.float32)
                    cur = orc.forward_f64(LL, xin.astype(np.float32).reshape(1, K)).reshape(N).astype(np.float16).astype(np.float64)
                    want.append(cur)
                for o, w in zip(outs, want):
                    got = o.float().cpu().numpy().astype(np.float64)
                    assert got.shape == w.shape
                    worst = max(worst, float(np.abs(got - w).max() / np.abs(w).max()))
            dist.barrier()
        results[rank] = worst
    finally:
        torch.cuda.synchronize()
        os._exit(0 if results.get(rank) is not None else 1)


def test_sequence_two_gpus_matches_oracle_chain():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, results)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert not p.is_alive(), "worker hung"
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert len(results) == 2 and max(results.values()) < 2 * REL_TOL, dict(results)
