"""Latency of a small all-reduce inside a CUDA graph on N GPUs: NCCL vs torch symmetric-memory one-shot.
   torchrun --nproc-per-node N tests/perf/allreduce_latency.py"""
import os
import sys
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm

def bench(fn, iters=200):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay(); torch.cuda.synchronize(); dist.barrier()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters

for n in (4096, 12288, 22016):
    x = torch.ones(n, device=dev, dtype=torch.float16)
    t_nccl = bench(lambda: dist.all_reduce(x))
    msg = f"n={n} fp16 nccl {t_nccl:.2f} us"
    try:
        buf = symm.empty(n, dtype=torch.float16, device=dev)
        hdl = symm.rendezvous(buf, dist.group.WORLD)
        buf.fill_(1.0)
        gname = dist.group.WORLD.group_name
        t_one = bench(lambda: torch.ops.symm_mem.one_shot_all_reduce(buf, "sum", gname))
        out = torch.ops.symm_mem.one_shot_all_reduce(buf, "sum", gname)
        ok = bool((out == world).all().item())
        msg += f" | symm one_shot {t_one:.2f} us correct={ok}"
        t_two = bench(lambda: torch.ops.symm_mem.two_shot_all_reduce_(buf, "sum", gname))
        msg += f" | two_shot_ {t_two:.2f} us"
    except Exception as e:  # noqa: BLE001
        msg += f" | symm failed: {type(e).__name__}: {str(e)[:200]}"
    if rank == 0:
        print(msg, flush=True)
sys.stdout.flush()
torch.cuda.synchronize()
os._exit(0)
