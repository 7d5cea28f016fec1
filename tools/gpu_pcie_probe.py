"""Developer probe (GPU box, N ranks under torchrun): pinned host <-> device copy bandwidth per rank, one rank at a time
and all ranks at once, with and without binding the rank to its GPU's NUMA node.  Names the limiter of the end-to-end
numbers at N = 8 (every rank moves ~20 GB/s of results to the host at batch 1).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/gpu_pcie_probe.py
"""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import pin_to_gpu_numa_node  # noqa: E402


def bw(dev, h, d, d2h, reps=20):
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        for _ in range(2):
            (h.copy_(d, non_blocking=True) if d2h else d.copy_(h, non_blocking=True))
        s.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            (h.copy_(d, non_blocking=True) if d2h else d.copy_(h, non_blocking=True))
        s.synchronize()
    return reps * h.numel() / (time.perf_counter() - t0) / 1e9


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {}
    for pinned_node in (False, True):
        aff = pin_to_gpu_numa_node(local) if pinned_node else None
        h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
        res = {"affinity": aff}
        for d2h in (True, False):
            key = "d2h" if d2h else "h2d"
            alone = 0.0
            for r in range(world):          # one rank at a time
                if world > 1:
                    dist.barrier()
                if r == rank:
                    alone = bw(dev, h, d, d2h)
            if world > 1:
                dist.barrier()
            together = bw(dev, h, d, d2h, reps=40)
            res[key] = {"alone_GBs": round(alone, 1), "all_ranks_GBs": round(together, 1)}
        out["numa_bound" if pinned_node else "unbound"] = res
        del h, d
    t = torch.tensor([out[k][c][m] for k in ("unbound", "numa_bound") for c in ("d2h", "h2d") for m in ("alone_GBs", "all_ranks_GBs")],
                     dtype=torch.float64, device=dev)
    allt = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allt, t)
    else:
        allt = [t]
    if rank == 0:
        names = [f"{k}.{c}.{m}" for k in ("unbound", "numa_bound") for c in ("d2h", "h2d") for m in ("alone", "all")]
        table = {n: [round(float(a[i]), 1) for a in allt] for i, n in enumerate(names)}
        print(json.dumps({"world": world, "per_rank_GBs": table, "sum_all_GBs": {n: round(sum(v), 1) for n, v in table.items() if n.endswith("all")},
                          "affinity_rank0": out["numa_bound"]["affinity"]}))
        try:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout)
        except Exception as e:
            print("topo:", e)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
