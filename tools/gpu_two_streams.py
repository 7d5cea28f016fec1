"""Developer probe (GPU box): throughput of independent batch-1 forwards issued on 1, 2 or 3 streams (one engine
handle and workspace per stream) - how much of the per-layer latency (launch gaps, epilogue tails, the small ESA
kernels) can be hidden behind another request's kernels."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import esr_oracle as O  # noqa: E402
from ntire2022_esr_b200 import Engine  # noqa: E402

w = O.load_weights(os.path.join(ROOT, "tests", "golden", "weights", "rfdn_baseline.npz"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for ns in (1, 2, 3, 4):
    engs = [Engine("rfdn", 0).load_state_dict(w) for _ in range(ns)]
    streams = [torch.cuda.Stream() for _ in range(ns)]
    xs = [(torch.rand(B, 3, 256, 256) * 255).half().cuda() for _ in range(ns)]
    ys = [torch.empty(B, 3, 1024, 1024, dtype=torch.float16, device="cuda") for _ in range(ns)]
    def run(n):
        for i in range(n):
            k = i % ns
            with torch.cuda.stream(streams[k]):
                engs[k].forward(xs[k], out=ys[k])
    run(40)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 600
    run(n)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"streams={ns} batch={B}: {n * B / dt:.0f} img/s  ({dt / n * 1e6:.1f} us per forward issued)", flush=True)
