"""Developer probe (GPU box): hammer the pipelined host entry points (4 requests in flight, each on its own stream and
workspace, so the kernels of different requests share the GPU) and compare every result with the first one.

    python tools/gpu_stress_e2e.py [--n 20000] [--arch rfdn] [--size 256 256] [--u8 1]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import esr_oracle as O  # noqa: E402

IDS = {"imdn": -1, "rfdn": 0, "rlfn": 4, "bsrn": 18, "fmen": 3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--arch", default="rfdn")
    ap.add_argument("--size", type=int, nargs=2, default=[256, 256])
    ap.add_argument("--u8", type=int, default=1)
    ap.add_argument("--opt", nargs="*", default=[], help="engine options key=value")
    a = ap.parse_args()
    import torch
    from ntire2022_esr_b200 import Engine, _cabi
    mid = IDS[a.arch]
    w = O.load_weights(os.path.join(ROOT, "tests", "golden", "weights", O.MODELS[mid]["weights"] + ".npz"))
    dr = O.MODELS[mid]["data_range"]
    eng = Engine(a.arch, 0)
    for kv in a.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    eng.load_state_dict(w)
    h, wd = a.size
    nbuf = 6
    g = torch.Generator().manual_seed(1)
    if a.u8:
        xs = [(torch.rand(1, h, wd, 3, generator=g) * 255).to(torch.uint8).pin_memory() for _ in range(nbuf)]
        ys = [torch.empty(1, 4 * h, 4 * wd, 3, dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
        submit = lambda x, y: eng.forward_host_u8_async_ptr(x.data_ptr(), y.data_ptr(), 1, h, wd, dr, _cabi.DTYPE_F16)
    else:
        xs = [(torch.rand(1, 3, h, wd, generator=g) * dr).half().pin_memory() for _ in range(nbuf)]
        ys = [torch.empty(1, 3, 4 * h, 4 * wd, dtype=torch.float16).pin_memory() for _ in range(nbuf)]
        submit = lambda x, y: eng.forward_host_async_ptr(x.data_ptr(), y.data_ptr(), 1, h, wd, _cabi.DTYPE_F16)
    for i in range(nbuf):
        submit(xs[i], ys[i])
    eng.host_wait(-1)
    want = [y.clone() for y in ys]
    tickets, bad, t0 = [], 0, time.time()
    for i in range(a.n):
        k = i % nbuf
        if i >= nbuf:
            eng.host_wait(tickets[i - nbuf])
            if i % 97 == 0 and not torch.equal(ys[k], want[k]):
                bad += 1
        try:
            tickets.append(submit(xs[k], ys[k]))
        except Exception as e:
            print(f"STRESS FAILED at request {i}: {e}", flush=True)
            raise
    eng.host_wait(-1)
    bad += sum(0 if torch.equal(ys[k], want[k]) else 1 for k in range(nbuf))
    print(f"STRESS {a.arch} u8={a.u8} {h}x{wd}: {a.n} requests in {time.time() - t0:.1f} s, mismatching results: {bad}", flush=True)


if __name__ == "__main__":
    main()
