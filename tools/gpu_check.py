"""Developer probe (GPU box): one arch / dtype / engine-option combination vs the numpy oracle.

    python tools/gpu_check.py rfdn f16 --tc 1 --shift 1 --size 64 64 --batch 1
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
    ap.add_argument("arch")
    ap.add_argument("dtype", choices=["f32", "f16"])
    ap.add_argument("--tc", type=int, default=1)
    ap.add_argument("--shift", type=int, default=0)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--size", type=int, nargs=2, default=[64, 64])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--host", type=int, default=0)
    ap.add_argument("--time", type=int, default=0)
    ap.add_argument("--profile", type=int, default=0)
    ap.add_argument("--timeline", type=int, default=0)
    ap.add_argument("--dbg", type=int, default=0)
    ap.add_argument("--slots", type=int, default=4)
    ap.add_argument("--nocheck", type=int, default=0)
    ap.add_argument("--pdl", type=int, default=0)
    a = ap.parse_args()
    import torch
    from ntire2022_esr_b200 import Engine

    mid = IDS[a.arch]
    w = O.load_weights(os.path.join(ROOT, "tests", "golden", "weights", O.MODELS[mid]["weights"] + ".npz"))
    dr = O.MODELS[mid]["data_range"]
    rng = np.random.default_rng(1)
    h, wd = a.size
    x = (rng.random((a.batch, 3, h, wd), dtype=np.float32) * dr).astype(np.float16 if a.dtype == "f16" else np.float32)
    eng = Engine(a.arch, 0)
    eng.set_option("tc_enable", a.tc)
    eng.set_option("tc_shift_mode", a.shift)
    eng.set_option("use_graph", a.graph)
    eng.set_option("tc_rows_per_item", a.rows)
    eng.set_option("tc_timeline", a.timeline)
    eng.set_option("tc_dbg_flags", a.dbg)
    eng.set_option("tc_acc_slots", a.slots)
    eng.set_option("use_pdl", a.pdl)
    eng.load_state_dict(w)
    if a.host:
        y = eng.forward_host(x)
    else:
        xt = torch.from_numpy(x).cuda()
        yt = eng.forward(xt)
        torch.cuda.synchronize()
        y = yt.cpu().numpy()
    if a.nocheck:
        print(f"NOCHECK {a.arch} {a.dtype} {a.batch}x{h}x{wd} finite={np.isfinite(y).all()} mean={float(np.mean(y.astype(np.float64))):.4f}", flush=True)
        if not a.host:
            timing(a, eng, xt, yt)
            extras(a, eng, xt, yt)
        return
    t0 = time.time()
    ref = O.forward(a.arch, w, x.astype(np.float32), dtype=np.float64 if a.dtype == "f32" else np.float32)
    t1 = time.time()
    err = np.abs(y.astype(np.float64) - ref).max() / dr
    mse = np.mean((y.astype(np.float64) - ref) ** 2) / dr ** 2
    psnr = 10 * np.log10(1.0 / mse) if mse > 0 else float("inf")
    print(f"CHECK {a.arch} {a.dtype} tc={a.tc} shift={a.shift} graph={a.graph} {a.batch}x{h}x{wd} host={a.host}: "
          f"max|err|/range={err:.3e} psnr={psnr:.2f} dB finite={np.isfinite(y).all()} (oracle {t1 - t0:.1f}s)", flush=True)
    if not a.host:
        timing(a, eng, xt, yt)
        extras(a, eng, xt, yt)


def timing(a, eng, xt, yt):
    import torch
    h, wd = a.size
    if a.time:
        for _ in range(5):
            eng.forward(xt, out=yt)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(a.time):
            eng.forward(xt, out=yt)
        ev1.record()
        torch.cuda.synchronize()
        print(f"TIME {a.arch} {a.dtype} tc={a.tc} graph={a.graph} {a.batch}x{h}x{wd}: {ev0.elapsed_time(ev1) / a.time * 1e3:.1f} us/forward",
              flush=True)


def extras(a, eng, xt, yt):
    import torch
    if a.profile:
        for _ in range(200):
            eng.forward(xt, out=yt)
        torch.cuda.synchronize()
        for name, fl, ms in eng.profile_launches(xt, yt, reps=a.profile):
            print(f"PROF {name:45s} {ms * 1e3:8.2f} us  {fl / 1e9:7.3f} GF  {fl / (ms * 1e-3) / 1e12 if ms > 0 else 0:7.1f} TF/s", flush=True)
    if a.timeline:
        names = [n for n in eng.launch_names(a.batch, a.size[0], a.size[1], 1) if n.startswith("conv_tc")]
        eng.set_option("use_graph", 0)
        for _ in range(3):
            eng.forward(xt, out=yt)
        tl = eng.debug_timeline(len(names))
        for i, n in enumerate(names[: a.timeline]):
            t = tl[i]
            t0 = t[0, 0]
            fmt = lambda v: " ".join(f"{int(x - t0):6d}" if x else "     -" for x in v)
            print(f"TL {n}")
            print("   cta: start,setup,prod_done,w_ready,store_done,end:", fmt(t[0, :6]))
            print("   tma strips issued:", fmt(t[1, :12]))
            print("   epi (ld done, math+stage done) per tile:", fmt(t[1, 16:28]))
            print("   mma (begin,commit) per tile:", fmt(t[2, :10]))
            print("   epi (wait,got,stored) per tile:", fmt(t[3, :15]), flush=True)


if __name__ == "__main__":
    main()
