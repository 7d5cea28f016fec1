"""Developer probe (GPU box): fused chain kernel (conv_chain.cuh) against the per-layer tcgen05 path and the oracle.

    python tools/gpu_chain_check.py [--archs rfdn imdn rlfn] [--time 1]
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
SHAPES = [(1, 16, 40), (1, 64, 64), (1, 40, 150), (3, 33, 47), (1, 130, 260), (1, 256, 256), (5, 256, 256)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--archs", nargs="+", default=["rfdn", "imdn", "rlfn"])
    ap.add_argument("--time", type=int, default=0)
    ap.add_argument("--oracle", type=int, default=1)
    ap.add_argument("--store_all", type=int, default=0)
    ap.add_argument("--mask", type=int, default=-1)
    ap.add_argument("--timeline", type=int, default=0)
    ap.add_argument("--dbg", type=int, default=0)
    ap.add_argument("--force", type=int, default=0, help="1: chain_enable = 2 (fused kernel also when a cluster gets several bands)")
    ap.add_argument("--shapes", type=int, nargs="*", default=None)
    a = ap.parse_args()
    import torch
    from ntire2022_esr_b200 import Engine

    worst = 0.0
    for arch in a.archs:
        mid = IDS[arch]
        w = O.load_weights(os.path.join(ROOT, "tests", "golden", "weights", O.MODELS[mid]["weights"] + ".npz"))
        dr = O.MODELS[mid]["data_range"]
        eng = Engine(arch, 0)
        eng.load_state_dict(w)
        for si, (b, h, wd) in enumerate(SHAPES):
            if a.shapes is not None and si not in a.shapes:
                continue
            rng = np.random.default_rng(si)
            x = (rng.random((b, 3, h, wd), dtype=np.float32) * dr).astype(np.float16)
            xt = torch.from_numpy(x).cuda()
            ys = {}
            for chain in (0, 1):
                eng.set_option("chain_enable", (2 if a.force else 1) if chain else 0)
                eng.set_option("chain_store_all", a.store_all)
                eng.set_option("chain_mask", a.mask)
                for graph in ((0, 1) if chain else (0,)):
                    eng.set_option("use_graph", graph)
                    yt = eng.forward(xt)
                    yt2 = eng.forward(xt)      # second call: flags must have been re-armed
                    torch.cuda.synchronize()
                    ys[(chain, graph)] = yt.float().cpu().numpy()
                    if not torch.equal(yt, yt2):
                        dd = (yt.float() - yt2.float()).abs().amax(dim=(0, 1))
                        rows = torch.nonzero(dd.amax(dim=1) > 0).flatten().cpu().numpy() // 4
                        cols = torch.nonzero(dd.amax(dim=0) > 0).flatten().cpu().numpy() // 4
                        print(f"  NOT REPEATABLE chain={chain} graph={graph} {b}x{h}x{wd}: max {float(dd.max()):.3e}; LR rows {np.unique(rows)[:24]} cols {np.unique(cols)[:24]}", flush=True)
            y0, y1, y1g = ys[(0, 0)], ys[(1, 0)], ys[(1, 1)]
            d = float(np.abs(y1 - y0).max() / dr)
            dg = float(np.abs(y1g - y1).max() / dr)
            nbad = int((np.abs(y1 - y0) / dr > 2e-3).sum())
            msg = f"CHAIN {arch} {b}x{h}x{wd}: max|chain - layered|/range={d:.3e} (>2e-3: {nbad})  graph vs direct={dg:.1e} finite={np.isfinite(y1).all()}"
            if a.oracle and b * h * wd <= 130 * 260:
                ref = O.forward(arch, w, x.astype(np.float32), dtype=np.float32)
                ps = [10 * np.log10(dr * dr / max(np.mean((y.astype(np.float64) - ref) ** 2), 1e-30)) for y in (y0, y1)]
                msg += f"  psnr vs oracle: layered {ps[0]:.2f} chain {ps[1]:.2f} dB"
            print(msg, flush=True)
            if d > 2e-3:
                dd = np.abs(y1 - y0).max(axis=(0, 1)) / dr
                rows = np.unique(np.nonzero(dd.max(axis=1) > 2e-3)[0] // 4)
                cols = np.unique(np.nonzero(dd.max(axis=0) > 2e-3)[0] // 4)
                print(f"  bad LR rows ({len(rows)}): {rows[:40]}\n  bad LR cols ({len(cols)}): {cols[:40]}", flush=True)
            worst = max(worst, d)
            if a.time and (b, h, wd) in ((1, 256, 256), (5, 256, 256)):
                for chain in (0, 1):
                    eng.set_option("chain_enable", (2 if a.force else 1) if chain else 0)
                    eng.set_option("use_graph", 1)
                    yt = eng.forward(xt)
                    for _ in range(20):
                        eng.forward(xt, out=yt)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(200):
                        eng.forward(xt, out=yt)
                    e1.record()
                    torch.cuda.synchronize()
                    print(f"TIME {arch} chain={chain} {b}x{h}x{wd}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us/forward", flush=True)
                    if a.time > 1:
                        for name, fl, ms in eng.profile_launches(xt, yt, reps=10):
                            print(f"PROF {name[:60]:60s} {ms * 1e3:8.2f} us  {fl / 1e9:7.3f} GF  {fl / (ms * 1e-3) / 1e12 if ms > 0 else 0:7.1f} TF/s", flush=True)
        if a.timeline:
            x = (np.random.default_rng(0).random((1, 3, 256, 256), dtype=np.float32) * dr).astype(np.float16)
            xt = torch.from_numpy(x).cuda()
            eng.set_option("chain_enable", 1)
            eng.set_option("use_graph", 0)
            eng.set_option("tc_timeline", 1)
            eng.set_option("tc_dbg_flags", a.dbg)
            for _ in range(3):
                eng.forward(xt)
            torch.cuda.synchronize()
            tl = eng.debug_timeline(240).reshape(240 * 128)
            for ci in range(a.timeline):
                bt = tl[(100 + 5 * ci) * 128:(100 + 5 * ci) * 128 + 148 * 4].reshape(148, 4)
                bt = bt[bt[:, 0] > 0]
                if len(bt):
                    z = bt[:, 0].min()
                    print(f"TLB chain {ci}: {len(bt)} blocks; entry {int(bt[:, 0].min() - z)}..{int(bt[:, 0].max() - z)} ns, set-up done "
                          f"{int(bt[:, 1].min() - z)}..{int(bt[:, 1].max() - z)}, roles done {int(bt[:, 2].min() - z)}..{int(bt[:, 2].max() - z)} "
                          f"(median {int(np.median(bt[:, 2]) - z)}), exit {int(bt[:, 3].min() - z)}..{int(bt[:, 3].max() - z)}")
                    for bi in (0, 1, 2, 3, 60, 61, 128, 129):
                        if bi < len(bt):
                            print(f"   block {bi}: entry {int(bt[bi, 0] - z)} setup {int(bt[bi, 1] - z)} roles {int(bt[bi, 2] - z)} exit {int(bt[bi, 3] - z)}")
                    order = np.argsort(bt[:, 2])
                    print("   roles-done time by block (sorted, every 16th):", [(int(o), int(bt[o, 2] - z)) for o in order[::16]])
                t = tl[(224 + 2 * ci) * 128:(226 + 2 * ci) * 128]
                t0 = t[0]
                print(f"TLC chain {ci}: kernel start 0 end {int(t[1] - t0)} cycles")
                for g in range(8):
                    mm = t[64 + g * 6: 64 + g * 6 + 6]
                    ep = t[128 + g * 4: 128 + g * 4 + 4]
                    if mm.max() == 0:
                        continue
                    print(f"   layer {g}: mma step commit (i=5..0): " + " ".join(f"{int(v - t0):7d}" for v in mm[::-1]) +
                          "   epilogue done (j=3..0): " + " ".join(f"{int(v - t0):7d}" for v in ep[::-1]), flush=True)
                for j in (3, 2, 1, 0):
                    e5 = t[8 + j * 8: 8 + j * 8 + 5]
                    if e5.max() > 0:
                        print(f"      pw epi row {j}: begin {int(e5[0] - t0)} +sd wait {int(e5[1] - e5[0])} +bar waits {int(e5[2] - e5[1])} +units {int(e5[3] - e5[2])} +fence/arrive {int(e5[4] - e5[3])}")
                w4 = t[192 + 48: 192 + 60]
                if w4.max() > 0:
                    print("      layer 2 mid-step wait for the next step (i=5..0) begin/+blocked: " + "  ".join(f"{int(w4[2 * i] - t0)}/+{int(w4[2 * i + 1] - w4[2 * i])}" for i in range(5, -1, -1) if w4[2 * i] > 0))
                for g in range(4):
                    w = t[192 + g * 12: 192 + g * 12 + 12]
                    if w.max() == 0:
                        continue
                    print(f"      issuer layer {g} (i=5..0) waits done / mma issued: " +
                          "  ".join(f"{int(w[2 * i] - t0)}/{int(w[2 * i + 1] - t0)}" for i in (5, 4, 3, 2, 1, 0)), flush=True)
    print(f"WORST {worst:.3e}")


if __name__ == "__main__":
    main()
