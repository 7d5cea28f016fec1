"""Measures the BASELINE.json configs on one GPU (per-GPU share of the 8-GPU configs) and the same graphs in
PyTorch eager on the same GPU (oracle/esr_oracle_torch.py = the ATen ops the reference's nn.Modules call:
cuDNN / cuBLAS), fp32 and fp16.  Prints one JSON line per measurement."""
import json, os, random, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import esr_oracle as O
from oracle import esr_oracle_torch as OT
from ntire2022_esr_b200 import build_model
from ntire2022_esr_b200.sharded import forward_bucketed

IDS = {"imdn": -1, "rfdn": 0, "rlfn": 4, "bsrn": 18}


def weights(arch):
    return O.load_weights(os.path.join(ROOT, "tests", "golden", "weights", O.MODELS[IDS[arch]]["weights"] + ".npz"))


def timeit(fn, min_s=0.6, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    while True:
        fn(); n += 1
        if n % 5 == 0 and time.perf_counter() - t0 > min_s:
            break
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def div2k_shapes(n=32):
    rnd = random.Random(0)
    pool = [339] * 20 + [384] * 4 + [342] * 3 + [288] * 2 + [324, 408, 510]
    out = []
    for _ in range(n):
        s = rnd.choice(pool)
        hw = (s, 510)
        if rnd.random() < 0.25:
            hw = (510, s)
        out.append(hw)
    return out


def main():
    dev = torch.device("cuda:0")
    res = []
    for arch, B, (h, w), tag in [("rfdn", 1, (256, 256), "config2 RFDN b1 256x256"),
                                 ("rlfn", 8, (256, 256), "config4 RLFN 8 img/GPU 256x256"),
                                 ("bsrn", 16, (270, 480), "config5 BSRN 16 img/GPU 270x480"),
                                 ("imdn", 1, (256, 256), "config1-shape IMDN b1 256x256"),
                                 ("rfdn", 16, (256, 256), "RFDN b16 256x256")]:
        wts = weights(arch)
        dr = O.MODELS[IDS[arch]]["data_range"]
        m = build_model(IDS[arch], state_dict=wts).eval().to(dev)
        x16 = (torch.rand(B, 3, h, w, device=dev) * dr).half()
        y = m(x16)
        ms = timeit(lambda: m.engine(dev).forward(x16, out=y))
        r = {"what": tag, "impl": "esr_b200 fp16", "ms": ms, "img_s": B / ms * 1e3}
        print(json.dumps(r), flush=True)
        if arch in ("rfdn", "imdn") and B == 1:
            x32 = x16.float()
            y32 = m(x32)
            ms = timeit(lambda: m.engine(dev).forward(x32, out=y32))
            print(json.dumps({"what": tag, "impl": "esr_b200 fp32 (parity mode)", "ms": ms, "img_s": B / ms * 1e3}), flush=True)
        # PyTorch eager on the same GPU (cudnn.benchmark off like test_demo.py:490)
        for dt, name in [(torch.float32, "torch eager fp32"), (torch.float16, "torch eager fp16")]:
            wt = {k: torch.as_tensor(v).to(dev).to(dt) for k, v in wts.items()}
            xx = x16.to(dt)
            with torch.no_grad():
                ms = timeit(lambda: OT.FORWARD[arch](wt, xx))
            print(json.dumps({"what": tag, "impl": name + " (ATen ops of the reference modules, B200)", "ms": ms, "img_s": B / ms * 1e3}), flush=True)
    # config 3: batch 32 of DIV2K-shaped LR images, shape-bucketed, no padding
    m = build_model(0, state_dict=weights("rfdn")).eval().to(dev)
    shapes = div2k_shapes(32)
    imgs = [(torch.rand(3, h, w, device=dev) * 255).half() for h, w in shapes]
    ms = timeit(lambda: forward_bucketed(m, imgs), min_s=1.0, warm=2)
    print(json.dumps({"what": "config3 RFDN b32 DIV2K-shaped LR (bucketed)", "impl": "esr_b200 fp16", "ms": ms, "img_s": 32 / ms * 1e3,
                      "buckets": len(set(shapes))}), flush=True)
    x = (torch.rand(1, 3, 339, 510, device=dev) * 255).half()
    y = m(x)
    ms = timeit(lambda: m.engine(dev).forward(x, out=y))
    print(json.dumps({"what": "RFDN b1 339x510", "impl": "esr_b200 fp16", "ms": ms, "img_s": 1 / ms * 1e3}), flush=True)


if __name__ == "__main__":
    main()
