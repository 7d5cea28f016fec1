"""Developer probe: fp16 accuracy of the engine paths on test.bmp (256x256) vs the engine's own fp32 mode
(which is pinned to the reference at 1e-5): PSNR(out16, out32) for the tcgen05 path and the CUDA-core fp16 path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import esr_oracle as O
from ntire2022_esr_b200 import build_model

img = np.load(os.path.join(ROOT, "tests", "golden", "test_bmp.npz"))["img"]
for mid, arch in [(0, "rfdn"), (4, "rlfn"), (-1, "imdn"), (18, "bsrn")]:
    w = O.load_weights(os.path.join(ROOT, "tests", "golden", "weights", O.MODELS[mid]["weights"] + ".npz"))
    dr = O.MODELS[mid]["data_range"]
    x = torch.from_numpy(O.uint2tensor4(img, dr)).cuda()
    m = build_model(mid, state_dict=w).eval().cuda()
    y32 = m(x).double()
    res = {}
    for name, tc in [("tcgen05", 1), ("cuda-core-fp16", 0)]:
        m2 = build_model(mid, state_dict=w).eval().cuda()
        m2.set_engine_option("tc_enable", tc)
        y16 = m2(x.half()).double()
        mse = torch.mean((y16 - y32) ** 2).item() / dr ** 2
        res[name] = 10 * np.log10(1 / mse)
    # fp16-rounded weights on the CUDA-core path: isolates weight rounding from the algebraic folds
    wh = {k: v.astype(np.float16).astype(np.float32) for k, v in w.items()}
    m3 = build_model(mid, state_dict=wh).eval().cuda()
    m3.set_engine_option("tc_enable", 0)
    y16 = m3(x.half()).double()
    mse = torch.mean((y16 - y32) ** 2).item() / dr ** 2
    res["cuda-core-fp16+fp16 weights"] = 10 * np.log10(1 / mse)
    print(arch, {k: round(v, 2) for k, v in res.items()}, flush=True)
