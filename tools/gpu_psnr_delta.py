"""Developer probe (GPU box): fp16 engine vs the fp32 reference (oracle torch port) on pseudo pairs.

Protocol A (the DIV2K protocol): HR = test.bmp and its flips / transpose (256x256), LR = uint8 of the
MATLAB-bicubic x1/4 (utils_image.imresize_np), SR through uint2tensor4 -> model -> tensor2uint, PSNR(border 4).
Protocol B (round 1): LR = test.bmp (256x256), pseudo-HR = 4x pixel replication.
Prints, per network, the float-domain and the uint8-domain PSNR deltas (ours fp16 - reference fp32).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ntire2022_esr_b200 import build_model  # noqa: E402
from oracle import esr_oracle as O  # noqa: E402
from oracle import esr_oracle_torch as OT  # noqa: E402


def psnr_f(a, b, peak):
    return 10 * np.log10(peak * peak / np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))


def main():
    img = np.load(os.path.join(ROOT, "tests", "golden", "test_bmp.npz"))["img"]
    views = [img, img[:, ::-1], img[::-1], img.transpose(1, 0, 2)]
    out = {}
    for mid, arch in [(0, "rfdn"), (4, "rlfn"), (-1, "imdn"), (18, "bsrn"), (22, "rfdn40"), (40, "rfdn_pruned"), (26, "imdn_nb7")]:
        w = O.load_weights(os.path.join(ROOT, "tests", "golden", "weights", O.MODELS[mid]["weights"] + ".npz"))
        dr = O.MODELS[mid]["data_range"]
        m = build_model(mid, state_dict=w).eval().cuda()
        wt = OT.prepare(w)
        res = {}
        for proto in ("A_imresize", "B_replicate"):
            dfl, du8, p16 = [], [], []
            for v in views:
                v = np.ascontiguousarray(v)
                if proto == "A_imresize":
                    hr = v
                    lr = np.clip(np.round(O.imresize_np(hr.astype(np.float32) / 255.0, 1 / 4) * 255.0), 0, 255).astype(np.uint8)
                else:
                    lr = v
                    hr = np.repeat(np.repeat(lr, 4, axis=0), 4, axis=1)
                x = O.uint2tensor4(lr, dr)
                ref = OT.forward(O.MODELS[mid]["arch"], wt, x).numpy()
                y16 = m(torch.from_numpy(x).cuda().half()).float().cpu().numpy()
                hr_f = hr.astype(np.float64).transpose(2, 0, 1)[None] * (dr / 255.0)
                dfl.append(psnr_f(y16, hr_f, dr) - psnr_f(ref, hr_f, dr))
                du8.append(O.psnr(O.tensor2uint(y16, dr), hr, border=4) - O.psnr(O.tensor2uint(ref, dr), hr, border=4))
                p16.append(psnr_f(y16, ref, dr))
            res[proto] = {"float_delta_db": [round(float(d), 6) for d in dfl], "uint8_delta_db_mean": round(float(np.mean(du8)), 6),
                          "uint8_delta_db": [round(float(d), 6) for d in du8], "psnr_vs_ref_db": [round(float(p), 2) for p in p16]}
        out[arch] = res
        print("PSNRDELTA", arch, json.dumps(res), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "psnr_delta.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
