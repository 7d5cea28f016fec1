"""Generate golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

Imports the reference's own `test_demo.select_model` / `forward` (SURVEY Appendix A recipe: stub
matplotlib, force map_location, chdir to the reference root) and writes, next to this file:

  weights/<name>.npz      state-dicts of the four in-scope checkpoints as fp32 numpy arrays
  test_bmp.npz            utils/test.bmp as uint8 HWC RGB (via the reference's imread_uint)
  metrics_golden.npz      the reference's calculate_psnr / calculate_ssim / modcrop on a deterministic image pair (-101)
  test_bmp_lr.npz         the reference's imresize_np (MATLAB bicubic) of test.bmp: x1/4 (64x64), an odd crop, x2
  ref_<arch>_small.npz    seeded small inputs + FULL reference outputs (several odd sizes)
  ref_<arch>_256.npz      reference output on test.bmp (256x256): crops, strided subsample, stats
  ref_rfdn_tiled.npz      reference tiled forward (tile=32, overlap=8) on a 48x40 input
  ref_<arch>_<H>x<W>.npz  BASELINE.json config shapes (RFDN 339x510 = configs[2], BSRN 270x480 = configs[4]): numpy-seeded
                          input (regenerated from the seed at test time), crops + strided subsample + sums of the output

    python tests/golden/make_golden.py 22         # only the listed model ids

Nothing at test/bench time reads /root/reference; these files are the pin.
"""
import copy
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("ESR_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

for m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
os.chdir(REF)
_load = torch.load
torch.load = lambda p, **k: _load(p, **{"map_location": "cpu", **k})
import test_demo  # noqa: E402
from utils import utils_image as util  # noqa: E402

torch.set_num_threads(os.cpu_count())

MODELS = {-1: ("imdn", "imdn_baseline.pth"), 0: ("rfdn", "rfdn_baseline.pth"),
          4: ("rlfn", "team04_rlfn.pth"), 18: ("bsrn", "team18_bsrn.pth"),
          22: ("rfdn40", "team22_rep_rfdn.pth"), 40: ("rfdn_pruned", "team40_rfdn_pruned.pth"),
          26: ("imdn_nb7", "team26_imdn_nb7.pth"), 3: ("fmen", "team03_fmen.pth")}   # id 22 = RFDN at nf = 40 (test_demo.py:175-181): same graph, SURVEY row N1
SMALL_SIZES = [(15, 15), (24, 20), (33, 47), (64, 64)]
CROPS = [(0, 0), (0, 992), (992, 0), (992, 992), (500, 500), (0, 480), (700, 0), (301, 777)]
SHAPED = {0: (339, 510), 18: (270, 480)}   # model id -> LR shape of its BASELINE.json config


def shaped_input(mid, h, w, data_range):
    """The seeded input of the BASELINE-shape goldens (tests regenerate it with the same two lines)."""
    rng = np.random.default_rng(1000 + mid)
    return (rng.random((1, 3, h, w), dtype=np.float32) * np.float32(data_range)).astype(np.float16).astype(np.float32)


def main():
    only = [int(a) for a in sys.argv[1:]]   # optional: model ids to (re)generate; default = all
    os.makedirs(os.path.join(HERE, "weights"), exist_ok=True)
    img = util.imread_uint(os.path.join(REF, "utils", "test.bmp"), n_channels=3)
    if not only:
        np.savez_compressed(os.path.join(HERE, "test_bmp.npz"), img=img)
    if not only or -100 in only:
        # the DIV2K LR protocol on the one natural image the reference ships: LR = imresize_np(HR / 255, 1/4)
        # (utils/utils_image.py:704-774), also on an odd-sized crop; pins the oracle's restatement
        hr = img.astype(np.float32) / 255.
        np.savez_compressed(os.path.join(HERE, "test_bmp_lr.npz"), lr64=util.imresize_np(hr.copy(), 1 / 4),
                            lr_odd=util.imresize_np(hr[:130, :77].copy(), 1 / 4), up2=util.imresize_np(hr[:40, :36].copy(), 2))
    if not only or -101 in only:
        # the harness metrics (row N3): the reference's calculate_psnr / calculate_ssim / modcrop on a deterministic pair
        # (test.bmp and a perturbed copy), borders 0 and 4, RGB and grey  -> metrics_golden.npz
        rng = np.random.default_rng(5)
        a = img[:203, :187].copy()
        b = np.clip(a.astype(np.int32) + rng.integers(-9, 10, a.shape), 0, 255).astype(np.uint8)
        np.savez_compressed(os.path.join(HERE, "metrics_golden.npz"), a=a, b=b,
                            psnr=np.array([util.calculate_psnr(a, b, border=k) for k in (0, 4)]),
                            ssim=np.array([util.calculate_ssim(a, b, border=k) for k in (0, 4)]),
                            ssim_grey=np.float64(util.calculate_ssim(a[..., 0], b[..., 0], border=4)),
                            modcrop_shape=np.array(util.modcrop(a, 4).shape))
    for mid, (arch, fname) in MODELS.items():
        if only and mid not in only:
            continue
        args = types.SimpleNamespace(model_id=mid)
        model, name, data_range, tile = test_demo.select_model(args, torch.device("cpu"))
        sd = {k: v.detach().cpu().numpy().astype(np.float32) for k, v in model.state_dict().items()}
        np.savez(os.path.join(HERE, "weights", fname.replace(".pth", ".npz")), **sd)
        small = {"data_range": np.float32(data_range), "name": np.array(name)}
        for i, (h, w) in enumerate(SMALL_SIZES):
            g = torch.Generator().manual_seed(100 + i)
            nb = 2 if i == 1 else 1
            x = torch.rand(nb, 3, h, w, generator=g) * data_range
            with torch.no_grad():
                y = test_demo.forward(x, model, tile)
            small[f"x{i}"] = x.numpy()
            small[f"y{i}"] = y.numpy()
        np.savez_compressed(os.path.join(HERE, f"ref_{arch}_small.npz"), **small)

        x = util.uint2tensor4(img, data_range)
        with torch.no_grad():
            y = test_demo.forward(x, model, tile).numpy().copy()
            y64 = copy.deepcopy(model).double()(x.double()).numpy().copy()
        big = {"data_range": np.float32(data_range),
               "crops_yx": np.array(CROPS, dtype=np.int32),
               "crops": np.stack([y[0, :, a:a + 32, b:b + 32] for a, b in CROPS]),
               "sub16": y[0, :, ::16, ::16].copy(),
               "sum_c": y.astype(np.float64).sum(axis=(0, 2, 3)),
               "abs_sum_c": np.abs(y.astype(np.float64)).sum(axis=(0, 2, 3)),
               "minmax": np.array([y.min(), y.max()], dtype=np.float32),
               "uint8_sub": util.tensor2uint(torch.from_numpy(y.copy()), data_range)[::8, ::8].copy()}
        if y64 is not None:
            big["fp32_vs_fp64_maxabs"] = np.float64(np.abs(y - y64).max())
        np.savez_compressed(os.path.join(HERE, f"ref_{arch}_256.npz"), **big)
        print(name, "ok; fp32-vs-fp64 max abs", big.get("fp32_vs_fp64_maxabs"))

        if mid in SHAPED:
            h, w = SHAPED[mid]
            xs = shaped_input(mid, h, w, data_range)   # fp16-representable values: the fp16 engine sees the same input
            with torch.no_grad():
                ys = test_demo.forward(torch.from_numpy(xs), model, tile).numpy().copy()
            Ho, Wo = 4 * h, 4 * w
            cr = [(0, 0), (0, Wo - 32), (Ho - 32, 0), (Ho - 32, Wo - 32), (Ho // 2, Wo // 2), (4 * 127, 4 * 128 - 16), (4 * 255, 4 * 383)]
            cr = [(min(a, Ho - 32), min(b, Wo - 32)) for a, b in cr]
            np.savez_compressed(os.path.join(HERE, f"ref_{arch}_{h}x{w}.npz"), data_range=np.float32(data_range),
                                seed=np.int64(1000 + mid), crops_yx=np.array(cr, dtype=np.int32),
                                crops=np.stack([ys[0, :, a:a + 32, b:b + 32] for a, b in cr]), sub16=ys[0, :, ::16, ::16].copy(),
                                sum_c=ys.astype(np.float64).sum(axis=(0, 2, 3)),
                                sq_sum_c=(ys.astype(np.float64) ** 2).sum(axis=(0, 2, 3)),
                                minmax=np.array([ys.min(), ys.max()], dtype=np.float32))
            print(name, "shaped golden", h, w)

        if arch == "rfdn":
            g = torch.Generator().manual_seed(7)
            xt = torch.rand(1, 3, 48, 40, generator=g) * data_range
            with torch.no_grad():
                yt = test_demo.forward(xt, model, tile=32, tile_overlap=8)
            np.savez_compressed(os.path.join(HERE, "ref_rfdn_tiled.npz"), x=xt.numpy(), y=yt.numpy())


if __name__ == "__main__":
    main()
