"""The reference's evaluation harness on top of the B200 engine: `select_dataset`, `run`, `main`
and the image utilities they use, with the reference's names, argument meaning, result keys and file outputs
(test_demo.py:344-361, 394-477, 480-577; utils/utils_image.py:122-141, 442-455, 490-554).

    python -m ntire2022_esr_b200.harness --data_dir <DIV2K root> --save_dir <out> --model_id 0 [--include_test] [--ssim]

What differs from the reference, on purpose:
 * the network runs in the CUDA engine (`demo_api.select_model` / `forward`); the LR image goes in as uint8 HWC and the SR
   image comes back as uint8 HWC (`uint2tensor4` and `tensor2uint` run on the device, `demo_api.forward_uint8`) unless the
   model asks for tiles, in which case the reference's float pipeline is used as is;
 * `--half` (default on) selects the fp16 engine; `--no-half` the fp32 parity mode;
 * #Activations / #Conv2d / FLOPs come from the same forward hooks the reference registers (`model_summary` below restates
   utils/model_summary.py:230-318, 390-438); the drop-in serves them with shape-only tensors (module_tree.py), so no
   second forward pass is needed.
PSNR / SSIM / modcrop / image I/O are host-side numpy + OpenCV like the reference's (cv2 is the reference's own choice)."""
from __future__ import annotations

import argparse
import json
import logging
import math
import os
from pprint import pprint

import numpy as np
import torch
import torch.nn as nn

from . import demo_api


# ---------------------------------------------------------------------------------------------------------------------
# image utilities (utils/utils_image.py)
# ---------------------------------------------------------------------------------------------------------------------
def imread_uint(path: str, n_channels: int = 3) -> np.ndarray:
    """utils_image.py:122-134: HxWx3 RGB uint8 (or HxWx1 for n_channels = 1)."""
    import cv2
    if n_channels == 1:
        return np.expand_dims(cv2.imread(path, 0), axis=2)
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 2:
        return cv2.cvtColor(img, cv2.COLOR_GRAY2RGB)
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def imsave(img: np.ndarray, img_path: str) -> None:
    """utils_image.py:137-141"""
    import cv2
    img = np.squeeze(img)
    if img.ndim == 3:
        img = img[:, :, [2, 1, 0]]
    cv2.imwrite(img_path, img)


def modcrop(img_in: np.ndarray, scale: int) -> np.ndarray:
    """utils_image.py:442-455: crop H and W down to multiples of `scale`."""
    img = np.copy(img_in)
    if img.ndim not in (2, 3):
        raise ValueError("Wrong img ndim: [{:d}].".format(img.ndim))
    H, W = img.shape[:2]
    return img[:H - H % scale, :W - W % scale]


def calculate_psnr(img1: np.ndarray, img2: np.ndarray, border: int = 0) -> float:
    """utils_image.py:490-503; images in [0, 255]."""
    if img1.shape != img2.shape:
        raise ValueError("Input images must have the same dimensions.")
    h, w = img1.shape[:2]
    a = img1[border:h - border, border:w - border].astype(np.float64)
    b = img2[border:h - border, border:w - border].astype(np.float64)
    mse = np.mean((a - b) ** 2)
    return float("inf") if mse == 0 else 20 * math.log10(255.0 / math.sqrt(mse))


def ssim(img1: np.ndarray, img2: np.ndarray) -> float:
    """utils_image.py:533-554: 11x11 Gaussian window (sigma 1.5), 'valid' region."""
    import cv2
    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    a, b = img1.astype(np.float64), img2.astype(np.float64)
    k = cv2.getGaussianKernel(11, 1.5)
    win = np.outer(k, k.transpose())
    f = lambda v: cv2.filter2D(v, -1, win)[5:-5, 5:-5]
    mu1, mu2 = f(a), f(b)
    s1, s2, s12 = f(a * a) - mu1 * mu1, f(b * b) - mu2 * mu2, f(a * b) - mu1 * mu2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    return float(m.mean())


def calculate_ssim(img1: np.ndarray, img2: np.ndarray, border: int = 0) -> float:
    """utils_image.py:509-530.  (For RGB input the reference averages three evaluations of the WHOLE 3-channel pair -
    its loop never indexes the channel - which equals one evaluation; kept.)"""
    if img1.shape != img2.shape:
        raise ValueError("Input images must have the same dimensions.")
    h, w = img1.shape[:2]
    a, b = img1[border:h - border, border:w - border], img2[border:h - border, border:w - border]
    if a.ndim == 2:
        return ssim(a, b)
    if a.ndim == 3 and a.shape[2] == 3:
        return float(np.mean([ssim(a, b) for _ in range(3)]))
    if a.ndim == 3 and a.shape[2] == 1:
        return ssim(np.squeeze(a), np.squeeze(b))
    raise ValueError("Wrong input image dimensions.")


# ---------------------------------------------------------------------------------------------------------------------
# model summary (utils/model_summary.py)
# ---------------------------------------------------------------------------------------------------------------------
def model_summary(model: nn.Module, input_dim=(3, 256, 256)):
    """(#activations, #Conv2d, FLOPs, #parameters) as `get_model_activation` / `get_model_flops(..., False)` / the parameter
    count of test_demo.py:522-533 report them: conv FLOPs = k*k*Cin*(Cout/groups) MACs per output position (bias not
    counted), one FLOP per ReLU / LeakyReLU output element, in*out per nn.Linear row (model_summary.py:283-318, 428-438)."""
    from .module_tree import fire_forward_hooks
    c = {"flops": 0, "act": 0, "nconv": 0}

    def conv_hook(mod, inp, out):
        c["flops"] += int(np.prod(mod.kernel_size) * mod.in_channels * (mod.out_channels // mod.groups)) * int(out.shape[0] * np.prod(out.shape[2:]))
        c["act"] += out.numel()
        c["nconv"] += 1

    def relu_hook(mod, inp, out):
        c["flops"] += out.numel()

    def linear_hook(mod, inp, out):
        c["flops"] += int(inp[0].shape[0] * inp[0].shape[1] * out.shape[1])

    handles = []
    for mod in model.modules():
        if isinstance(mod, nn.Conv2d):
            handles.append(mod.register_forward_hook(conv_hook))
        elif isinstance(mod, (nn.ReLU, nn.LeakyReLU)):
            handles.append(mod.register_forward_hook(relu_hook))
        elif isinstance(mod, nn.Linear):
            handles.append(mod.register_forward_hook(linear_hook))
    try:
        if isinstance(model, demo_api.B200SRModel):
            fire_forward_hooks(model, model.arch, model.nf, model.nblocks, model._esa_f, 1, input_dim[1], input_dim[2])
        else:
            with torch.no_grad():
                model(torch.zeros(1, *input_dim, device=next(model.parameters()).device))
    finally:
        for h in handles:
            h.remove()
    return c["act"], c["nconv"], c["flops"], sum(p.numel() for p in model.parameters())


# ---------------------------------------------------------------------------------------------------------------------
# test_demo.py
# ---------------------------------------------------------------------------------------------------------------------
def select_dataset(data_dir: str, mode: str):
    """test_demo.py:344-361: (LR path, HR path) pairs of DIV2K test (0901-1000) or validation (0801-0900, LR named *x4.png)."""
    if mode == "test":
        return [(os.path.join(data_dir, f"DIV2K_test_LR/{i:04}.png"), os.path.join(data_dir, f"DIV2K_test_HR/{i:04}.png"))
                for i in range(901, 1001)]
    return [(os.path.join(data_dir, f"DIV2K_valid_LR/{i:04}x4.png"), os.path.join(data_dir, f"DIV2K_valid_HR/{i:04}.png"))
            for i in range(801, 901)]


def run(model, model_name, data_range, tile, logger, device, args, mode="test"):
    """test_demo.py:394-477: per image read LR, super-resolve (timed with CUDA events around the forward only), convert to
    uint8, PSNR (and SSIM) against the mod-cropped HR with a border of `sf` pixels, save the SR image; returns the
    reference's result dictionary (`{mode}_runtime` [ms], `{mode}_psnr`, `{mode}_ssim`, `{mode}_memory` [MB],
    `{mode}_ave_*`).  Pairs whose files do not exist are skipped (the reference would raise), so a partial dataset works."""
    sf = 4
    border = sf
    results = {f"{mode}_runtime": [], f"{mode}_psnr": []}
    if args.ssim:
        results[f"{mode}_ssim"] = []
    pairs = [p for p in select_dataset(args.data_dir, mode) if os.path.exists(p[0]) and os.path.exists(p[1])]
    if not pairs:
        raise FileNotFoundError(f"no DIV2K {mode} pairs under {args.data_dir}")
    save_path = os.path.join(args.save_dir, model_name, "test" if mode == "test" else "valid")
    os.makedirs(save_path, exist_ok=True)
    half = getattr(args, "half", True)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for lr_path, hr_path in pairs:
        img_name, ext = os.path.splitext(os.path.basename(hr_path))
        img_lr = imread_uint(lr_path, n_channels=3)
        if tile is None and isinstance(model, demo_api.B200SRModel):
            x = torch.from_numpy(np.ascontiguousarray(img_lr)).to(device)
            start.record()
            img_sr = demo_api.forward_uint8(x, model, data_range, half=half)
            end.record()
            torch.cuda.synchronize()
            img_sr = img_sr.cpu().numpy()
        else:   # the reference's float pipeline (utils_image.py:190-193, 204-208)
            x = torch.from_numpy(np.ascontiguousarray(img_lr)).permute(2, 0, 1).float().div(255. / data_range).unsqueeze(0).to(device)
            x = x.half() if half else x
            start.record()
            y = demo_api.forward(x, model, tile)
            end.record()
            torch.cuda.synchronize()
            y = y.data.squeeze().float().clamp_(0, data_range).cpu().numpy()
            img_sr = np.uint8((np.transpose(y, (1, 2, 0)) * 255.0 / data_range).round())
        results[f"{mode}_runtime"].append(start.elapsed_time(end))
        img_hr = modcrop(imread_uint(hr_path, n_channels=3).squeeze(), sf)
        psnr = calculate_psnr(img_sr, img_hr, border=border)
        results[f"{mode}_psnr"].append(psnr)
        if args.ssim:
            s = calculate_ssim(img_sr, img_hr, border=border)
            results[f"{mode}_ssim"].append(s)
            logger.info("{:s} - PSNR: {:.2f} dB; SSIM: {:.4f}.".format(img_name + ext, psnr, s))
        else:
            logger.info("{:s} - PSNR: {:.2f} dB".format(img_name + ext, psnr))
        imsave(img_sr, os.path.join(save_path, img_name[:4] + ext))
    n = len(results[f"{mode}_runtime"])
    results[f"{mode}_memory"] = torch.cuda.max_memory_allocated(torch.cuda.current_device()) / 1024 ** 2
    results[f"{mode}_ave_runtime"] = sum(results[f"{mode}_runtime"]) / n
    results[f"{mode}_ave_psnr"] = sum(results[f"{mode}_psnr"]) / n
    if args.ssim:
        results[f"{mode}_ave_ssim"] = sum(results[f"{mode}_ssim"]) / n
    logger.info("{:>16s} : {:<.3f} [M]".format("Max Memery", results[f"{mode}_memory"]))
    logger.info("------> Average runtime of ({}) is : {:.6f} seconds".format("test" if mode == "test" else "valid", results[f"{mode}_ave_runtime"]))
    return results


def _results_table(results: dict, include_test: bool) -> str:
    """test_demo.py:536-562"""
    if include_test:
        fmt = "{:20s}\t{:10s}\t{:10s}\t{:14s}\t{:14s}\t{:14s}\t{:10s}\t{:10s}\t{:8s}\t{:8s}\t{:8s}\n"
        s = fmt.format("Model", "Val PSNR", "Test PSNR", "Val Time [ms]", "Test Time [ms]", "Ave Time [ms]", "Params [M]", "FLOPs [G]", "Acts [M]", "Mem [M]", "Conv")
    else:
        fmt = "{:20s}\t{:10s}\t{:14s}\t{:10s}\t{:10s}\t{:8s}\t{:8s}\t{:8s}\n"
        s = fmt.format("Model", "Val PSNR", "Val Time [ms]", "Params [M]", "FLOPs [G]", "Acts [M]", "Mem [M]", "Conv")
    for k, v in results.items():
        cols = [k, f"{v['valid_ave_psnr']:2.2f}"]
        if include_test:
            cols += [f"{v['test_ave_psnr']:2.2f}", f"{v['valid_ave_runtime']:3.2f}", f"{v['test_ave_runtime']:3.2f}",
                     f"{(v['valid_ave_runtime'] + v['test_ave_runtime']) / 2:3.2f}"]
        else:
            cols += [f"{v['valid_ave_runtime']:3.2f}"]
        cols += [f"{v['num_parameters']:2.3f}", f"{v['flops']:2.2f}", f"{v['activations']:2.2f}", f"{v['valid_memory']:2.2f}", f"{v['num_conv']:4d}"]
        s += fmt.format(*cols)
    return s


def main(args):
    """test_demo.py:480-563: results.json is read-modified-written in the working directory, results.txt rewritten."""
    logger = logging.getLogger("NTIRE2022-EfficientSR")
    if not logger.handlers:
        logger.setLevel(logging.INFO)
        fmt = logging.Formatter("%(asctime)s.%(msecs)03d : %(message)s", datefmt="%y-%m-%d %H:%M:%S")
        for h in (logging.FileHandler("NTIRE2022-EfficientSR.log", mode="a"), logging.StreamHandler()):
            h.setFormatter(fmt)
            logger.addHandler(h)
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    json_dir = os.path.join(os.getcwd(), "results.json")
    results = {}
    if os.path.exists(json_dir):
        with open(json_dir, "r") as f:
            results = json.load(f)
    model, model_name, data_range, tile = demo_api.select_model(args, device)
    logger.info(model_name)
    results[model_name] = run(model, model_name, data_range, tile, logger, device, args, mode="valid")
    if args.include_test:
        results[model_name].update(run(model, model_name, data_range, tile, logger, device, args, mode="test"))
    activations, num_conv, flops, num_parameters = model_summary(model, (3, 256, 256))
    activations, flops, num_parameters = activations / 10 ** 6, flops / 10 ** 9, num_parameters / 10 ** 6
    logger.info("{:>16s} : {:<.4f} [M]".format("#Activations", activations))
    logger.info("{:>16s} : {:<d}".format("#Conv2d", num_conv))
    logger.info("{:>16s} : {:<.4f} [G]".format("FLOPs", flops))
    logger.info("{:>16s} : {:<.4f} [M]".format("#Params", num_parameters))
    results[model_name].update({"activations": activations, "num_conv": num_conv, "flops": flops, "num_parameters": num_parameters})
    with open(json_dir, "w") as f:
        json.dump(results, f)
    with open(os.path.join(os.getcwd(), "results.txt"), "w") as f:
        f.write(_results_table(results, args.include_test))
    return results


def build_parser() -> argparse.ArgumentParser:
    """test_demo.py:566-574 (+ --half / --no-half for the engine's storage type)"""
    parser = argparse.ArgumentParser("NTIRE2022-EfficientSR")
    parser.add_argument("--data_dir", default="/cluster/work/cvl/yawli/data/NTIRE2022_Challenge", type=str)
    parser.add_argument("--save_dir", default="/cluster/work/cvl/yawli/data/NTIRE2022_Challenge/results", type=str)
    parser.add_argument("--model_id", default=0, type=int)
    parser.add_argument("--include_test", action="store_true", help="Inference on the DIV2K test set")
    parser.add_argument("--ssim", action="store_true", help="Calculate SSIM")
    parser.add_argument("--half", dest="half", action="store_true", default=True, help="fp16 engine (default)")
    parser.add_argument("--no-half", dest="half", action="store_false", help="fp32 parity mode")
    return parser


if __name__ == "__main__":
    _args = build_parser().parse_args()
    pprint(_args)
    main(_args)
