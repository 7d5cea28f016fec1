"""The evaluation harness (test_demo.py run / main / select_dataset + the image utilities they call).
CPU part: metrics against values produced by the reference's own functions (tests/golden/metrics_golden.npz), dataset
layout, result table, model summary numbers.  GPU part: a miniature DIV2K tree built from test.bmp goes through main()."""
import json
import os
import types

import numpy as np
import pytest

from oracle import esr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch = pytest.importorskip("torch")


def _golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))


def test_metrics_match_the_reference_functions():
    from ntire2022_esr_b200 import harness as H
    g = _golden("metrics_golden.npz")
    a, b = g["a"], g["b"]
    for k, border in enumerate((0, 4)):
        assert abs(H.calculate_psnr(a, b, border=border) - g["psnr"][k]) < 1e-12
        assert abs(H.calculate_ssim(a, b, border=border) - g["ssim"][k]) < 1e-12
    assert abs(H.calculate_ssim(a[..., 0], b[..., 0], border=4) - float(g["ssim_grey"])) < 1e-12
    assert H.calculate_psnr(a, a) == float("inf")
    assert tuple(H.modcrop(a, 4).shape) == tuple(g["modcrop_shape"]) == (200, 184, 3)
    with pytest.raises(ValueError):
        H.calculate_psnr(a, b[:-1])


def test_dataset_layout_and_parser_follow_the_reference():
    from ntire2022_esr_b200 import harness as H
    v = H.select_dataset("/d", "valid")
    t = H.select_dataset("/d", "test")
    assert len(v) == len(t) == 100
    assert v[0] == ("/d/DIV2K_valid_LR/0801x4.png", "/d/DIV2K_valid_HR/0801.png") and v[-1][1].endswith("0900.png")
    assert t[0] == ("/d/DIV2K_test_LR/0901.png", "/d/DIV2K_test_HR/0901.png") and t[-1][0].endswith("1000.png")
    a = H.build_parser().parse_args([])
    assert (a.model_id, a.include_test, a.ssim, a.half) == (0, False, False, True)
    assert H.build_parser().parse_args(["--model_id", "18", "--ssim", "--no-half"]).half is False


def test_model_summary_reports_the_reference_numbers():
    from test_host_cpu import REFERENCE_SUMMARY, _weights
    from ntire2022_esr_b200 import build_model
    from ntire2022_esr_b200.harness import model_summary
    for mid, want in REFERENCE_SUMMARY.items():
        assert model_summary(build_model(mid, state_dict=_weights(mid)), (3, 256, 256)) == want, mid


def test_results_table_columns():
    from ntire2022_esr_b200.harness import _results_table
    r = {"00_RFDN_baseline": {"valid_ave_psnr": 29.04, "valid_ave_runtime": 0.41, "valid_memory": 120.5, "test_ave_psnr": 28.75,
                              "test_ave_runtime": 0.43, "num_parameters": 0.433, "flops": 27.1, "activations": 112.03, "num_conv": 64}}
    lines = _results_table(r, True).splitlines()
    assert lines[0].split("\t")[:3] == ["Model               ", "Val PSNR  ", "Test PSNR "] and len(lines) == 2
    assert [c.strip() for c in lines[1].split("\t")] == ["00_RFDN_baseline", "29.04", "28.75", "0.41", "0.43", "0.42", "0.433", "27.10", "112.03", "120.50", "64"]
    assert len(_results_table(r, False).splitlines()[1].split("\t")) == 8


@pytest.mark.gpu
def test_main_on_a_miniature_div2k_tree(tmp_path, monkeypatch):
    """HR = two crops of test.bmp, LR = the reference's bicubic x1/4 (oracle.imresize_np, pinned on the reference): main()
    must write results.json / results.txt / the SR images and report the PSNR the oracle's fp32 forward gives (+- 0.01 dB)."""
    from ntire2022_esr_b200 import harness as H
    img = _golden("test_bmp.npz")["img"]
    data, save = tmp_path / "data", tmp_path / "out"
    (data / "DIV2K_valid_LR").mkdir(parents=True)
    (data / "DIV2K_valid_HR").mkdir()
    want = []
    w = O.load_weights(os.path.join(ROOT, "tests", "golden", "weights", O.MODELS[0]["weights"] + ".npz"))
    for k, (y0, x0, hh, ww) in enumerate([(0, 0, 128, 160), (90, 70, 163, 131)]):   # the second HR is not a multiple of 4: modcrop
        hr = img[y0:y0 + hh, x0:x0 + ww]
        hr4 = hr[:hh - hh % 4, :ww - ww % 4]
        lr = np.uint8((np.clip(O.imresize_np(hr4.astype(np.float32) / 255., 1 / 4), 0, 1) * 255.).round())
        H.imsave(hr, str(data / "DIV2K_valid_HR" / f"{801 + k:04}.png"))
        H.imsave(lr, str(data / "DIV2K_valid_LR" / f"{801 + k:04}x4.png"))
        assert np.array_equal(H.imread_uint(str(data / "DIV2K_valid_HR" / f"{801 + k:04}.png")), hr)       # PNG round trip, RGB order
        y = O.forward("rfdn", w, O.uint2tensor4(lr, 255.0), dtype=np.float32)
        want.append(H.calculate_psnr(O.tensor2uint(y, 255.0), hr4, border=4))
    # select_model loads model_zoo/<name>.pth relative to the working directory like the reference (test_demo.py:22-27)
    (tmp_path / "model_zoo").mkdir()
    torch.save({k: torch.from_numpy(np.asarray(v)) for k, v in w.items()}, tmp_path / "model_zoo" / "rfdn_baseline.pth")
    monkeypatch.chdir(tmp_path)
    args = types.SimpleNamespace(data_dir=str(data), save_dir=str(save), model_id=0, include_test=False, ssim=True, half=True)
    res = H.main(args)
    r = res["00_RFDN_baseline"]
    assert len(r["valid_psnr"]) == 2 and np.abs(np.array(r["valid_psnr"]) - np.array(want)).max() < 1e-2, (r["valid_psnr"], want)
    assert all(0.5 < s <= 1.0 for s in r["valid_ssim"]) and r["valid_ave_runtime"] > 0 and r["valid_memory"] >= 0
    assert (r["num_conv"], round(r["flops"], 4), round(r["num_parameters"], 6)) == (64, 27.1026, 0.433448)
    on_disk = json.load(open(tmp_path / "results.json"))
    assert on_disk["00_RFDN_baseline"]["valid_ave_psnr"] == r["valid_ave_psnr"]
    assert (tmp_path / "results.txt").read_text().startswith("Model")
    sr = H.imread_uint(str(save / "00_RFDN_baseline" / "valid" / "0801.png"))
    assert sr.shape == (128, 160, 3)
