"""FMEN (model id 3) without a GPU: the packed tcgen05 layers (esr_debug_tc_layer) are replayed on the CPU in the
network's topology (models/team03_fmen.py:68-75, 121-134) and the result is compared with the oracle.  This pins what
build_fmen adds on top of a plain sequence of 3x3 convolutions: the power-of-two range management (trunk at 2^-10, the
warm-up HFAB's inner path at a further 2^-4, undone in the excitate and tail weights) and the gate epilogue
(res_after = 2: out = sigmoid(acc + bias) * operand)."""
import os

import numpy as np

from oracle import esr_oracle as O
from test_tc_packing_cpu import _layers, _replay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S_T, S_H0 = 2.0 ** -10, 2.0 ** -4


def _w():
    return O.load_weights(os.path.join(ROOT, "tests", "golden", "weights", "team03_fmen.npz"))


def _h(a):
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float16).astype(np.float64)


def _run(L, x, gate=None):
    """one packed layer on an (H, W, 64) activation: replay the MMA entries, bias, activation / gate from the group record"""
    col0, ncols, act, has_res, res_after, mode, slope_bits, b9 = L["groups"][0]
    acc = _replay(L, x)[..., :ncols] + L["bias"][0, :ncols]
    slope = np.array([slope_bits], dtype=np.int32).view(np.float32)[0]
    if act == 1:
        acc = np.where(acc >= 0, acc, acc * float(slope))
    if has_res:
        assert gate is not None
        acc = gate[..., :ncols] / (1.0 + np.exp(-acc)) if res_after == 2 else acc + gate[..., :ncols]
    out = np.zeros(x.shape[:2] + (64,))
    out[..., :min(ncols, 64)] = acc[..., :64]
    return out, acc


def test_packed_fmen_graph_reproduces_the_oracle_and_the_scales_are_where_they_should_be():
    w = _w()
    Ls = _layers("fmen", w)
    assert len(Ls) == 33                                     # every convolution but the head
    # --- the scales, layer by layer: biases carry the scale of the layer's output, the two un-scaling layers none
    for name, L in Ls.items():
        b = w[name + ".bias"].astype(np.float64)
        got = L["bias"][0, :len(b)]
        if name.endswith("excitate") or name == "tail.0":
            want = b
        elif name.startswith("warmup.1."):
            want = b * S_T * S_H0
        else:
            want = b * S_T
        np.testing.assert_allclose(got, want.astype(np.float32), rtol=0, atol=0, err_msg=name)
        g = L["groups"][0]
        is_gate = name.endswith("excitate")
        assert (g[3] == 1 and g[4] == 2) == is_gate, name      # gate operand only on the excitate layers
        if name == "lr_conv":
            assert g[3] == 1 and g[4] == 0                     # + head output, before (no) activation
        if name == "tail.0":
            assert g[5] == 1 and L["acc_cols"] == 48          # pixel-shuffle store
    # --- the whole graph on a small image, fp64 arithmetic on the packed (fp16) weights
    rng = np.random.default_rng(0)
    x = rng.random((1, 3, 12, 10)) * 255.0
    wt = {k: v.astype(np.float64) for k, v in w.items()}
    fea_true = O.conv2d(x, wt["head.weight"], wt["head.bias"], 1, 1)[0].transpose(1, 2, 0)       # the head runs on the CUDA cores in fp32
    fea = np.zeros((12, 10, 64))
    fea[..., :50] = fea_true * S_T

    def hfab(p, up, xin):
        t, _ = _run(Ls[p + "squeeze"], xin)
        for k in range(up):
            t, _ = _run(Ls[p + f"convs.{k}.conv1.rep_conv"], t)
            t, _ = _run(Ls[p + f"convs.{k}.conv2.rep_conv"], t)
        return _run(Ls[p + "excitate"], t, gate=xin)[0]

    g0, _ = _run(Ls["warmup.0"], fea)
    h = hfab("warmup.1.", 2, g0)
    for i in range(4):
        t, _ = _run(Ls[f"basic_blocks.{i}.conv1.rep_conv"], h)
        gi, _ = _run(Ls[f"basic_blocks.{i}.conv2.rep_conv"], t)
        h = hfab(f"hfabs.{i}.", 1, gi)
    t, _ = _run(Ls["lr_conv"], h, gate=fea)
    _, y48 = _run(Ls["tail.0"], t)                               # (H, W, 48) at TRUE scale
    got = O.pixel_shuffle(y48.transpose(2, 0, 1)[None], 4)
    # reference: the oracle on weights rounded the way the packer rounds them (scaled by a power of two, then fp16)
    wr = {}
    for k, v in wt.items():
        if not k.endswith(".weight") or k == "head.weight":
            wr[k] = v
            continue
        n = k[:-len(".weight")]
        s = 1.0
        if n.endswith("excitate"):
            s = 1.0 / (S_T * (S_H0 if n.startswith("warmup.1.") else 1.0))
        elif n == "tail.0":
            s = 1.0 / S_T
        elif n == "warmup.1.squeeze":
            s = S_H0
        wr[k] = _h(v * s) / s
    ref = O.fmen_forward(wr, x, dtype=np.float64)
    assert got.shape == ref.shape == (1, 3, 48, 40)
    assert np.abs(got - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max()), np.abs(got - ref).max()
    # and the un-rounded reference is only the fp16 weight rounding away
    exact = O.fmen_forward(wt, x, dtype=np.float64)
    assert np.abs(got - exact).max() <= 2e-2 * np.abs(exact).max()
