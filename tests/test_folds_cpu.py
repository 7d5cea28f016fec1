"""The algebraic folds the graph builder applies before packing weights (csrc/graph_builder.cuh), restated in
numpy on the reference's real weights and checked in fp64 against the oracle's layer-by-layer evaluation.  They
are exact identities; the GPU parity tests cover the packed kernels, these pin the algebra itself on the CPU:

  * BSConvU (Linear, then zero-padded depthwise 3x3) == one dense 3x3 convolution with a border-class bias
  * ESA tail: conv4 commuted through the bilinear upsample, cf' = conv4(conv_f(c1_)) + b4
  * the fast GELU used by the fp16 engine (Abramowitz-Stegun 7.1.26 erf) against the exact-erf GELU
"""
import os

import numpy as np

from oracle import esr_oracle as O


def _w(golden_dir, mid):
    w = O.load_weights(os.path.join(golden_dir, "weights", O.MODELS[mid]["weights"] + ".npz"))
    return {k: v.astype(np.float64) for k, v in w.items()}


def test_bsconvu_as_dense_3x3_with_border_class_bias(golden_dir):
    """graph_builder.cuh::bsconv_dense.  models/team18_bsrn.py:82-88."""
    w = _w(golden_dir, 18)
    rng = np.random.default_rng(0)
    for name, shape in [("B2.c1_r.", (2, 48, 7, 9)), ("B4.c4.", (1, 48, 5, 3)), ("c2.", (1, 48, 3, 8))]:
        x = rng.standard_normal(shape)
        ref = O._bsconvu(w, name, x)
        pw, pb = w[name + "pw.weight"], w[name + "pw.bias"]                      # (O, I), (O,)
        dw, db = w[name + "dw.weight"][:, 0], w[name + "dw.bias"]                 # (O, 3, 3), (O,)
        dense = dw[:, None, :, :] * pw[:, :, None, None]                           # W[o][i][ky][kx]
        y = O.conv2d(x, dense, None, 1, 1)
        H, Wd = x.shape[2:]
        for yy in range(H):
            for xx in range(Wd):
                ky = slice(1 if yy == 0 else 0, 2 if yy == H - 1 else 3)          # taps that lie inside the image
                kx = slice(1 if xx == 0 else 0, 2 if xx == Wd - 1 else 3)
                y[:, :, yy, xx] += db + pb * dw[:, ky, kx].sum(axis=(1, 2))
        assert np.abs(y - ref).max() < 1e-12 * max(1.0, np.abs(ref).max()), name


def test_esa_tail_commuted_through_the_bilinear_upsample(golden_dir):
    """graph_builder.cuh::esa_tail_commuted / build_rfdn: M3 = conv4(c3) without bias on the pooled map, cf' from
    the c5 GEMM, mask = sigmoid(bilinear(M3) + cf').  models/rfdn_baseline/block.py:117-129."""
    w = _w(golden_dir, 0)
    p = "B3.esa."
    x = np.random.default_rng(1).standard_normal((1, 50, 33, 41)) * 20.0
    ref = O._esa_rfdn(w, p, x)
    c1_ = O._conv(w, p + "conv1", x)
    c1 = O._conv(w, p + "conv2", c1_, stride=2, padding=0)
    v = O.relu(O._conv(w, p + "conv_max", O.max_pool2d(c1, 7, 3), padding=1))
    c3 = O._conv(w, p + "conv3_", O.relu(O._conv(w, p + "conv3", v, padding=1)), padding=1)
    w4, b4 = w[p + "conv4.weight"], w[p + "conv4.bias"]
    m3 = O.conv2d(c3, w4, None)                                                   # low resolution, no bias
    cfp = O.conv2d(O._conv(w, p + "conv_f", c1_), w4, b4)                          # full resolution, carries b4
    y = x * O.sigmoid(O.interpolate_bilinear(m3, x.shape[2:]) + cfp)
    assert np.abs(y - ref).max() < 1e-11 * np.abs(ref).max()


def test_fast_gelu_of_the_fp16_engine():
    """kernels_generic.cuh::gelu_fast / conv_tc.cuh::tc_gelu16: erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7);
    its GELU must sit far below the fp16 rounding (relative 4.9e-4) that follows it in the engine."""
    v = np.linspace(-12.0, 12.0, 200001)
    x = np.abs(v) * 0.70710678118654752440
    t = 1.0 / (1.0 + 0.3275911 * x)
    q = ((((1.061405429 * t - 1.453152027) * t + 1.421413741) * t - 0.284496736) * t + 0.254829592) * t
    e = 1.0 - q * np.exp(-x * x)
    fast = 0.5 * v * (1.0 + np.copysign(e, v))
    ref = O.gelu(v)
    assert np.abs(fast - ref).max() < 1e-6
    big = np.abs(ref) > 1e-2
    assert (np.abs(fast - ref)[big] / np.abs(ref)[big]).max() < 2e-5
