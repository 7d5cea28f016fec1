"""Replays, on the CPU, exactly what conv_tc_kernel is handed for a tcgen05 layer (esr_debug_tc_layer: MMA entry
list, pre-swizzled fp16 weight blob, group table, biases) and compares the result with the oracle's evaluation of
the reference layers folded into that launch.  This pins the host side of the tensor-core path - weight packing,
SWIZZLE_128B block layout, tap / chunk / column-segment planning, the algebraic folds, the border-class bias -
without a GPU.  Weights are compared at their fp16-rounded values (what the blob holds); arithmetic is fp64."""
import ctypes
import os

import numpy as np
import pytest

from oracle import esr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _weights(mid):
    return O.load_weights(os.path.join(ROOT, "tests", "golden", "weights", O.MODELS[mid]["weights"] + ".npz"))


def _layers(arch, w, **kw):
    from ntire2022_esr_b200 import Engine, _cabi

    e = Engine(arch, device=-1, **kw)
    e.load_state_dict(w)
    lib = _cabi.lib
    n = lib.esr_debug_tc_layer(e._h, -1, None, 0, None, None, None, None, None, None, 0)
    out = {}
    for i in range(n):
        name = ctypes.create_string_buffer(256)
        meta = (ctypes.c_int32 * 12)()
        ent = (ctypes.c_int32 * (16 * 8))()
        grp = (ctypes.c_int32 * (3 * 8))()
        bias = (ctypes.c_float * (3 * 64))()
        bias9 = (ctypes.c_float * (9 * 64))()
        blob = (ctypes.c_uint8 * (1 << 18))()
        rc = lib.esr_debug_tc_layer(e._h, i, name, 256, meta, ent, grp, bias, bias9, blob, len(blob))
        assert rc == 0, lib.esr_last_error(e._h)
        m = list(meta)
        out[name.value.decode()] = dict(
            nchunks=m[0], halo=m[1], acc_cols=m[2], chunk_c0=m[8:12], has_bias9=bool(m[7]),
            entries=np.array(ent, dtype=np.int64).reshape(16, 8)[:m[3]],
            groups=np.array(grp, dtype=np.int32).reshape(3, 8)[:m[4]],
            bias=np.array(bias, dtype=np.float64).reshape(3, 64), bias9=np.array(bias9, dtype=np.float64).reshape(9, 64),
            blob=np.frombuffer(bytes(blob[:m[5]]), dtype=np.uint8))
    return out


def _sw128(n, k):   # tc_common.cuh::sw128_offset: byte offset of element (row n, k) in a K-major SWIZZLE_128B block
    return (n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 3) ^ (n & 7)) & 7) << 4) + (k & 7) * 2


def _replay(L, x_nhwc):
    """x_nhwc: (H, W, C) fp64 activations in the layer's input buffer (64-channel chunks).  Returns the
    accumulator (H, W, acc_cols) after all MMA entries, i.e. before bias / residual / activation."""
    H, W, _ = x_nhwc.shape
    acc = np.full((H, W, L["acc_cols"]), np.nan)
    for dy, dx, chunk, steps, n, dcol, first, boff in L["entries"]:
        blk = L["blob"][boff:boff + n * 128]
        idx = np.array([[_sw128(j, k) for k in range(64)] for j in range(n)])
        wt = (blk[idx].astype(np.uint16) | (blk[idx + 1].astype(np.uint16) << 8)).view(np.float16).astype(np.float64)   # (n, 64)
        K = 16 * steps
        assert not wt[:, K:].any(), "weights beyond the issued K steps must be zero"
        c0 = L["chunk_c0"][chunk] - L["chunk_c0"][0]   # x_nhwc starts at the layer's first input channel
        a = np.zeros((H, W, K))
        ys, xs = slice(max(0, -dy), min(H, H - dy)), slice(max(0, -dx), min(W, W - dx))
        a[ys, xs] = x_nhwc[ys.start + dy:ys.stop + dy, xs.start + dx:xs.stop + dx, c0:c0 + K]   # TMA OOB fill = zero padding
        contrib = a @ wt[:, :K].T
        if first:
            acc[..., dcol:dcol + n] = contrib
        else:
            acc[..., dcol:dcol + n] += contrib
    assert not np.isnan(acc).any(), "every accumulator column must be initialised by an overwriting MMA"
    return acc


def _h(a):   # the value the fp16 weight blob holds: double -> float (plane table) -> half, as the packer does
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float16).astype(np.float64)


def _nchw(x_nhwc, c):
    return x_nhwc[..., :c].transpose(2, 0, 1)[None]


def test_rfdn_stage_distillation_fold_and_identity_tap():
    """B2.c2_r+d: cols 0..49 = c2_r(x) + x (identity tap), cols 64..88 = c2_d(x) in the centre tap (block.py:153-155)."""
    w = _weights(0)
    L = _layers("rfdn", w)["B2.c2_r+d"]
    assert (L["nchunks"], L["halo"], L["acc_cols"]) == (1, 1, 96) and len(L["entries"]) == 10   # centre tap (N = 96) + 8 taps + identity
    rng = np.random.default_rng(0)
    x = np.zeros((9, 13, 64))
    x[..., :50] = rng.standard_normal((9, 13, 50))
    acc = _replay(L, x)
    xin = _nchw(x, 50)
    r = O.conv2d(xin, _h(w["B2.c2_r.weight"]), None, 1, 1)[0].transpose(1, 2, 0) + x[..., :50]
    d = O.conv2d(xin, _h(w["B2.c2_d.weight"]), None)[0].transpose(1, 2, 0)
    assert np.abs(acc[..., :50] - r).max() < 1e-10 and not acc[..., 50:64].any()
    assert np.abs(acc[..., 64:89] - d).max() < 1e-10 and not acc[..., 89:96].any()
    np.testing.assert_allclose(L["bias"][0, :50], w["B2.c2_r.bias"], rtol=0, atol=0)
    np.testing.assert_allclose(L["bias"][1, :25], w["B2.c2_d.bias"], rtol=0, atol=0)
    g = L["groups"]
    assert list(g[:, 1]) == [64, 32] and list(g[:, 2]) == [1, 1]          # LeakyReLU on both groups


def test_rfdn_c5_with_the_esa_entry_folded_in():
    """B1.c5+...: c5 over the four 32-channel slots of `dist`, c1_ = conv1(c5(.)), cf' = conv4(conv_f(c1_)) + b4."""
    w = {k: v.astype(np.float64) for k, v in _weights(0).items()}
    L = _layers("rfdn", _weights(0))["B1.c5+esa.conv1+esa.conv_f+esa.conv4"]
    assert (L["nchunks"], L["halo"], L["acc_cols"]) == (2, 0, 144)
    rng = np.random.default_rng(1)
    dist = np.zeros((5, 7, 128))
    cat = rng.standard_normal((5, 7, 100))
    for s in range(4):
        dist[..., 32 * s:32 * s + 25] = cat[..., 25 * s:25 * s + 25]
    acc = _replay(L, dist) + np.concatenate([L["bias"][0], L["bias"][1, :16], L["bias"][2]])[None, None]
    xin = cat.transpose(2, 0, 1)[None]
    c5 = O.conv2d(xin, w["B1.c5.weight"], w["B1.c5.bias"])
    c1_ = O.conv2d(c5, w["B1.esa.conv1.weight"], w["B1.esa.conv1.bias"])
    cfp = O.conv2d(O.conv2d(c1_, w["B1.esa.conv_f.weight"], w["B1.esa.conv_f.bias"]), w["B1.esa.conv4.weight"], w["B1.esa.conv4.bias"])
    for got, ref in [(acc[..., :50], c5), (acc[..., 64:76], c1_), (acc[..., 80:130], cfp)]:
        ref = ref[0].transpose(1, 2, 0)
        # the composed matrices are rounded to fp16 once: relative 2^-11 per weight
        assert np.abs(got - ref).max() < 3e-3 * np.abs(ref).max()


def test_bsrn_stage_dense_bsconvu_border_class_bias():
    """B3.c1_r.pw+dw+d: BSConvU as one dense 3x3 (rank-1 taps, bias by border class), distillation Linear in the
    centre tap, identity tap for `+ x` (team18_bsrn.py:82-88,150-152)."""
    w = {k: v.astype(np.float64) for k, v in _weights(18).items()}
    L = _layers("bsrn", _weights(18))["B3.c1_r.pw+dw+d"]
    assert L["has_bias9"] and (L["nchunks"], L["halo"], L["acc_cols"]) == (1, 1, 96)
    rng = np.random.default_rng(2)
    H, W = 6, 8
    x = np.zeros((H, W, 64))
    x[..., :48] = rng.standard_normal((H, W, 48))
    acc = _replay(L, x)
    cls = np.array([[3 * (0 if y == 0 else (2 if y == H - 1 else 1)) + (0 if xx == 0 else (2 if xx == W - 1 else 1))
                     for xx in range(W)] for y in range(H)])
    got = acc[..., :48] + L["bias9"][cls][..., :48]
    # reference with the dense weights at their fp16 values; the bias path stays exact
    pw, pb = w["B3.c1_r.pw.weight"], w["B3.c1_r.pw.bias"]
    dw, db = w["B3.c1_r.dw.weight"][:, 0], w["B3.c1_r.dw.bias"]
    dense = _h(dw[:, None] * pw[:, :, None, None])
    ref = O.conv2d(_nchw(x, 48), dense, None, 1, 1)[0].transpose(1, 2, 0) + x[..., :48]
    ones = O.conv2d(np.ones((1, 48, H, W)), dw[:, None], None, 1, 1, groups=48)[0].transpose(1, 2, 0)   # sum of in-image taps
    ref = ref + db + pb * ones
    assert np.abs(got - ref).max() < 2e-6          # bias table is fp32
    # and the whole thing is the reference's BSConvU + x up to the fp16 rounding of the weights
    exact = O._bsconvu(w, "B3.c1_r.", _nchw(x, 48))[0].transpose(1, 2, 0) + x[..., :48]
    assert np.abs(got - exact).max() < 3e-3 * np.abs(exact).max()
    d = O.linear_nchw(_nchw(x, 48), _h(w["B3.c1_d.weight"]), np.zeros(24))[0].transpose(1, 2, 0)
    assert np.abs(acc[..., 64:88] - d).max() < 1e-10
    assert list(L["groups"][:, 2]) == [3, 3]       # GELU on both groups


@pytest.mark.parametrize("arch,mid,kw", [("rfdn", 0, {}), ("rlfn", 4, {}), ("imdn", -1, {}), ("bsrn", 18, {}),
                                         ("rfdn", 22, {"nf": 40}), ("rfdn_pruned", 40, {})])
def test_every_layer_fits_the_kernel_tables(arch, mid, kw):
    """Structural invariants of every packed layer: entry count, column coverage, K steps, blob extents."""
    for name, L in _layers(arch, _weights(mid), **kw).items():
        e = L["entries"]
        assert 1 <= len(e) <= 16 and L["acc_cols"] % 16 == 0 and L["acc_cols"] <= 256, name
        cover = np.zeros(L["acc_cols"], int)
        for dy, dx, chunk, steps, n, dcol, first, boff in e:
            assert abs(dy) <= L["halo"] and abs(dx) <= L["halo"] and 0 <= chunk < L["nchunks"], name
            assert 1 <= steps <= 4 and n % 16 == 0 and 16 <= n <= 256 and dcol + n <= L["acc_cols"], name
            assert boff % 1024 == 0 or boff % 128 == 0, name
            assert boff + n * 128 <= len(L["blob"]), name
            if first:
                cover[dcol:dcol + n] += 1
        assert (cover == 1).all(), (name, "every accumulator column is overwritten exactly once per tile")
        for col0, ncols, act, has_res, res_after, mode, slope_bits, b9 in L["groups"]:
            assert ncols % 16 == 0 and col0 + ncols <= L["acc_cols"], name


def test_rlfn_block_residual_inside_the_c5_gemm():
    """B2.c5+...: c5(u + x) as one GEMM over the two halves [x | u] of a 128-channel buffer with c5's weights on
    both (team04_rlfn.py:109-122: out = c3_r(...) + input; c5(out); esa)."""
    w = {k: v.astype(np.float64) for k, v in _weights(4).items()}
    L = _layers("rlfn", _weights(4))["B2.c5+esa.conv1+esa.conv_f+esa.conv4"]
    assert (L["nchunks"], L["halo"], L["acc_cols"]) == (2, 0, 112) and list(L["chunk_c0"][:2]) == [0, 64]
    rng = np.random.default_rng(3)
    buf = np.zeros((4, 6, 128))
    buf[..., :46] = rng.standard_normal((4, 6, 46))            # x
    buf[..., 64:110] = rng.standard_normal((4, 6, 46))         # u = lrelu(c3_r(...))
    acc = _replay(L, buf) + np.concatenate([L["bias"][0, :48], L["bias"][1, :16], L["bias"][2, :48]])[None, None]
    s = (buf[..., :46] + buf[..., 64:110]).transpose(2, 0, 1)[None]
    c5 = O.conv2d(s, w["B2.c5.weight"], w["B2.c5.bias"])
    c1_ = O.conv2d(c5, w["B2.esa.conv1.weight"], w["B2.esa.conv1.bias"])
    cfp = O.conv2d(O.conv2d(c1_, w["B2.esa.conv_f.weight"], w["B2.esa.conv_f.bias"]), w["B2.esa.conv4.weight"], w["B2.esa.conv4.bias"])
    for got, ref in [(acc[..., :46], c5), (acc[..., 48:64], c1_), (acc[..., 64:110], cfp)]:
        ref = ref[0].transpose(1, 2, 0)
        assert np.abs(got - ref).max() < 3e-3 * np.abs(ref).max()
    assert not L["groups"][:, 3].any()                           # no epilogue residual anywhere in this launch


def test_imdn_split_as_permuted_output_columns_and_pixel_shuffle_tail():
    """model.1.sub.0.conv2: out channels permuted to [remaining 48 | distilled 16] so the next conv reads a
    contiguous K = 48 (basicblock.py:259-265); model.2: the tail writes through the fused PixelShuffle (mode 1)."""
    w = _weights(-1)
    Ls = _layers("imdn", w)
    L = Ls["model.1.sub.0.conv2"]
    rng = np.random.default_rng(4)
    x = np.zeros((5, 6, 64))
    x[..., :48] = rng.standard_normal((5, 6, 48))
    acc = _replay(L, x)
    full = O.conv2d(_nchw(x, 48), _h(w["model.1.sub.0.conv2.0.weight"]), None, 1, 1)[0].transpose(1, 2, 0)   # (H, W, 64)
    assert np.abs(acc[..., :48] - full[..., 16:]).max() < 1e-10      # remaining channels first
    assert np.abs(acc[..., 48:64] - full[..., :16]).max() < 1e-10    # then the distilled 16
    b = w["model.1.sub.0.conv2.0.bias"]
    np.testing.assert_array_equal(np.concatenate([L["bias"][0, :48], L["bias"][1, :16]]), np.concatenate([b[16:], b[:16]]))
    tail = Ls["model.2"]
    assert list(tail["groups"][:, 5]) == [1] and tail["acc_cols"] == 48     # one group, pixel-shuffle store
