"""CPU oracle for the x4 efficient-SR hot path (IMDN / RFDN / RLFN / BSRN forward).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`ntire2022_esr_b200/`) may import this
module; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs do, and only as the checker or the timed CPU arm.

This is a numpy restatement of the reference's PyTorch graphs.  The reference itself contains no
arithmetic: every op is delegated to PyTorch ATen (third-party, un-pinned by the reference; the
container has torch 2.11.0+cu128).  The ATen ops used on the path are restated here from their
published definitions (cross-correlation conv2d with zero padding, max_pool2d floor mode,
upsample_bilinear2d align_corners=False, pixel_shuffle, leaky_relu, exact-erf GELU, sigmoid).

Parity pin: `tests/golden/*.npz` hold outputs of the UNMODIFIED reference modules
(`/root/reference/models/...` + `model_zoo/*.pth`, run through `test_demo.select_model/forward`)
produced by `tests/golden/make_golden.py`; `tests/test_oracle.py` checks this file against them.

All functions take NCHW arrays; `dtype` selects float32 (default, like the reference) or float64
(tie-breaker).  Reference citations are relative to /root/reference.
"""
from __future__ import annotations

import math
import numpy as np

try:  # exact erf for nn.GELU(approximate='none')
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


# ------------------------------------------------------------------------------------------------
# primitive ops (ATen semantics)
# ------------------------------------------------------------------------------------------------
def conv2d(x, w, b=None, stride=1, padding=0, groups=1):
    """torch.nn.functional.conv2d (cross-correlation, zero padding).  x (N,C,H,W), w (O,C/g,kh,kw)."""
    n, c, h, wd = x.shape
    o, cg, kh, kw = w.shape
    if padding:
        x = np.pad(x, ((0, 0), (0, 0), (padding, padding), (padding, padding)))
    ho = (h + 2 * padding - kh) // stride + 1
    wo = (wd + 2 * padding - kw) // stride + 1
    if groups == 1:
        assert cg == c
        # accumulate tap by tap: y[n,o,:,:] += w[o,:,i,j] . x[n,:,i::s,j::s]
        y = np.zeros((n, o, ho, wo), dtype=x.dtype)
        for i in range(kh):
            for j in range(kw):
                xs = x[:, :, i:i + (ho - 1) * stride + 1:stride, j:j + (wo - 1) * stride + 1:stride]
                y += np.einsum("oc,nchw->nohw", w[:, :, i, j], xs, optimize=True)
    else:
        assert groups == c == o and cg == 1, "only depthwise grouping is on the path"
        y = np.zeros((n, o, ho, wo), dtype=x.dtype)
        for i in range(kh):
            for j in range(kw):
                xs = x[:, :, i:i + (ho - 1) * stride + 1:stride, j:j + (wo - 1) * stride + 1:stride]
                y += w[None, :, 0, i, j, None, None] * xs
    if b is not None:
        y += b[None, :, None, None]
    return y


def linear_nchw(x, w, b):
    """nn.Linear applied on the channel axis of an NCHW tensor (the reference permutes to NHWC,
    models/team18_bsrn.py:82-88,110,150) == 1x1 conv with w (O,I)."""
    return np.einsum("oc,nchw->nohw", w, x, optimize=True) + b[None, :, None, None]


def leaky_relu(x, slope):
    return np.where(x >= 0, x, x * x.dtype.type(slope))


def relu(x):
    return np.maximum(x, 0)


def gelu(x):
    """nn.GELU() default approximate='none': 0.5*x*(1+erf(x/sqrt(2)))."""
    return (0.5 * x * (1.0 + _erf(x.astype(np.float64) / math.sqrt(2.0)))).astype(x.dtype)


def sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(x.dtype)


def max_pool2d(x, k, s):
    """F.max_pool2d(kernel_size=k, stride=s), padding 0, floor mode."""
    n, c, h, w = x.shape
    ho = (h - k) // s + 1
    wo = (w - k) // s + 1
    if ho < 1 or wo < 1:
        raise ValueError(f"max_pool2d: input {h}x{w} smaller than window {k}")
    y = np.full((n, c, ho, wo), -np.inf, dtype=x.dtype)
    for i in range(k):
        for j in range(k):
            y = np.maximum(y, x[:, :, i:i + (ho - 1) * s + 1:s, j:j + (wo - 1) * s + 1:s])
    return y


def _bilinear_axis(n_in, n_out, dtype):
    d = np.arange(n_out, dtype=np.float64)
    src = np.maximum((d + 0.5) * (n_in / n_out) - 0.5, 0.0)
    i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    lam = (src - i0).astype(dtype)
    return i0, i1, lam


def interpolate_bilinear(x, size):
    """F.interpolate(mode='bilinear', align_corners=False)."""
    h_out, w_out = size
    n, c, h, w = x.shape
    y0, y1, ly = _bilinear_axis(h, h_out, x.dtype)
    x0, x1, lx = _bilinear_axis(w, w_out, x.dtype)
    top = x[:, :, y0, :]
    bot = x[:, :, y1, :]
    rows = top * (1 - ly)[None, None, :, None] + bot * ly[None, None, :, None]
    left = rows[:, :, :, x0]
    right = rows[:, :, :, x1]
    return left * (1 - lx)[None, None, None, :] + right * lx[None, None, None, :]


def pixel_shuffle(x, r):
    """nn.PixelShuffle(r): out[b,c,r*h+i,r*w+j] = in[b, c*r*r + i*r + j, h, w]."""
    n, c, h, w = x.shape
    co = c // (r * r)
    return x.reshape(n, co, r, r, h, w).transpose(0, 1, 4, 2, 5, 3).reshape(n, co, h * r, w * r)


def _cast(weights, dtype):
    return {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}


def _conv(wt, name, x, stride=1, padding=0, groups=1):
    return conv2d(x, wt[name + ".weight"], wt[name + ".bias"], stride, padding, groups)


# ------------------------------------------------------------------------------------------------
# RFDN  (models/rfdn_baseline/RFDN.py:29-41, block.py:117-129,148-166)
# ------------------------------------------------------------------------------------------------
def _esa_rfdn(wt, p, x):
    """ESA.forward, models/rfdn_baseline/block.py:117-129."""
    c1_ = _conv(wt, p + "conv1", x)
    c1 = _conv(wt, p + "conv2", c1_, stride=2, padding=0)
    v_max = max_pool2d(c1, 7, 3)
    v_range = relu(_conv(wt, p + "conv_max", v_max, padding=1))
    c3 = relu(_conv(wt, p + "conv3", v_range, padding=1))
    c3 = _conv(wt, p + "conv3_", c3, padding=1)
    c3 = interpolate_bilinear(c3, x.shape[2:])
    cf = _conv(wt, p + "conv_f", c1_)
    c4 = _conv(wt, p + "conv4", c3 + cf)
    return x * sigmoid(c4)


def _rfdb(wt, p, x, slope=0.05, residual=True):
    """RFDB.forward, models/rfdn_baseline/block.py:148-166 (residual add BEFORE the activation);
    residual=False: the pruned variant without the inner adds, models/team40_rfdn_pruned.py:148-166."""
    k = 1.0 if residual else 0.0
    d1 = leaky_relu(_conv(wt, p + "c1_d", x), slope)
    r1 = leaky_relu(_conv(wt, p + "c1_r", x, padding=1) + k * x, slope)
    d2 = leaky_relu(_conv(wt, p + "c2_d", r1), slope)
    r2 = leaky_relu(_conv(wt, p + "c2_r", r1, padding=1) + k * r1, slope)
    d3 = leaky_relu(_conv(wt, p + "c3_d", r2), slope)
    r3 = leaky_relu(_conv(wt, p + "c3_r", r2, padding=1) + k * r2, slope)
    r4 = leaky_relu(_conv(wt, p + "c4", r3, padding=1), slope)
    out = np.concatenate([d1, d2, d3, r4], axis=1)
    return _esa_rfdn(wt, p + "esa.", _conv(wt, p + "c5", out))


def rfdn_forward(weights, x, dtype=np.float32, return_intermediates=False, residual=True):
    """RFDN.forward, models/rfdn_baseline/RFDN.py:29-41 (channel counts come from the weights)."""
    wt = _cast(weights, dtype)
    x = np.asarray(x, dtype=dtype)
    fea = _conv(wt, "fea_conv", x, padding=1)
    b1 = _rfdb(wt, "B1.", fea, residual=residual)
    b2 = _rfdb(wt, "B2.", b1, residual=residual)
    b3 = _rfdb(wt, "B3.", b2, residual=residual)
    b4 = _rfdb(wt, "B4.", b3, residual=residual)
    out_b = leaky_relu(_conv(wt, "c.0", np.concatenate([b1, b2, b3, b4], axis=1)), 0.05)
    out_lr = _conv(wt, "LR_conv", out_b, padding=1) + fea
    y = pixel_shuffle(_conv(wt, "upsampler.0", out_lr, padding=1), 4)
    if return_intermediates:
        return y, dict(fea=fea, b1=b1, b2=b2, b3=b3, b4=b4, out_b=out_b, out_lr=out_lr)
    return y


def rfdn_pruned_forward(weights, x, dtype=np.float32):
    """RFDN.forward of models/team40_rfdn_pruned.py:186-213: the RFDN graph with RFDBs that drop the inner
    residual adds (:148-166) and an ESA of fixed width 12 (:106); the widths come from the weights."""
    return rfdn_forward(weights, x, dtype=dtype, residual=False)


# ------------------------------------------------------------------------------------------------
# IMDN  (models/imdn_baseline.py:32-65, models/basicblock.py:191-205,230-265,446-449)
# ------------------------------------------------------------------------------------------------
def _imdb(wt, p, x, slope=0.05, d_nc=16):
    """IMDBlock.forward, models/basicblock.py:259-265: the split is taken AFTER the LeakyReLU."""
    t = leaky_relu(_conv(wt, p + "conv1.0", x, padding=1), slope)
    d1, r1 = t[:, :d_nc], t[:, d_nc:]
    t = leaky_relu(_conv(wt, p + "conv2.0", r1, padding=1), slope)
    d2, r2 = t[:, :d_nc], t[:, d_nc:]
    t = leaky_relu(_conv(wt, p + "conv3.0", r2, padding=1), slope)
    d3, r3 = t[:, :d_nc], t[:, d_nc:]
    d4 = _conv(wt, p + "conv4", r3, padding=1)
    res = _conv(wt, p + "conv1x1", np.concatenate([d1, d2, d3, d4], axis=1))
    return x + res


def imdn_forward(weights, x, dtype=np.float32, nb=None):
    """IMDN.forward, models/imdn_baseline.py:63-65 with the Sequential built at :46-61."""
    wt = _cast(weights, dtype)
    x = np.asarray(x, dtype=dtype)
    if nb is None:
        nb = sum(1 for k in wt if k.startswith("model.1.sub.") and k.endswith(".conv1x1.weight"))
    head = _conv(wt, "model.0", x, padding=1)
    t = head
    for i in range(nb):
        t = _imdb(wt, f"model.1.sub.{i}.", t)
    t = _conv(wt, f"model.1.sub.{nb}", t, padding=1)
    t = head + t                                    # ShortcutBlock, basicblock.py:197-199
    return pixel_shuffle(_conv(wt, "model.2", t, padding=1), 4)


# ------------------------------------------------------------------------------------------------
# RLFN  (models/team04_rlfn.py:76-89,109-122,141-152)
# ------------------------------------------------------------------------------------------------
def _esa_rlfn(wt, p, x):
    """slim ESA, models/team04_rlfn.py:76-89: only conv3 after the pool, no ReLU."""
    c1_ = _conv(wt, p + "conv1", x)
    c1 = _conv(wt, p + "conv2", c1_, stride=2, padding=0)
    v_max = max_pool2d(c1, 7, 3)
    c3 = _conv(wt, p + "conv3", v_max, padding=1)
    c3 = interpolate_bilinear(c3, x.shape[2:])
    cf = _conv(wt, p + "conv_f", c1_)
    c4 = _conv(wt, p + "conv4", c3 + cf)
    return x * sigmoid(c4)


def _rlfb(wt, p, x, slope=0.05):
    """RLFB.forward, models/team04_rlfn.py:109-122."""
    t = leaky_relu(_conv(wt, p + "c1_r", x, padding=1), slope)
    t = leaky_relu(_conv(wt, p + "c2_r", t, padding=1), slope)
    t = leaky_relu(_conv(wt, p + "c3_r", t, padding=1), slope)
    t = t + x
    return _esa_rlfn(wt, p + "esa.", _conv(wt, p + "c5", t))


def rlfn_forward(weights, x, dtype=np.float32):
    """RLFN_cut.forward, models/team04_rlfn.py:141-152."""
    wt = _cast(weights, dtype)
    x = np.asarray(x, dtype=dtype)
    fea = _conv(wt, "fea_conv", x, padding=1)
    t = fea
    for b in ("B1.", "B2.", "B3.", "B4."):
        t = _rlfb(wt, b, t)
    out_lr = _conv(wt, "LR_conv", t, padding=1) + fea
    return pixel_shuffle(_conv(wt, "upsampler.0", out_lr, padding=1), 4)


# ------------------------------------------------------------------------------------------------
# BSRN  (models/team18_bsrn.py:82-88,109-122,150-172,217-236)
# ------------------------------------------------------------------------------------------------
def _bsconvu(wt, p, x):
    """BSConvU.forward, team18_bsrn.py:82-88: pointwise Linear THEN depthwise 3x3 (zero pad 1)."""
    t = linear_nchw(x, wt[p + "pw.weight"], wt[p + "pw.bias"])
    return conv2d(t, wt[p + "dw.weight"], wt[p + "dw.bias"], 1, 1, groups=t.shape[1])


def _lin(wt, name, x):
    return linear_nchw(x, wt[name + ".weight"], wt[name + ".bias"])


def _esa_bsrn(wt, p, x):
    """ESA.forward, team18_bsrn.py:109-122."""
    c1_ = _lin(wt, p + "conv1", x)
    c1 = _conv(wt, p + "conv2", c1_, stride=2, padding=0)
    v_max = max_pool2d(c1, 7, 3)
    v_range = gelu(_bsconvu(wt, p + "conv_max.", v_max))
    c3 = gelu(_bsconvu(wt, p + "conv3.", v_range))
    c3 = _bsconvu(wt, p + "conv3_.", c3)
    c3 = interpolate_bilinear(c3, x.shape[2:])
    cf = _lin(wt, p + "conv_f", c1_)
    c4 = _lin(wt, p + "conv4", c3 + cf)
    return x * sigmoid(c4)


def _rfdb_bsrn(wt, p, x):
    """RFDB.forward, team18_bsrn.py:150-172."""
    d1 = gelu(_lin(wt, p + "c1_d", x))
    r1 = gelu(_bsconvu(wt, p + "c1_r.", x) + x)
    d2 = gelu(_lin(wt, p + "c2_d", r1))
    r2 = gelu(_bsconvu(wt, p + "c2_r.", r1) + r1)
    d3 = gelu(_lin(wt, p + "c3_d", r2))
    r3 = gelu(_bsconvu(wt, p + "c3_r.", r2) + r2)
    r4 = gelu(_bsconvu(wt, p + "c4.", r3))
    out = _lin(wt, p + "c5", np.concatenate([d1, d2, d3, r4], axis=1))
    fused = _esa_bsrn(wt, p + "esa.", out) * wt[p + "cw"].reshape(1, -1, 1, 1)
    return _lin(wt, p + "conv_out", fused) + x


def bsrn_forward(weights, x, dtype=np.float32):
    """BSRN.forward, team18_bsrn.py:217-236 (num_feat=48, num_block=5, test_demo.py:155-156)."""
    wt = _cast(weights, dtype)
    x = np.asarray(x, dtype=dtype)
    x4 = np.concatenate([x, x, x, x], axis=1)
    fea = _bsconvu(wt, "fea_conv.", x4)
    outs = []
    t = fea
    nblk = sum(1 for k in wt if k.endswith(".cw"))
    for i in range(1, nblk + 1):
        t = _rfdb_bsrn(wt, f"B{i}.", t)
        outs.append(t)
    out_b = gelu(_lin(wt, "c1", np.concatenate(outs, axis=1)))
    out_lr = _bsconvu(wt, "c2.", out_b) + fea
    return pixel_shuffle(_conv(wt, "upsampler.upsampleOneStep.0", out_lr, padding=1), 4)


# ------------------------------------------------------------------------------------------------
# FMEN  (models/team03_fmen.py:10-134; model id 3, test_demo.py:45-51): 3x3 convolutions only, LeakyReLU(0.1),
# high-frequency attention blocks that gate their input with a sigmoid
# ------------------------------------------------------------------------------------------------
def _fmen_basic(wt, p, x):
    """BasicBlock.forward, team03_fmen.py:37-42: RepConv - LeakyReLU(0.1) - RepConv"""
    return _conv(wt, p + "conv2.rep_conv", leaky_relu(_conv(wt, p + "conv1.rep_conv", x, padding=1), 0.1), padding=1)


def _fmen_hfab(wt, p, x):
    """HFAB.forward, team03_fmen.py:68-75"""
    out = leaky_relu(_conv(wt, p + "squeeze", x, padding=1), 0.1)
    k = 0
    while p + f"convs.{k}.conv1.rep_conv.weight" in wt:
        out = _fmen_basic(wt, p + f"convs.{k}.", out)
        k += 1
    out = leaky_relu(out, 0.1)
    return sigmoid(_conv(wt, p + "excitate", out, padding=1)) * x


def fmen_forward(weights, x, dtype=np.float32):
    """FMEN.forward, team03_fmen.py:121-134 (n_feats 50, four down blocks, warm-up HFAB with two basic blocks)."""
    wt = _cast(weights, dtype)
    x = _conv(wt, "head", np.asarray(x, dtype=dtype), padding=1)
    h = _fmen_hfab(wt, "warmup.1.", _conv(wt, "warmup.0", x, padding=1))
    i = 0
    while f"basic_blocks.{i}.conv1.rep_conv.weight" in wt:
        h = _fmen_hfab(wt, f"hfabs.{i}.", _fmen_basic(wt, f"basic_blocks.{i}.", h))
        i += 1
    h = _conv(wt, "lr_conv", h, padding=1) + x
    return pixel_shuffle(_conv(wt, "tail.0", h, padding=1), 4)


# ------------------------------------------------------------------------------------------------
# registry mirroring test_demo.select_model (test_demo.py:13-30,52-58,150-157)
# ------------------------------------------------------------------------------------------------
MODELS = {
    # id 3: FMEN (test_demo.py:45-51)
    3: dict(arch="fmen", name="03_FMEN", data_range=255.0, weights="team03_fmen", fn=fmen_forward),
    -1: dict(arch="imdn", name="-1_IMDN_baseline", data_range=1.0, weights="imdn_baseline", fn=imdn_forward),
    0: dict(arch="rfdn", name="00_RFDN_baseline", data_range=255.0, weights="rfdn_baseline", fn=rfdn_forward),
    4: dict(arch="rlfn", name="04_RLFN", data_range=255.0, weights="team04_rlfn", fn=rlfn_forward),
    18: dict(arch="bsrn", name="18_RFDNFINALB5", data_range=1.0, weights="team18_bsrn", fn=bsrn_forward),
    # id 22: the same RFDN graph at nf = 40 (models/team22_rep_rfdn.py:101-165, test_demo.py:175-181)
    22: dict(arch="rfdn", name="22_RFDN40", data_range=1.0, weights="team22_rep_rfdn", fn=rfdn_forward),
    # id 26: IMDN with seven blocks (test_demo.py:203-209); imdn_forward takes the block count from the state dict
    26: dict(arch="imdn", name="26_IMDN", data_range=1.0, weights="team26_imdn_nb7", fn=imdn_forward),
    # id 40: pruned RFDN, nf = 40, no inner residuals, ESA width 12 (test_demo.py:302-308)
    40: dict(arch="rfdn_pruned", name="40_RFDNPrune", data_range=255.0, weights="team40_rfdn_pruned", fn=rfdn_pruned_forward),
}
FORWARD = {"imdn": imdn_forward, "rfdn": rfdn_forward, "rlfn": rlfn_forward, "bsrn": bsrn_forward,
           "rfdn_pruned": rfdn_pruned_forward, "fmen": fmen_forward}


def forward(arch, weights, x, dtype=np.float32):
    return FORWARD[arch](weights, x, dtype=dtype)


def forward_tiled(arch, weights, x, tile, tile_overlap=32, scale=4, dtype=np.float32):
    """test_demo.forward tiled branch, test_demo.py:368-389 (E/W accumulate and divide)."""
    x = np.asarray(x, dtype=dtype)
    b, c, h, w = x.shape
    tile = min(tile, h, w)
    stride = tile - tile_overlap
    hs = list(range(0, h - tile, stride)) + [h - tile]
    ws = list(range(0, w - tile, stride)) + [w - tile]
    e = np.zeros((b, c, h * scale, w * scale), dtype=dtype)
    wsum = np.zeros_like(e)
    for hi in hs:
        for wi in ws:
            o = forward(arch, weights, x[..., hi:hi + tile, wi:wi + tile], dtype=dtype)
            e[..., hi * scale:(hi + tile) * scale, wi * scale:(wi + tile) * scale] += o
            wsum[..., hi * scale:(hi + tile) * scale, wi * scale:(wi + tile) * scale] += 1
    return e / wsum


def tensor2uint(y, data_range):
    """utils/utils_image.py:204-208 for a (1,3,H,W) array -> HWC uint8."""
    img = np.clip(np.asarray(y, dtype=np.float32)[0], 0, data_range).transpose(1, 2, 0)
    return np.uint8(np.round(img * np.float32(255.0) / np.float32(data_range)))


def uint2tensor4(img, data_range):
    """utils/utils_image.py:190-193: HWC uint8 -> (1,3,H,W) float32 scaled to [0,data_range]."""
    return (img.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0 / data_range))[None]


def psnr(a, b, border=0, peak=255.0):
    """utils/utils_image.py:490-503."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if border:
        a = a[..., border:-border, border:-border] if a.ndim == 4 else a[border:-border, border:-border]
        b = b[..., border:-border, border:-border] if b.ndim == 4 else b[border:-border, border:-border]
    mse = np.mean((a - b) ** 2)
    return float("inf") if mse == 0 else 20 * math.log10(peak / math.sqrt(mse))


# ---------------------------------------------------------------------------------------------
# MATLAB-style bicubic resize (the DIV2K LR protocol), utils/utils_image.py:565-626 (cubic,
# calculate_weights_indices) and :704-774 (imresize_np).  Restated per axis as a dense weight matrix:
# output k sits at u = k / scale + 0.5 (1 - 1 / scale) (1-based), takes ceil(4 / scale) + 2 taps starting at
# floor(u - 2 / scale) with the antialiasing kernel scale * cubic(scale * d), rows normalised to 1, and taps that
# fall outside the image are mirrored about the border (the reference's symmetric padding).
# ---------------------------------------------------------------------------------------------
def _cubic(d):
    a = np.abs(d)
    return np.where(a <= 1, 1.5 * a ** 3 - 2.5 * a ** 2 + 1, np.where(a <= 2, -0.5 * a ** 3 + 2.5 * a ** 2 - 4 * a + 2, 0.0))


def _resize_matrix(n_in, scale, antialiasing=True):
    n_out = int(np.ceil(n_in * scale))
    kw = 4.0 / scale if (scale < 1 and antialiasing) else 4.0
    k = np.arange(1, n_out + 1, dtype=np.float64)
    u = k / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(u - kw / 2)
    P = int(np.ceil(kw)) + 2
    idx = left[:, None] + np.arange(P)[None, :]                      # 1-based input positions
    d = u[:, None] - idx
    wgt = scale * _cubic(d * scale) if (scale < 1 and antialiasing) else _cubic(d)
    wgt = wgt / wgt.sum(axis=1, keepdims=True)
    idx = idx.astype(np.int64)
    idx = np.where(idx < 1, 1 - idx, idx)                            # mirror (edge pixel repeated)
    idx = np.where(idx > n_in, 2 * n_in + 1 - idx, idx)
    M = np.zeros((n_out, n_in), dtype=np.float64)
    np.add.at(M, (np.repeat(np.arange(n_out), P), (idx - 1).ravel()), wgt.ravel())
    return M


def imresize_np(img, scale, antialiasing=True):
    """img: (H, W, C) or (H, W) float array in [0, 1]; returns float32 like the reference (no rounding)."""
    a = np.asarray(img, dtype=np.float64)
    squeeze = a.ndim == 2
    if squeeze:
        a = a[:, :, None]
    Mh, Mw = _resize_matrix(a.shape[0], scale, antialiasing), _resize_matrix(a.shape[1], scale, antialiasing)
    out = np.einsum("ih,hwc->iwc", Mh, a)
    out = np.einsum("jw,iwc->ijc", Mw, out)
    out = out.astype(np.float32)
    return out[:, :, 0] if squeeze else out


def shaped_input(seed, h, w, data_range):
    """Seeded (1,3,h,w) input of the BASELINE-shape goldens (tests/golden/make_golden.py::shaped_input): uniform in
    [0, data_range), rounded to fp16-representable values so the fp16 engine and the fp32 reference see the same tensor."""
    rng = np.random.default_rng(int(seed))
    return (rng.random((1, 3, h, w), dtype=np.float32) * np.float32(data_range)).astype(np.float16).astype(np.float32)


def load_weights(path):
    """Load an .npz written by tests/golden/make_golden.py (state-dict names -> fp32 arrays)."""
    with np.load(path) as z:
        return {k: z[k] for k in z.files}
