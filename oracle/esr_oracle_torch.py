"""CPU oracle, second restatement: the same four graphs as oracle/esr_oracle.py but expressed with
`torch.nn.functional` calls, i.e. the very ATen CPU kernels (oneDNN conv2d / linear, max_pool2d,
upsample_bilinear2d, pixel_shuffle) the reference's nn.Modules dispatch to.

TEST INFRASTRUCTURE ONLY (same rule as esr_oracle.py): imported by tests/, smoke() and bench.py's CPU
baseline / `--impl reference` arm, never by the product package.  It exists because the reference is
pure PyTorch and cannot travel to the GPU box: this file is what `bench.py --impl reference` times as
"the reference's own CPU implementation of the path" (all host threads, fp32, no_grad), and its
numerics are pinned against the same golden outputs of the unmodified reference (tests/test_oracle.py).

Reference graphs: models/rfdn_baseline/{RFDN.py:29-41, block.py:117-129,148-166},
models/imdn_baseline.py:46-65 + models/basicblock.py:259-265, models/team04_rlfn.py:76-152,
models/team18_bsrn.py:82-236.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _t(weights, dtype=torch.float32):
    return {k: torch.as_tensor(v).to(dtype) for k, v in weights.items()}


def _conv(w, name, x, stride=1, padding=0, groups=1):
    return F.conv2d(x, w[name + ".weight"], w[name + ".bias"], stride, padding, 1, groups)


def _lin(w, name, x):  # nn.Linear on the channel axis of NCHW (the reference permutes to NHWC)
    return F.linear(x.permute(0, 2, 3, 1), w[name + ".weight"], w[name + ".bias"]).permute(0, 3, 1, 2)


def _bsconv(w, p, x):
    t = _lin(w, p + "pw", x)
    return F.conv2d(t, w[p + "dw.weight"], w[p + "dw.bias"], 1, 1, 1, t.shape[1])


def _esa(w, p, x, kind):
    lin = kind == "bsrn"
    one = (lambda n, t: _lin(w, p + n, t)) if lin else (lambda n, t: _conv(w, p + n, t))
    c1_ = one("conv1", x)
    c1 = _conv(w, p + "conv2", c1_, stride=2, padding=0)
    v = F.max_pool2d(c1, kernel_size=7, stride=3)
    if kind == "rfdn":
        v = F.relu(_conv(w, p + "conv_max", v, padding=1))
        c3 = F.relu(_conv(w, p + "conv3", v, padding=1))
        c3 = _conv(w, p + "conv3_", c3, padding=1)
    elif kind == "rlfn":
        c3 = _conv(w, p + "conv3", v, padding=1)
    else:
        v = F.gelu(_bsconv(w, p + "conv_max.", v))
        c3 = F.gelu(_bsconv(w, p + "conv3.", v))
        c3 = _bsconv(w, p + "conv3_.", c3)
    c3 = F.interpolate(c3, (x.size(2), x.size(3)), mode="bilinear", align_corners=False)
    return x * torch.sigmoid(one("conv4", c3 + one("conv_f", c1_)))


def rfdn_forward(w, x, residual=True):
    """residual=False: models/team40_rfdn_pruned.py:148-166 (RFDB without the inner adds)."""
    a = lambda t: F.leaky_relu(t, 0.05)
    k = 1.0 if residual else 0.0
    fea = _conv(w, "fea_conv", x, padding=1)
    outs, t = [], fea
    nb = sum(1 for k in w if k.endswith(".c5.weight"))
    for b in range(1, nb + 1):
        p = f"B{b}."
        d1 = a(_conv(w, p + "c1_d", t)); r1 = a(_conv(w, p + "c1_r", t, padding=1) + k * t)
        d2 = a(_conv(w, p + "c2_d", r1)); r2 = a(_conv(w, p + "c2_r", r1, padding=1) + k * r1)
        d3 = a(_conv(w, p + "c3_d", r2)); r3 = a(_conv(w, p + "c3_r", r2, padding=1) + k * r2)
        r4 = a(_conv(w, p + "c4", r3, padding=1))
        t = _esa(w, p + "esa.", _conv(w, p + "c5", torch.cat([d1, d2, d3, r4], 1)), "rfdn")
        outs.append(t)
    out_b = a(_conv(w, "c.0", torch.cat(outs, 1)))
    out_lr = _conv(w, "LR_conv", out_b, padding=1) + fea
    return F.pixel_shuffle(_conv(w, "upsampler.0", out_lr, padding=1), 4)


def imdn_forward(w, x):
    a = lambda t: F.leaky_relu(t, 0.05)
    nb = sum(1 for k in w if k.endswith(".conv1x1.weight"))
    head = _conv(w, "model.0", x, padding=1)
    t = head
    for i in range(nb):
        p = f"model.1.sub.{i}."
        u = a(_conv(w, p + "conv1.0", t, padding=1)); d1, r = u[:, :16], u[:, 16:]
        u = a(_conv(w, p + "conv2.0", r, padding=1)); d2, r = u[:, :16], u[:, 16:]
        u = a(_conv(w, p + "conv3.0", r, padding=1)); d3, r = u[:, :16], u[:, 16:]
        d4 = _conv(w, p + "conv4", r, padding=1)
        t = t + _conv(w, p + "conv1x1", torch.cat([d1, d2, d3, d4], 1))
    t = head + _conv(w, f"model.1.sub.{nb}", t, padding=1)
    return F.pixel_shuffle(_conv(w, "model.2", t, padding=1), 4)


def rlfn_forward(w, x):
    a = lambda t: F.leaky_relu(t, 0.05)
    fea = _conv(w, "fea_conv", x, padding=1)
    t = fea
    nb = sum(1 for k in w if k.endswith(".c5.weight"))
    for b in range(1, nb + 1):
        p = f"B{b}."
        u = a(_conv(w, p + "c1_r", t, padding=1))
        u = a(_conv(w, p + "c2_r", u, padding=1))
        u = a(_conv(w, p + "c3_r", u, padding=1)) + t
        t = _esa(w, p + "esa.", _conv(w, p + "c5", u), "rlfn")
    out_lr = _conv(w, "LR_conv", t, padding=1) + fea
    return F.pixel_shuffle(_conv(w, "upsampler.0", out_lr, padding=1), 4)


def bsrn_forward(w, x):
    fea = _bsconv(w, "fea_conv.", torch.cat([x, x, x, x], 1))
    outs, t = [], fea
    nb = sum(1 for k in w if k.endswith(".cw"))
    for b in range(1, nb + 1):
        p = f"B{b}."
        d1 = F.gelu(_lin(w, p + "c1_d", t)); r1 = F.gelu(_bsconv(w, p + "c1_r.", t) + t)
        d2 = F.gelu(_lin(w, p + "c2_d", r1)); r2 = F.gelu(_bsconv(w, p + "c2_r.", r1) + r1)
        d3 = F.gelu(_lin(w, p + "c3_d", r2)); r3 = F.gelu(_bsconv(w, p + "c3_r.", r2) + r2)
        r4 = F.gelu(_bsconv(w, p + "c4.", r3))
        out = _lin(w, p + "c5", torch.cat([d1, d2, d3, r4], 1))
        fused = _esa(w, p + "esa.", out, "bsrn") * w[p + "cw"].reshape(1, -1, 1, 1)
        t = _lin(w, p + "conv_out", fused) + t
        outs.append(t)
    out_b = F.gelu(_lin(w, "c1", torch.cat(outs, 1)))
    out_lr = _bsconv(w, "c2.", out_b) + fea
    return F.pixel_shuffle(_conv(w, "upsampler.upsampleOneStep.0", out_lr, padding=1), 4)


def fmen_forward(w, x):
    """models/team03_fmen.py:37-42 (BasicBlock), :68-75 (HFAB), :121-134 (FMEN)"""
    a = lambda t: F.leaky_relu(t, 0.1)
    basic = lambda p, t: _conv(w, p + "conv2.rep_conv", a(_conv(w, p + "conv1.rep_conv", t, padding=1)), padding=1)

    def hfab(p, t):
        out = a(_conv(w, p + "squeeze", t, padding=1))
        k = 0
        while p + f"convs.{k}.conv1.rep_conv.weight" in w:
            out = basic(p + f"convs.{k}.", out)
            k += 1
        return torch.sigmoid(_conv(w, p + "excitate", a(out), padding=1)) * t

    x = _conv(w, "head", x, padding=1)
    h = hfab("warmup.1.", _conv(w, "warmup.0", x, padding=1))
    i = 0
    while f"basic_blocks.{i}.conv1.rep_conv.weight" in w:
        h = hfab(f"hfabs.{i}.", basic(f"basic_blocks.{i}.", h))
        i += 1
    return F.pixel_shuffle(_conv(w, "tail.0", _conv(w, "lr_conv", h, padding=1) + x, padding=1), 4)


FORWARD = {"imdn": imdn_forward, "rfdn": rfdn_forward, "rlfn": rlfn_forward, "bsrn": bsrn_forward,
           "rfdn_pruned": lambda w, x: rfdn_forward(w, x, residual=False), "fmen": fmen_forward}


def forward(arch, weights, x, dtype=torch.float32):
    """weights: name -> array/tensor (or an already converted dict); x: array/tensor NCHW."""
    w = weights if all(isinstance(v, torch.Tensor) and v.dtype == dtype for v in weights.values()) else _t(weights, dtype)
    with torch.no_grad():
        return FORWARD[arch](w, torch.as_tensor(x).to(dtype))


prepare = _t
