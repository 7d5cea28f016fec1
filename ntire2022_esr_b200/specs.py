"""Parameter-name / shape contracts of the four accelerated networks (what `load_state_dict(strict=True)`
accepts in the reference).  Derived from the module constructors:
  IMDN  models/imdn_baseline.py:32-61 + models/basicblock.py:230-257
  RFDN  models/rfdn_baseline/RFDN.py:11-26 + block.py:104-115,133-146
  RLFN  models/team04_rlfn.py:62-75,91-107,125-139
  BSRN  models/team18_bsrn.py:44-81,91-107,125-148,183-215
and the model registry of test_demo.py:13-30,52-58,150-157,203-209.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple


def _conv(d, name, o, i, k):
    d[name + ".weight"] = (o, i, k, k)
    d[name + ".bias"] = (o,)


def _lin(d, name, o, i):
    d[name + ".weight"] = (o, i)
    d[name + ".bias"] = (o,)


def imdn_spec(nc=64, nb=8) -> "OrderedDict[str, Tuple[int, ...]]":
    d = OrderedDict()
    dn = nc // 4
    rn = nc - dn
    _conv(d, "model.0", nc, 3, 3)
    for b in range(nb):
        p = f"model.1.sub.{b}."
        _conv(d, p + "conv1.0", nc, nc, 3)
        _conv(d, p + "conv2.0", nc, rn, 3)
        _conv(d, p + "conv3.0", nc, rn, 3)
        _conv(d, p + "conv4", dn, rn, 3)
        _conv(d, p + "conv1x1", nc, dn * 4, 1)
    _conv(d, f"model.1.sub.{nb}", nc, nc, 3)
    _conv(d, "model.2", 48, nc, 3)
    return d


def rfdn_spec(nf=50, nb=4, f=None):
    d = OrderedDict()
    dc = nf // 2
    f = f or nf // 4
    _conv(d, "fea_conv", nf, 3, 3)
    for b in range(1, nb + 1):
        p = f"B{b}."
        for s in (1, 2, 3):
            _conv(d, p + f"c{s}_d", dc, nf, 1)
            _conv(d, p + f"c{s}_r", nf, nf, 3)
        _conv(d, p + "c4", dc, nf, 3)
        _conv(d, p + "c5", nf, dc * 4, 1)
        e = p + "esa."
        _conv(d, e + "conv1", f, nf, 1)
        _conv(d, e + "conv_f", f, f, 1)
        _conv(d, e + "conv_max", f, f, 3)
        _conv(d, e + "conv2", f, f, 3)
        _conv(d, e + "conv3", f, f, 3)
        _conv(d, e + "conv3_", f, f, 3)
        _conv(d, e + "conv4", nf, f, 1)
    _conv(d, "c.0", nf, nf * nb, 1)
    _conv(d, "LR_conv", nf, nf, 3)
    _conv(d, "upsampler.0", 48, nf, 3)
    return d


def rlfn_spec(nf=46, nb=4, mf=48, f=16):
    d = OrderedDict()
    _conv(d, "fea_conv", nf, 3, 3)
    for b in range(1, nb + 1):
        p = f"B{b}."
        _conv(d, p + "c1_r", mf, nf, 3)
        _conv(d, p + "c2_r", mf, mf, 3)
        _conv(d, p + "c3_r", nf, mf, 3)
        _conv(d, p + "c5", nf, nf, 1)
        e = p + "esa."
        _conv(d, e + "conv1", f, nf, 1)
        _conv(d, e + "conv_f", f, f, 1)
        _conv(d, e + "conv2", f, f, 3)
        _conv(d, e + "conv3", f, f, 3)
        _conv(d, e + "conv4", nf, f, 1)
    _conv(d, "LR_conv", nf, nf, 3)
    _conv(d, "upsampler.0", 48, nf, 3)
    return d


def bsrn_spec(nf=48, nb=5):
    d = OrderedDict()
    dc, f = nf // 2, nf // 4

    def bsconv(name, o, i):
        _lin(d, name + ".pw", o, i)
        d[name + ".dw.weight"] = (o, 1, 3, 3)
        d[name + ".dw.bias"] = (o,)

    bsconv("fea_conv", nf, 12)
    for b in range(1, nb + 1):
        p = f"B{b}."
        d[p + "cw"] = (1, nf)
        for s in (1, 2, 3):
            _lin(d, p + f"c{s}_d", dc, nf)
            bsconv(p + f"c{s}_r", nf, nf)
        bsconv(p + "c4", dc, nf)
        _lin(d, p + "c5", nf, dc * 4)
        e = p + "esa."
        _lin(d, e + "conv1", f, nf)
        _lin(d, e + "conv_f", f, f)
        bsconv(e + "conv_max", f, f)
        _conv(d, e + "conv2", f, f, 3)
        bsconv(e + "conv3", f, f)
        bsconv(e + "conv3_", f, f)
        _lin(d, e + "conv4", nf, f)
        _lin(d, p + "conv_out", nf, nf)
    _lin(d, "c1", nf, nf * nb)
    bsconv("c2", nf, nf)
    _conv(d, "upsampler.upsampleOneStep.0", 48, nf, 3)
    return d


def rfdn_pruned_spec(nf=40, nb=4):
    """models/team40_rfdn_pruned.py:103-115,133-146: RFDN shapes with the ESA width fixed at 50 // 4 = 12."""
    return rfdn_spec(nf, nb, f=12)


def fmen_spec(nf=50, nb=4):
    """models/team03_fmen.py:83-118: head, warm-up conv + HFAB(two basic blocks, 12 channels), nb x [BasicBlock(nf),
    HFAB(one basic block, 16 channels)], lr_conv, tail conv in front of PixelShuffle(4).  Every convolution is 3x3."""
    d = OrderedDict()

    def hfab(p, up, mid):
        _conv(d, p + "squeeze", mid, nf, 3)
        for k in range(up):
            _conv(d, p + f"convs.{k}.conv1.rep_conv", mid, mid, 3)
            _conv(d, p + f"convs.{k}.conv2.rep_conv", mid, mid, 3)
        _conv(d, p + "excitate", nf, mid, 3)

    _conv(d, "head", nf, 3, 3)
    _conv(d, "warmup.0", nf, nf, 3)
    hfab("warmup.1.", 2, 12)
    for i in range(nb):
        _conv(d, f"basic_blocks.{i}.conv1.rep_conv", nf, nf, 3)
        _conv(d, f"basic_blocks.{i}.conv2.rep_conv", nf, nf, 3)
    for i in range(nb):
        hfab(f"hfabs.{i}.", 1, 16)
    _conv(d, "lr_conv", nf, nf, 3)
    _conv(d, "tail.0", 48, nf, 3)
    return d


SPECS = {"imdn": imdn_spec, "rfdn": rfdn_spec, "rlfn": rlfn_spec, "bsrn": bsrn_spec, "rfdn_pruned": rfdn_pruned_spec,
         "fmen": fmen_spec}

# model registry: id -> (arch, ctor kwargs, checkpoint file, state-dict wrapper key, name, data_range)
# (test_demo.py:17-23 IMDN, :24-30 RFDN, :45-51 FMEN, :52-58 RLFN, :150-157 BSRN, :175-181 RFDN40, :203-209 IMDN nb=7, :302-308 pruned RFDN)
REGISTRY: Dict[int, dict] = {
    -1: dict(arch="imdn", kwargs=dict(nf=64, nblocks=8), file="imdn_baseline.pth", wrap=None,
             name="IMDN_baseline", data_range=1.0),
    # FMEN (models/team03_fmen.py:78-134, test_demo.py:45-51)
    3: dict(arch="fmen", kwargs=dict(nf=50, nblocks=4), file="team03_fmen.pth", wrap=None,
            name="FMEN", data_range=255.0),
    0: dict(arch="rfdn", kwargs=dict(nf=50, nblocks=4), file="rfdn_baseline.pth", wrap=None,
            name="RFDN_baseline", data_range=255.0),
    4: dict(arch="rlfn", kwargs=dict(nf=46, nblocks=4), file="team04_rlfn.pth", wrap=None,
            name="RLFN", data_range=255.0),
    18: dict(arch="bsrn", kwargs=dict(nf=48, nblocks=5), file="team18_bsrn.pth", wrap="params",
             name="RFDNFINALB5", data_range=1.0),
    # RFDN40 (models/team22_rep_rfdn.py:134-165, test_demo.py:175-181): the RFDN graph at nf = 40, data range 1
    22: dict(arch="rfdn", kwargs=dict(nf=40, nblocks=4), file="team22_rep_rfdn.pth", wrap=None,
             name="RFDN40", data_range=1.0),
    # IMDN with seven blocks (test_demo.py:203-209)
    26: dict(arch="imdn", kwargs=dict(nf=64, nblocks=7), file="team26_imdn_nb7.pth", wrap=None,
             name="IMDN", data_range=1.0),
    # pruned RFDN (models/team40_rfdn_pruned.py:186-213, test_demo.py:302-308): nf = 40, RFDBs without the inner
    # residual adds, ESA width 12
    40: dict(arch="rfdn_pruned", kwargs=dict(nf=40, nblocks=4), file="team40_rfdn_pruned.pth", wrap=None,
             name="RFDNPrune", data_range=255.0),
}
