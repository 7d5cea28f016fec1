"""Reference-shaped module tree of the drop-in model, and the shape trace that keeps `utils/model_summary.py` working.

The reference harness reports #Params, #Conv2d, #Activations and FLOPs of the selected model
(test_demo.py:522-533) by registering forward hooks on its nn.Conv2d / nn.Linear / nn.ReLU / nn.LeakyReLU leaves
(utils/model_summary.py:230-245, 390-400) and running one forward.  The drop-in computes the forward inside the CUDA
engine, so (a) its parameters live in REAL leaf modules under the reference's names (`B1.esa.conv2` is an
nn.Conv2d(12, 12, 3, stride=2), `B1.c1_r.pw` an nn.Linear, ...; the parameter-free activation modules the reference
hooks exist too), and (b) after an engine forward, every leaf that carries forward hooks gets them called with
shape-only (`meta` device) tensors of exactly the shapes that layer sees in the reference's forward, in the
reference's call order.  The hooks only read shapes and module attributes, so the harness prints the reference's
numbers for the drop-in (tests/test_host_cpu.py pins them to values produced by the reference itself).

Reference graphs: models/rfdn_baseline/{RFDN.py:29-41, block.py:117-129,148-166}, models/imdn_baseline.py:46-65 +
models/basicblock.py:259-265, models/team04_rlfn.py:76-152, models/team18_bsrn.py:82-236,
models/team22_rep_rfdn.py:87-165, models/team40_rfdn_pruned.py:103-213, models/team03_fmen.py:10-134.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Tuple

import torch
import torch.nn as nn


class _Node(nn.Module):
    """Anonymous container so dotted reference names ('B1.esa.conv1.weight') map to module paths."""


def _leaf_for(name: str, shape: Tuple[int, ...]) -> nn.Module:
    """The nn leaf that owns `<name>.weight` of the given shape in the reference."""
    last = name.split(".")[-1]
    if len(shape) == 2:                                     # nn.Linear (BSRN's pointwise layers)
        return nn.Linear(shape[1], shape[0], bias=True)
    out_c, in_g, k, _ = shape
    groups = out_c if (last == "dw" and in_g == 1) else 1   # BSConvU depthwise 3x3
    stride, pad = (2, 0) if (last == "conv2" and ".esa." in "." + name) else (1, (k - 1) // 2)
    return nn.Conv2d(in_g * groups, out_c, k, stride, pad, groups=groups, bias=True)


def _act_modules(arch: str, nb: int) -> Dict[str, nn.Module]:
    """Parameter-free modules of the reference that model_summary hooks (nn.ReLU / nn.LeakyReLU)."""
    acts: Dict[str, nn.Module] = {}
    if arch in ("rfdn", "rfdn_pruned", "rlfn"):
        for b in range(1, nb + 1):
            acts[f"B{b}.act"] = nn.LeakyReLU(0.05, inplace=True)
            acts[f"B{b}.esa.relu"] = nn.ReLU(inplace=True)
        if arch != "rlfn":
            acts["c.1"] = nn.LeakyReLU(0.05, inplace=True)
    elif arch == "imdn":
        for b in range(nb):
            for c in (1, 2, 3):
                acts[f"model.1.sub.{b}.conv{c}.1"] = nn.LeakyReLU(0.05, inplace=True)
    elif arch == "fmen":
        # FMEN's LeakyReLU(0.1) is a module-level global, not a sub-module (team03_fmen.py:6-7): model_summary never sees
        # it.  The HFABs own an nn.Sigmoid (team03_fmen.py:66), which carries no hooks either but is part of .modules()
        acts["warmup.1.sigmoid"] = nn.Sigmoid()
        for b in range(nb):
            acts[f"hfabs.{b}.sigmoid"] = nn.Sigmoid()
    return acts


def build_tree(root: nn.Module, arch: str, nb: int, spec) -> None:
    """Registers, under `root`, one leaf per reference layer (weights zero-initialised, requires_grad False) plus the
    hooked activation modules, in the order of the reference's state dict."""
    def parent_of(path: str) -> Tuple[nn.Module, str]:
        parts = path.split(".")
        node = root
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _Node())
            node = node._modules[p]
        return node, parts[-1]

    done = set()
    for name, shape in spec.items():
        mod_path, pname = name.rsplit(".", 1)
        if pname not in ("weight", "bias"):                  # a bare nn.Parameter (BSRN's per-channel scale `cw`)
            node, leaf = parent_of(name)
            node.register_parameter(leaf, nn.Parameter(torch.zeros(shape), requires_grad=False))
            continue
        if mod_path in done:
            continue
        done.add(mod_path)
        leaf_mod = _leaf_for(mod_path, tuple(spec[mod_path + ".weight"]))
        for p in leaf_mod.parameters():
            p.requires_grad = False
            p.data.zero_()
        node, leaf = parent_of(mod_path)
        node.add_module(leaf, leaf_mod)
    for path, mod in _act_modules(arch, nb).items():
        node, leaf = parent_of(path)
        node.add_module(leaf, mod)


# ---------------------------------------------------------------------------------------------------------------------
# shape traces: (module path, input shape, output shape) for every call of a hooked leaf, in the reference's order
# ---------------------------------------------------------------------------------------------------------------------
Call = Tuple[str, Tuple[int, ...], Tuple[int, ...]]


def _esa_dims(H: int, W: int) -> Tuple[int, int, int, int]:
    H2, W2 = (H - 3) // 2 + 1, (W - 3) // 2 + 1             # conv2: 3x3 stride 2 pad 0
    return H2, W2, (H2 - 7) // 3 + 1, (W2 - 7) // 3 + 1     # max_pool2d(7, 3)


def _trace_rfdn(nf: int, nb: int, f: int, B: int, H: int, W: int, rlfn: bool = False, mf: int = 48) -> Iterator[Call]:
    H2, W2, H3, W3 = _esa_dims(H, W)
    full = lambda c: (B, c, H, W)
    dc = nf // 2
    yield "fea_conv", full(3), full(nf)
    for b in range(1, nb + 1):
        p = f"B{b}."
        if rlfn:                                             # team04_rlfn.py:109-122
            for name, ci, co in (("c1_r", nf, mf), ("c2_r", mf, mf), ("c3_r", mf, nf)):
                yield p + name, full(ci), full(co)
                yield p + "act", full(co), full(co)
            yield p + "c5", full(nf), full(nf)
        else:                                                # block.py:148-166
            for s in (1, 2, 3):
                yield p + f"c{s}_d", full(nf), full(dc)
                yield p + "act", full(dc), full(dc)
                yield p + f"c{s}_r", full(nf), full(nf)
                yield p + "act", full(nf), full(nf)
            yield p + "c4", full(nf), full(dc)
            yield p + "act", full(dc), full(dc)
            yield p + "c5", full(4 * dc), full(nf)
        e = p + "esa."                                       # block.py:117-129 / team04_rlfn.py:76-89
        yield e + "conv1", full(nf), full(f)
        yield e + "conv2", full(f), (B, f, H2, W2)
        low = (B, f, H3, W3)
        if not rlfn:
            yield e + "conv_max", low, low
            yield e + "relu", low, low
            yield e + "conv3", low, low
            yield e + "relu", low, low
            yield e + "conv3_", low, low
        else:
            yield e + "conv3", low, low
        yield e + "conv_f", full(f), full(f)
        yield e + "conv4", full(f), full(nf)
    if not rlfn:
        yield "c.0", full(nf * nb), full(nf)
        yield "c.1", full(nf), full(nf)
    yield "LR_conv", full(nf), full(nf)
    yield "upsampler.0", full(nf), full(48)


def _trace_imdn(nc: int, nb: int, B: int, H: int, W: int) -> Iterator[Call]:
    full = lambda c: (B, c, H, W)
    dn, rn = nc // 4, nc - nc // 4
    yield "model.0", full(3), full(nc)
    for b in range(nb):                                      # basicblock.py:259-265
        p = f"model.1.sub.{b}."
        cin = nc
        for c in (1, 2, 3):
            yield p + f"conv{c}.0", full(cin), full(nc)
            yield p + f"conv{c}.1", full(nc), full(nc)
            cin = rn
        yield p + "conv4", full(rn), full(dn)
        yield p + "conv1x1", full(4 * dn), full(nc)
    yield f"model.1.sub.{nb}", full(nc), full(nc)
    yield "model.2", full(nc), full(48)


def _trace_bsrn(nf: int, nb: int, f: int, B: int, H: int, W: int) -> Iterator[Call]:
    H2, W2, H3, W3 = _esa_dims(H, W)
    nchw = lambda c, h=H, w=W: (B, c, h, w)
    nhwc = lambda c, h=H, w=W: (B, h, w, c)                  # the Linears run on permuted (NHWC) tensors (team18_bsrn.py:83-88)

    def bsconv(p, ci, co, h=H, w=W):
        yield p + "pw", nhwc(ci, h, w), nhwc(co, h, w)
        yield p + "dw", nchw(co, h, w), nchw(co, h, w)

    dc = nf // 2
    yield from bsconv("fea_conv.", 12, nf)
    for b in range(1, nb + 1):                               # team18_bsrn.py:150-172
        p = f"B{b}."
        for s in (1, 2, 3):
            yield p + f"c{s}_d", nhwc(nf), nhwc(dc)
            yield from bsconv(p + f"c{s}_r.", nf, nf)
        yield from bsconv(p + "c4.", nf, dc)
        yield p + "c5", nhwc(4 * dc), nhwc(nf)
        e = p + "esa."                                       # team18_bsrn.py:109-122
        yield e + "conv1", nhwc(nf), nhwc(f)
        yield e + "conv2", nchw(f), nchw(f, H2, W2)
        for name in ("conv_max.", "conv3.", "conv3_."):
            yield from bsconv(e + name, f, f, H3, W3)
        yield e + "conv_f", nhwc(f), nhwc(f)
        yield e + "conv4", nhwc(f), nhwc(nf)
        yield p + "conv_out", nhwc(nf), nhwc(nf)
    yield "c1", nhwc(nf * nb), nhwc(nf)
    yield from bsconv("c2.", nf, nf)
    yield "upsampler.upsampleOneStep.0", nchw(nf), nchw(48)


def _trace_fmen(nf: int, nb: int, B: int, H: int, W: int) -> Iterator[Call]:
    full = lambda c: (B, c, H, W)

    def hfab(p, up, mid):                                    # team03_fmen.py:68-75
        yield p + "squeeze", full(nf), full(mid)
        for k in range(up):
            yield p + f"convs.{k}.conv1.rep_conv", full(mid), full(mid)
            yield p + f"convs.{k}.conv2.rep_conv", full(mid), full(mid)
        yield p + "excitate", full(mid), full(nf)

    yield "head", full(3), full(nf)                          # team03_fmen.py:121-134
    yield "warmup.0", full(nf), full(nf)
    yield from hfab("warmup.1.", 2, 12)
    for i in range(nb):
        yield f"basic_blocks.{i}.conv1.rep_conv", full(nf), full(nf)
        yield f"basic_blocks.{i}.conv2.rep_conv", full(nf), full(nf)
        yield from hfab(f"hfabs.{i}.", 1, 16)
    yield "lr_conv", full(nf), full(nf)
    yield "tail.0", full(nf), full(48)


def trace(arch: str, nf: int, nb: int, esa_f: int, B: int, H: int, W: int) -> List[Call]:
    if arch in ("rfdn", "rfdn_pruned"):
        return list(_trace_rfdn(nf, nb, esa_f, B, H, W))
    if arch == "rlfn":
        return list(_trace_rfdn(nf, nb, esa_f, B, H, W, rlfn=True))
    if arch == "imdn":
        return list(_trace_imdn(nf, nb, B, H, W))
    if arch == "bsrn":
        return list(_trace_bsrn(nf, nb, esa_f, B, H, W))
    if arch == "fmen":
        return list(_trace_fmen(nf, nb, B, H, W))
    raise NotImplementedError(arch)


def fire_forward_hooks(root: nn.Module, arch: str, nf: int, nb: int, esa_f: int, B: int, H: int, W: int) -> None:
    """Calls the forward hooks registered on the leaves (if any) with shape-only tensors, in the reference's order."""
    mods = dict(root.named_modules())
    if not any(m._forward_hooks for m in mods.values()):
        return
    for path, in_shape, out_shape in trace(arch, nf, nb, esa_f, B, H, W):
        m = mods.get(path)
        if m is None or not m._forward_hooks:
            continue
        x = torch.empty(in_shape, device="meta")
        y = torch.empty(out_shape, device="meta")
        for hook in list(m._forward_hooks.values()):
            hook(m, (x,), y)
