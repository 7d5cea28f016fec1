"""Drop-in for the reference's inference seam: `select_model(args, device)` and
`forward(img_lq, model, tile=None, tile_overlap=32, scale=4)` (test_demo.py:13-341, 364-391).

Same names, argument meaning, return values and error behaviour; the returned model is an
`nn.Module` whose parameters carry the reference's state-dict names (so `load_state_dict(strict=True)`,
`.eval()`, `.to()`, `.parameters()`, `.modules()` keep working for the reference's callers) and whose
`forward` runs the hand-written sm_100a engine through the C ABI.  There is no PyTorch compute path:
calling the model on a CPU tensor raises.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn

from . import module_tree, specs
from .engine import Engine, EsrError

MODEL_ZOO_ENV = "ESR_MODEL_ZOO"


class B200SRModel(nn.Module):
    """nn.Module facade over an esr_b200 `Engine`.

    arch in {'imdn','rfdn','rlfn','bsrn','rfdn_pruned','fmen'}; nf / nblocks follow the reference constructors
    (IMDN(nc=64, nb=8), RFDN(nf=50, 4 blocks), RLFN_cut(46), BSRN(num_feat=48, num_block=5), pruned RFDN(nf=40)).
    """

    def __init__(self, arch: str, nf: int = 0, nblocks: int = 0):
        super().__init__()
        if arch not in specs.SPECS:
            raise NotImplementedError(f"architecture {arch!r} is not implemented")
        self.arch = arch
        defaults = {"imdn": (64, 8), "rfdn": (50, 4), "rlfn": (46, 4), "bsrn": (48, 5), "rfdn_pruned": (40, 4), "fmen": (50, 4)}[arch]
        self.nf = nf or defaults[0]
        self.nblocks = nblocks or defaults[1]
        self._spec = specs.SPECS[arch](self.nf, self.nblocks)
        # real nn.Conv2d / nn.Linear leaves under the reference's names (plus the activation modules model_summary hooks):
        # they hold the weights and are never called - the forward runs in the CUDA engine
        module_tree.build_tree(self, arch, self.nblocks, self._spec)
        self._esa_f = self._spec.get("B1.esa.conv1.weight", (0,))[0]
        self._engine: Optional[Engine] = None
        self._engine_key = None
        self.engine_options = {}

    # any change of device / dtype / values invalidates the packed copy inside the engine
    def _apply(self, fn, *a, **k):
        self._drop_engine()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._drop_engine()
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _drop_engine(self):
        if getattr(self, "_engine", None) is not None:
            self._engine.close()
        self._engine = None
        self._engine_key = None

    def set_engine_option(self, key: str, value: int):
        self.engine_options[key] = int(value)
        if self._engine is not None:
            self._engine.set_option(key, value)

    def engine(self, device: torch.device) -> Engine:
        key = (device.type, device.index)
        if self._engine is None or self._engine_key != key:
            self._drop_engine()
            if device.type != "cuda":
                raise EsrError(-5, "B200SRModel runs only on a CUDA (sm_100) device; there is no CPU fallback")
            idx = device.index if device.index is not None else torch.cuda.current_device()
            eng = Engine(self.arch, idx, self.nf, self.nblocks)
            for k, v in self.engine_options.items():
                eng.set_option(k, v)
            eng.load_state_dict({k: v for k, v in self.state_dict().items()})
            self._engine, self._engine_key = eng, key
        return self._engine

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.engine(x.device).forward(x)
        # utils/model_summary.py (test_demo.py:522-530) counts FLOPs / activations through forward hooks on the leaves:
        # serve them with shape-only tensors of the reference's forward (no-op when nothing is hooked)
        module_tree.fire_forward_hooks(self, self.arch, self.nf, self.nblocks, self._esa_f, *x.shape[:1], *x.shape[2:])
        return y


def forward_uint8(img_lr, model: B200SRModel, data_range: float, half: bool = True):
    """The per-image body of the reference's run() (test_demo.py:422-434) in one device call:
    `util.tensor2uint(forward(util.uint2tensor4(img_lr, data_range).to(device), model), data_range)`.

    img_lr: uint8 (H,W,3) or (B,H,W,3), a numpy array (host path: copies in, runs, copies out) or a CUDA tensor.
    Returns uint8 (4H,4W,3) / (B,4H,4W,3) of the same kind."""
    if isinstance(img_lr, torch.Tensor):
        return model.engine(img_lr.device).forward_uint8(img_lr, data_range, half=half)
    dev = next(model.parameters()).device
    return model.engine(dev).forward_host_uint8(img_lr, data_range, half=half)


def _find_checkpoint(fname: str) -> str:
    # the reference resolves 'model_zoo/<file>' against the CWD (test_demo.py:21,28,56,154)
    cands = [os.path.join("model_zoo", fname)]
    if os.environ.get(MODEL_ZOO_ENV):
        cands.append(os.path.join(os.environ[MODEL_ZOO_ENV], fname))
    for c in cands:
        if os.path.exists(c):
            return c
    raise FileNotFoundError(f"{cands[0]} not found (set {MODEL_ZOO_ENV} to the reference's model_zoo directory)")


def build_model(model_id: int, state_dict=None) -> B200SRModel:
    """Construct + load; `state_dict` (name -> array/tensor) overrides the model_zoo lookup."""
    if model_id in specs.REGISTRY:
        reg = specs.REGISTRY[model_id]
    else:
        raise NotImplementedError(f"Model {model_id} is not implemented.")
    model = B200SRModel(reg["arch"], **reg["kwargs"])
    if state_dict is None:
        sd = torch.load(_find_checkpoint(reg["file"]), map_location="cpu")
        if reg["wrap"]:
            sd = sd[reg["wrap"]]
    else:
        sd = {k: torch.as_tensor(v) for k, v in state_dict.items()}
    model.load_state_dict(sd, strict=True)
    return model


def select_model(args, device):
    """test_demo.select_model for the accelerated ids (-1 IMDN, 0 RFDN, 3 FMEN, 4 RLFN, 18 BSRN, 22 RFDN40, 26 IMDN nb=7,
    40 pruned RFDN)."""
    model_id = args.model_id
    if model_id in specs.REGISTRY:
        name, data_range = f"{model_id:02}_{specs.REGISTRY[model_id]['name']}", specs.REGISTRY[model_id]["data_range"]
    else:
        raise NotImplementedError(f"Model {model_id} is not implemented.")
    model = build_model(model_id)
    model.eval()
    tile = None   # (the reference returns tile = 256 for model id 2 only, test_demo.py:337; id 2 is not on the accelerated path)
    for _, v in model.named_parameters():
        v.requires_grad = False
    model = model.to(device)
    return model, name, data_range, tile


TILE_BATCH = 64   # tiles per engine call in the tiled forward


def forward(img_lq, model, tile=None, tile_overlap=32, scale=4):
    """test_demo.forward: whole image, or overlapping tiles accumulated in E and normalised by the
    per-pixel coverage count W (test_demo.py:368-389).

    The tiles of the reference's double loop are independent and equally sized, so they go through the engine
    as batches of up to TILE_BATCH (every image of a batch is bit-identical to its single-image run); they are
    then accumulated in the reference's loop order, which keeps the fp32 sums bit-identical to the
    tile-by-tile evaluation."""
    if tile is None:
        return model(img_lq)
    b, c, h, w = img_lq.size()
    tile = min(tile, h, w)
    stride = tile - tile_overlap
    ys = list(range(0, h - tile, stride)) + [h - tile]
    xs = list(range(0, w - tile, stride)) + [w - tile]
    coords = [(y0, x0) for y0 in ys for x0 in xs]
    acc = torch.zeros(b, c, h * scale, w * scale, dtype=img_lq.dtype, device=img_lq.device)
    cover = torch.zeros_like(acc)
    # any other nn.Module is evaluated tile by tile exactly like the reference (it may couple the images of a batch)
    per_call = TILE_BATCH if isinstance(model, B200SRModel) else 1
    for k0 in range(0, len(coords), per_call):
        chunk = coords[k0:k0 + per_call]
        # (tiles, b, c, tile, tile) -> one batch of len(chunk) * b images
        patches = torch.stack([img_lq[..., y0:y0 + tile, x0:x0 + tile] for y0, x0 in chunk])
        out = model(patches.reshape(len(chunk) * b, c, tile, tile).contiguous())
        out = out.reshape(len(chunk), b, c, tile * scale, tile * scale)
        for k, (y0, x0) in enumerate(chunk):
            sl = (..., slice(y0 * scale, (y0 + tile) * scale), slice(x0 * scale, (x0 + tile) * scale))
            acc[sl] += out[k]
            cover[sl] += 1
    return acc.div_(cover)
