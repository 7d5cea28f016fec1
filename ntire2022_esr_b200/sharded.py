"""Batched-image path: independent LR images sharded over the ranks of one box.

Images never interact (no BatchNorm; ESA statistics are per image), so the partition is the batch axis
and the only communication is the scatter of inputs / gather of outputs around the forward - there is
no collective inside the kernels and results are bit-identical to the single-GPU run of the same images
(tests/test_parity_gpu.py::test_batch_invariance...).  One process per GPU, `torch.distributed` (NCCL on
GPUs, gloo in the CPU tests of this host logic).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_images: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [start, end) image ranges; the first n % world ranks get one extra image."""
    base, extra = divmod(n_images, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < extra else 0)
        out.append((s, e))
        s = e
    return out


def _p2p(ops) -> None:
    """All point-to-point transfers of one phase as ONE grouped operation (NCCL: ncclGroupStart/End, the transfers
    run concurrently over NVLink instead of one after the other)."""
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()


def forward_sharded(model: Callable[[torch.Tensor], torch.Tensor], images: Optional[torch.Tensor], n_images: int,
                    shape_chw: Sequence[int], dtype: torch.dtype, device: torch.device, root: int = 0,
                    gather: bool = True, group=None, out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Rank `root` holds `images` (n_images, 3, H, W); every rank runs `model` on its shard.

    Returns the (n_images, 3, 4H, 4W) result on `root` when gather=True (None elsewhere; `out`, if given on the root,
    receives it); with gather=False every rank returns its own shard (true data-parallel serving: outputs stay sharded).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    c, h, w = shape_chw
    bounds = shard_bounds(n_images, world)
    s, e = bounds[rank]
    if world == 1:
        x = images.to(device)
    else:
        if rank == root:
            src = images.to(device)
            x = src[s:e]
            _p2p([dist.P2POp(dist.isend, src[rs:re_], r, group) for r, (rs, re_) in enumerate(bounds) if r != root and re_ > rs])
        else:
            x = torch.empty((e - s, c, h, w), dtype=dtype, device=device)
            if e > s:
                _p2p([dist.P2POp(dist.irecv, x, root, group)])
    y = model(x.contiguous()) if e > s else torch.empty((0, c, 4 * h, 4 * w), dtype=dtype, device=device)
    if not gather or world == 1:
        return y
    if rank == root:
        if out is None:
            out = torch.empty((n_images, c, 4 * h, 4 * w), dtype=dtype, device=device)
        elif tuple(out.shape) != (n_images, c, 4 * h, 4 * w) or out.dtype != dtype or not out.is_contiguous():
            raise ValueError("forward_sharded: `out` must be a contiguous (n_images, C, 4H, 4W) tensor of the model dtype")
        out[s:e].copy_(y)
        _p2p([dist.P2POp(dist.irecv, out[rs:re_], r, group) for r, (rs, re_) in enumerate(bounds) if r != root and re_ > rs])
        return out
    if e > s:
        _p2p([dist.P2POp(dist.isend, y.contiguous(), root, group)])
    return None


def forward_bucketed(model: Callable[[torch.Tensor], torch.Tensor], images: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Mixed-size batch (BASELINE.json configs[2]: DIV2K-shaped LR images): images of equal (H, W) are stacked
    and run as one batch, results come back in the original order.  No padding - padding an image changes
    its output (zero-padded convolutions, ESA pooling), so shapes are never mixed inside a launch."""
    buckets = {}
    for i, im in enumerate(images):
        if im.dim() != 3:
            raise ValueError("forward_bucketed expects (3, H, W) tensors")
        buckets.setdefault((im.shape[1], im.shape[2], im.dtype, im.device), []).append(i)
    out: List[Optional[torch.Tensor]] = [None] * len(images)
    for idxs in buckets.values():
        y = model(torch.stack([images[i] for i in idxs]))
        for k, i in enumerate(idxs):
            out[i] = y[k]
    return out  # type: ignore[return-value]
