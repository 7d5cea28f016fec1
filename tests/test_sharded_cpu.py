"""world_size-2 gloo test of the batch-sharding host logic (no GPU, stand-in model)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stand_in(x):  # x4 "SR" that depends on each image only
    return torch.nn.functional.interpolate(x, scale_factor=4, mode="nearest") * 2 + x.mean(dim=(1, 2, 3), keepdim=True)


def _worker(rank, world, port, n_images, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ntire2022_esr_b200.sharded import forward_sharded, shard_bounds

    g = torch.Generator().manual_seed(11)
    imgs = torch.rand(n_images, 3, 8, 6, generator=g) if rank == 0 else None
    out = forward_sharded(_stand_in, imgs, n_images, (3, 8, 6), torch.float32, torch.device("cpu"))
    mine = forward_sharded(_stand_in, imgs, n_images, (3, 8, 6), torch.float32, torch.device("cpu"), gather=False)
    s, e = shard_bounds(n_images, world)[rank]
    assert mine.shape[0] == e - s
    if rank == 0:
        q.put(out.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [5, 2, 1])
def test_scatter_forward_gather_world2(n_images):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(11)
    imgs = torch.rand(n_images, 3, 8, 6, generator=g)
    np.testing.assert_array_equal(out, _stand_in(imgs).numpy())


def test_shard_bounds():
    from ntire2022_esr_b200.sharded import shard_bounds

    assert shard_bounds(64, 8) == [(8 * i, 8 * i + 8) for i in range(8)]
    assert shard_bounds(5, 2) == [(0, 3), (3, 5)]
    assert shard_bounds(1, 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
    for n, w in [(128, 8), (7, 3), (0, 2)]:
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def test_forward_bucketed_groups_equal_shapes_and_keeps_order():
    from ntire2022_esr_b200.sharded import forward_bucketed

    calls = []

    def model(x):
        calls.append(tuple(x.shape))
        return _stand_in(x)

    g = torch.Generator().manual_seed(5)
    shapes = [(9, 7), (6, 8), (9, 7), (5, 5), (6, 8), (9, 7)]
    imgs = [torch.rand(3, h, w, generator=g) for h, w in shapes]
    outs = forward_bucketed(model, imgs)
    assert sorted(calls) == sorted([(3, 3, 9, 7), (2, 3, 6, 8), (1, 3, 5, 5)])
    for im, o in zip(imgs, outs):
        torch.testing.assert_close(o, _stand_in(im[None])[0])
