"""The multi-GPU host logic (tile dealing + gather + frame assembly, povray_b200/shard.py) on CPU with the gloo backend,
world_size 2 and 3: every pixel of the frame comes from exactly one rank and lands where it belongs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from povray_b200 import shard
from povray_b200.scene import tiles


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_render(rects, width):
    """Stands in for pvgpu_render_device: pixel value encodes its frame position."""
    out = []
    for l, t, r, b in rects:
        ys, xs = np.mgrid[t:b + 1, l:r + 1]
        px = np.stack([xs, ys, ys * width + xs, np.ones_like(xs)], axis=-1).astype(np.float32)
        out.append(px.reshape(-1, 4))
    return np.concatenate(out, axis=0) if out else np.zeros((0, 4), dtype=np.float32)


def _worker(rank, world, port, w, h, block, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rects_all = tiles(w, h, block)
    mine = shard.deal(rects_all, rank, world)
    n = shard.padded_pixels(rects_all, world)
    local = torch.zeros(n * 4, dtype=torch.float32)
    px = _fake_render(mine, w)
    local[:px.size] = torch.from_numpy(px.reshape(-1))
    img = shard.gather_frame(local, rects_all, w, h, dist, rank, world)
    counts = torch.tensor([float(shard.area(mine))])
    dist.all_reduce(counts)
    if rank == 0:
        np.save(result_path, img)
        assert counts.item() == w * h
    else:
        assert img is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,w,h,block", [(2, 100, 70, 32), (3, 64, 64, 16), (2, 33, 5, 32)])
def test_tiles_are_dealt_and_gathered(tmp_path, world, w, h, block):
    path = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(world, _free_port(), w, h, block, path), nprocs=world, join=True)
    img = np.load(path)
    ys, xs = np.mgrid[0:h, 0:w]
    assert np.array_equal(img[..., 0], xs) and np.array_equal(img[..., 1], ys)
    assert np.array_equal(img[..., 2], (ys * w + xs).astype(np.float32))
    assert (img[..., 3] == 1).all()                     # every pixel written exactly once (frame starts at zero)


def test_deal_is_a_partition():
    rects = tiles(1920, 1080, 32)
    assert len(rects) == 60 * 34
    for world in (1, 2, 4, 8):
        parts = [shard.deal(rects, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(rects)
        sizes = [shard.area(p) for p in parts]
        assert max(sizes) - min(sizes) <= 32 * 32 * 2     # balanced to within two tiles
