"""Tile sharding of one frame over the ranks of a `torch.distributed` job (one process per GPU).

The reference hands 32x32 rectangles to its render threads from one mutex-protected counter
(`ViewData::GetNextRectangle`, source/backend/scene/view.cpp:236-271); tiles are independent, so across GPUs they are
dealt round-robin in the same serial order (statistically balanced, no exchange while tracing).  The only data that
crosses NVLink is the finished tiles: one `gather` to the rank that owns the frame.  Backend-agnostic (nccl on the
GPU box, gloo in the CPU tests).
"""
import numpy as np

from .scene import tiles as frame_tiles


def deal(rects, rank, world):
    """Rectangles of `rank`: every `world`-th tile in the reference's serial order."""
    return rects[rank::world]


def area(rects):
    return sum((r - l + 1) * (b - t + 1) for l, t, r, b in rects)


def padded_pixels(rects_all, world):
    """Length (pixels) of the per-rank buffers: the largest share, so that `gather` sees equal shapes."""
    return max(area(deal(rects_all, r, world)) for r in range(world))


def gather_frame(local, rects_all, width, height, dist, rank, world, dst=0):
    """`local`: torch tensor [padded_pixels * 4] float32 holding this rank's tiles rect-major.  Returns the assembled
    H x W x 4 numpy frame on `dst`, None elsewhere."""
    import torch
    bufs = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    if world > 1:
        dist.gather(local, bufs, dst=dst)
    else:
        bufs = [local]
    if rank != dst:
        return None
    img = np.zeros((height, width, 4), dtype=np.float32)
    for r in range(world):
        px = bufs[r].detach().cpu().numpy().reshape(-1, 4)
        pos = 0
        for l, t, rr, b in deal(rects_all, r, world):
            n = (rr - l + 1) * (b - t + 1)
            img[t:b + 1, l:rr + 1] = px[pos:pos + n].reshape(b - t + 1, rr - l + 1, 4)
            pos += n
    return img


__all__ = ["deal", "area", "padded_pixels", "gather_frame", "frame_tiles"]
