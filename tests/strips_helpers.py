"""Test support for the N > 1 host logic (tests/test_strips_gloo.py, tests/test_abi.py): how a frame is cut into `world` strips of whole cell rows
(a cell row = 8 pixel rows), the scene is replicated, every rank renders its strip and the RGB8 strips are
gathered to rank 0.  Pixel id = py*W + px with row 0 at the bottom, so a strip is a contiguous byte range of the
image and the gather is a concatenation.  Backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Tuple

TILE = 8


def strip_bounds(nby: int, rank: int, world: int) -> Tuple[int, int]:
    """Cell rows [lo, hi) of `rank` - the same integer split the C library uses (ug_set_strip)."""
    return nby * rank // world, nby * (rank + 1) // world


def strip_pixel_rows(height: int, world: int) -> List[int]:
    nby = height // TILE
    return [(strip_bounds(nby, k, world)[1] - strip_bounds(nby, k, world)[0]) * TILE for k in range(world)]


def strip_nbytes(width: int, height: int, world: int) -> List[int]:
    return [3 * width * rows for rows in strip_pixel_rows(height, world)]


def gather_strips(strip, width: int, height: int, rank: int, world: int, gathered: Optional[list] = None, group=None):
    """Gather the (padded, equal-size) strip tensors to rank 0.  `strip` is a uint8 tensor of max(strip_nbytes)
    bytes whose first strip_nbytes[rank] bytes are this rank's RGB8 rows.  Returns the list of tensors on rank 0."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [strip]
    if rank == 0 and gathered is None:
        gathered = [torch.empty_like(strip) for _ in range(world)]
    dist.gather(strip, gathered if rank == 0 else None, dst=0, group=group)
    return gathered if rank == 0 else None


def assemble(gathered, width: int, height: int, world: int, out=None):
    """Concatenate the valid part of every gathered strip into the full RGB8 image (bottom row first)."""
    import torch
    sizes = strip_nbytes(width, height, world)
    if out is None:
        out = torch.empty(3 * width * height, dtype=torch.uint8, device=gathered[0].device)
    off = 0
    for k in range(world):
        out[off:off + sizes[k]].copy_(gathered[k][:sizes[k]], non_blocking=True)
        off += sizes[k]
    return out
