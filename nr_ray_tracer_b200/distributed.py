"""Multi-GPU rendering: one process per GPU, tile (row-block) partition, one framebuffer gather.

Camera::render is embarrassingly parallel over pixels (camera.rs:315-317): each rank renders the row-blocks
b with b % world == rank end to end (same Philox streams keyed by the global pixel index, so the assembled
image is bit-identical for every GPU count) and the only exchange is one gather of the owned rows to rank 0
over NCCL/NVLink (24.9 MB for 1080p).  No data-path collective exists inside the render itself.

The partition arithmetic here must match owned_pixel()/owned_rows() in csrc/nrrt_device.cu.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

DEFAULT_ROWS_PER_BLOCK = 8


def owned_row_ranges(height: int, rank: int, world: int, rows_per_block: int = DEFAULT_ROWS_PER_BLOCK
                     ) -> List[Tuple[int, int]]:
    """[(y0, y1), ...] half-open row ranges owned by `rank`."""
    out = []
    b = rank
    while b * rows_per_block < height:
        y0 = b * rows_per_block
        out.append((y0, min(y0 + rows_per_block, height)))
        b += world
    return out


def owned_rows(height: int, rank: int, world: int, rows_per_block: int = DEFAULT_ROWS_PER_BLOCK) -> np.ndarray:
    rs = owned_row_ranges(height, rank, world, rows_per_block)
    if not rs:
        return np.zeros(0, dtype=np.int64)
    return np.concatenate([np.arange(a, b, dtype=np.int64) for a, b in rs])


def gather_framebuffer(fb: torch.Tensor, rank: int, world: int, rows_per_block: int = DEFAULT_ROWS_PER_BLOCK,
                       group: Optional[dist.ProcessGroup] = None, dst: int = 0) -> Optional[torch.Tensor]:
    """fb: (H, W, 3) float32 tensor (CPU for gloo, CUDA for nccl) whose rows owned by `rank` are valid.
    Returns the assembled image on rank `dst`, None elsewhere."""
    H = fb.shape[0]
    if world == 1:
        return fb
    max_rows = max(len(owned_rows(H, r, world, rows_per_block)) for r in range(world))
    mine = torch.from_numpy(owned_rows(H, rank, world, rows_per_block)).to(fb.device)
    packed = torch.zeros((max_rows,) + tuple(fb.shape[1:]), dtype=fb.dtype, device=fb.device)
    packed[: mine.numel()] = fb.index_select(0, mine)
    if rank == dst:
        parts = [torch.empty_like(packed) for _ in range(world)]
        dist.gather(packed, gather_list=parts, dst=dst, group=group)
        full = torch.empty_like(fb)
        for r in range(world):
            ys = torch.from_numpy(owned_rows(H, r, world, rows_per_block)).to(fb.device)
            full.index_copy_(0, ys, parts[r][: ys.numel()])
        return full
    dist.gather(packed, gather_list=None, dst=dst, group=group)
    return None
