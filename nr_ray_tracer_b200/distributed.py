"""Multi-GPU rendering: one process per GPU, tile (row-block) partition, one framebuffer gather.

Camera::render is embarrassingly parallel over pixels (camera.rs:315-317): each rank renders the row-blocks
b with b % world == rank end to end (same Philox streams keyed by the global pixel index, so the assembled
image is bit-identical for every GPU count) and the only exchange is one gather of the owned rows to rank 0
over NCCL/NVLink (24.9 MB for 1080p).  No data-path collective exists inside the render itself.

What keeps the gather off the critical path (round 1 lost 4 % of the 8-GPU step here):
  * the kernels write a rank's rows PACKED (NRRT_RENDER_OUT_PACKED), so there is no pack step;
  * every buffer and the row permutation are built once per (image size, world) and reused (`FramebufferGather`);
  * rank 0 receives all parts into one stacked buffer and places the rows with ONE index_select kernel;
  * rows are interleaved singly for world > 1 (rows_per_block = 1): 1080 rows over 8 GPUs is 135 rows each, exactly, and
    every rank gets the same share of every part of the image.  (Measured at N = 8: blocks of 8 rows leave 4 % between
    the fastest and the slowest rank on the Cornell box — bright rows are cheap, floor rows are not — and the step waits
    for the slowest.)  The kernels size their work tiles to the row-block (csrc/nrrt_device.cu owned_pixel): with single
    rows a warp's block of 128 items is a 128-pixel piece of one row, so it stays compact in the image.

The partition arithmetic here must match owned_pixel()/owned_rows() in csrc/nrrt_device.cu.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

DEFAULT_ROWS_PER_BLOCK = 8  # single-GPU default of the C ABI (any value gives the same image)


def rows_per_block_for(world: int) -> int:
    """Row-block height used by render_distributed / bench.py: single rows once the image is shared out."""
    return 1 if world > 1 else DEFAULT_ROWS_PER_BLOCK


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


class FramebufferGather:
    """Reusable gather of packed per-rank rows into the full image on rank `dst`.

    `packed` (max_rows, W, 3) is the buffer a rank renders into (its first n_rows rows are valid, in ascending row
    order); gather() returns the assembled (H, W, 3) image on rank dst, None elsewhere.  All tensors are allocated
    once; a call issues one collective and, on dst, one index_select."""

    def __init__(self, height: int, width: int, rank: int, world: int, rows_per_block: int, device,
                 dtype=torch.float32, group: Optional[dist.ProcessGroup] = None, dst: int = 0):
        self.H, self.W, self.rank, self.world, self.R, self.dst, self.group = height, width, rank, world, rows_per_block, dst, group
        rows = [owned_rows(height, r, world, rows_per_block) for r in range(world)]
        self.n_rows = len(rows[rank])
        self.max_rows = max(len(r) for r in rows)
        self.packed = torch.zeros((self.max_rows, width, 3), dtype=dtype, device=device)
        self.stacked = self.parts = self.perm = self.full = None
        if rank == dst:
            self.stacked = torch.zeros((world * self.max_rows, width, 3), dtype=dtype, device=device)
            self.parts = [self.stacked[r * self.max_rows:(r + 1) * self.max_rows] for r in range(world)]
            perm = np.zeros(height, dtype=np.int64)  # destination row y comes from stacked row perm[y]
            for r in range(world):
                perm[rows[r]] = r * self.max_rows + np.arange(len(rows[r]))
            self.perm = torch.from_numpy(perm).to(device)
            self.full = torch.zeros((height, width, 3), dtype=dtype, device=device)

    def gather(self) -> Optional[torch.Tensor]:
        if self.world == 1:
            return self.packed[: self.n_rows]
        if self.rank == self.dst:
            dist.gather(self.packed, gather_list=self.parts, dst=self.dst, group=self.group)
            torch.index_select(self.stacked, 0, self.perm, out=self.full)
            return self.full
        dist.gather(self.packed, gather_list=None, dst=self.dst, group=self.group)
        return None


_GATHERS: Dict[tuple, FramebufferGather] = {}


def gather_framebuffer(fb: torch.Tensor, rank: int, world: int, rows_per_block: int = DEFAULT_ROWS_PER_BLOCK,
                       group: Optional[dist.ProcessGroup] = None, dst: int = 0) -> Optional[torch.Tensor]:
    """fb: (H, W, 3) tensor (CPU for gloo, CUDA for nccl) whose rows owned by `rank` are valid (a full-size,
    un-packed framebuffer).  Returns the assembled image on rank `dst`, None elsewhere.  Convenience wrapper that
    packs the owned rows first; render straight into FramebufferGather.packed to skip that copy."""
    if world == 1:
        return fb
    H, W = int(fb.shape[0]), int(fb.shape[1])
    key = (H, W, rank, world, rows_per_block, str(fb.device), fb.dtype, id(group), dst)
    g = _GATHERS.get(key)
    if g is None:
        g = _GATHERS[key] = FramebufferGather(H, W, rank, world, rows_per_block, fb.device, fb.dtype, group, dst)
        g.mine = torch.from_numpy(owned_rows(H, rank, world, rows_per_block)).to(fb.device)
    if g.n_rows:
        torch.index_select(fb, 0, g.mine, out=g.packed[: g.n_rows])
    return g.gather()


def render_distributed(ctx, cam, rank: int, world: int, seed: int = 0, mode: Optional[int] = None,
                       gather: Optional[FramebufferGather] = None, host_out: Optional[torch.Tensor] = None):
    """Scene::render over `world` GPUs (this process drives GPU `rank`'s context `ctx`, scene already uploaded).

    Renders this rank's rows packed into the gather buffer, gathers to rank 0 and — when host_out (a pinned
    (H, W, 3) float32 tensor) is given — copies the assembled image to the host on rank 0.
    Returns (image tensor on rank 0 or None, render stats, the FramebufferGather to pass back in next time)."""
    from . import _abi as A
    dev = torch.device("cuda", torch.cuda.current_device())
    R = rows_per_block_for(world)
    if gather is None:
        gather = FramebufferGather(cam.height, cam.width, rank, world, R, dev)
    _, st = ctx.render(cam, seed=seed, mode=A.MODE_AUTO if mode is None else mode, rank=rank, world=world,
                       rows_per_block=R, out_device_ptr=gather.packed.data_ptr(), packed=True)
    full = gather.gather()
    if world == 1:
        full = gather.packed.view(cam.height, cam.width, 3)
    if rank == 0 and host_out is not None:
        host_out.copy_(full, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host_out, st, gather
    return full, st, gather
