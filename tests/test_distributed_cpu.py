"""N>1 host path on CPU: row-block partition arithmetic and the framebuffer gather over torch.distributed
(gloo, world_size 2 and 3).  Each rank fills its owned rows from the CPU oracle (standing in for the GPU render,
which owns exactly the same pixels), the gather must reproduce the single-rank image bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nr_ray_tracer_b200 import distributed as D
from tests.scenes_util import load


def test_row_partition_is_a_disjoint_cover():
    for H in (1, 7, 8, 9, 27, 225, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            for R in (1, 8, 16):
                rows = [D.owned_rows(H, r, world, R) for r in range(world)]
                allr = np.sort(np.concatenate(rows))
                assert allr.tolist() == list(range(H)), (H, world, R)
                for r in range(world):
                    assert rows[r].tolist() == [y for a, b in D.owned_row_ranges(H, r, world, R) for y in range(a, b)]
                    for y in rows[r]:
                        assert (y // R) % world == r
    # balance at the benchmark size: 1080 rows in blocks of 8 over 8 GPUs -> 135 blocks, 16 or 17 per rank
    counts = [len(D.owned_rows(1080, r, 8, 8)) for r in range(8)]
    assert max(counts) - min(counts) <= 8 and sum(counts) == 1080


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, R, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        g = load("cornell-box-scene.json", width=W, height=H, samples_per_pixel=2)
        sc = O.OracleScene(g)
        cam = O.camera_build(g.camera.to_builder_config())
        fb = np.full((H, W, 3), np.nan, dtype=np.float32)  # rows of other ranks stay NaN
        for y0, y1 in D.owned_row_ranges(H, rank, world, R):
            part, _ = sc.render(cam, seed=9, pixel_range=(y0 * W, y1 * W), n_threads=1)
            fb[y0:y1] = part[y0:y1]
        full = D.gather_framebuffer(torch.from_numpy(fb), rank, world, R)
        if rank == 0:
            np.save(out_path, full.numpy())
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def _worker_packed(rank, world, port, H, W, out_path):
    """The production path: rows rendered PACKED into FramebufferGather.packed (NRRT_RENDER_OUT_PACKED layout: the
    rank's rows in ascending order), the production row-block height, one gather, one index_select on rank 0."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        R = D.rows_per_block_for(world)
        assert R == 1   # single rows: every rank gets the same share of every part of the image
        G = D.FramebufferGather(H, W, rank, world, R, torch.device("cpu"))
        for rep in range(2):  # buffers are reused across calls
            rows = D.owned_rows(H, rank, world, R)
            assert G.n_rows == len(rows)
            for k, y in enumerate(rows):  # a recognisable pattern: value = row * 1000 + column + repetition
                G.packed[k] = torch.arange(W, dtype=torch.float32)[:, None] + float(y) * 1000.0 + rep
            full = G.gather()
            if rank == 0:
                want = torch.arange(H, dtype=torch.float32)[:, None, None] * 1000.0 + \
                    torch.arange(W, dtype=torch.float32)[None, :, None] + torch.zeros(3) + rep
                assert torch.equal(full, want)
            else:
                assert full is None
        if rank == 0:
            np.save(out_path, full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H", [(2, 21), (3, 50)])
def test_gloo_packed_gather_of_row_blocks(tmp_path, world, H):
    out_path = str(tmp_path / "full.npy")
    mp.spawn(_worker_packed, args=(world, _free_port(), H, 5, out_path), nprocs=world, join=True)
    assert np.load(out_path).shape == (H, 5, 3)


@pytest.mark.parametrize("world,H,R", [(2, 20, 8), (3, 13, 2)])
def test_gloo_gather_reassembles_the_image(tmp_path, world, H, R):
    W = 16
    out_path = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), H, W, R, out_path), nprocs=world, join=True)
    full = np.load(out_path)
    from oracle import oracle as O
    g = load("cornell-box-scene.json", width=W, height=H, samples_per_pixel=2)
    ref, _ = O.OracleScene(g).render(O.camera_build(g.camera.to_builder_config()), seed=9)
    assert not np.isnan(full).any()
    assert np.array_equal(full, ref)
