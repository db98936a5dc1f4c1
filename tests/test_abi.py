"""The C-ABI library loads on a CPU-only box, exports every symbol include/nrrt.h declares, its struct
layouts match the ctypes mirror, and it fails loudly (no fallback) when there is no GPU."""
import ctypes as C
import os

import numpy as np
import re

import pytest

from nr_ray_tracer_b200 import _abi as A
from nr_ray_tracer_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "nrrt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(nrrt_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if not n.endswith("_fn")))


def test_library_exports_every_declared_symbol():
    L = api.lib()
    fns = declared_functions()
    assert len(fns) >= 12
    for name in fns:
        assert hasattr(L, name), f"{name} declared in include/nrrt.h but not exported"


def test_struct_sizes_match_header():
    L = api.lib()
    for i, s in enumerate(api.ABI_STRUCTS):
        assert L.nrrt_abi_sizeof(i) == C.sizeof(s), (i, s.__name__)
    assert L.nrrt_abi_sizeof(len(api.ABI_STRUCTS)) == 0
    assert C.sizeof(A.Node) == 64  # one traversal node = half a 128-byte line


def test_host_build_rejects_malformed_graphs():
    from nr_ray_tracer_b200.scene_config import SceneGraph
    g = SceneGraph()
    t = g.add_texture(kind=A.TEX_SOLID, color=(1, 1, 1))
    m = g.add_material(A.MAT_LAMBERTIAN, t)
    s = g.add_object(A.OBJ_SPHERE, m, v=(0, 0, 0, 1))
    g.root = s  # root must be a GROUP
    with pytest.raises(api.NrrtError):
        api.HostScene(g)
    g2 = SceneGraph()
    g2.add_material(A.MAT_LAMBERTIAN, 5)  # texture out of range
    g2.root = g2.add_object(A.OBJ_GROUP, children=[])
    with pytest.raises(api.NrrtError):
        api.HostScene(g2)
    g3 = SceneGraph()
    t = g3.add_texture(kind=A.TEX_SOLID, color=(1, 1, 1))
    m = g3.add_material(A.MAT_LAMBERTIAN, t)
    s = g3.add_object(A.OBJ_SPHERE, m, v=(0, 0, 0, 1))
    for _ in range(A.MAX_INSTANCE_DEPTH + 1):  # nesting deeper than the compiled-in limit
        grp = g3.add_object(A.OBJ_GROUP, children=[s, s, s])
        s = g3.add_object(A.OBJ_TRANSLATE, children=[grp], v=(1, 0, 0))
    g3.root = g3.add_object(A.OBJ_GROUP, children=[s])
    with pytest.raises(api.NrrtError):
        api.HostScene(g3)


def test_empty_scene_builds():
    from nr_ray_tracer_b200.scene_config import SceneGraph
    g = SceneGraph()
    g.root = g.add_object(A.OBJ_GROUP, children=[])
    hs = api.HostScene(g)
    assert hs.desc.root == A.REF_NONE and hs.desc.n_nodes == 0


def test_no_cpu_fallback_without_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(api.NrrtError) as e:
        api.Context(0)
    assert e.value.code == A.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nr_ray_tracer_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "from oracle" not in text and "import oracle" not in text, f
                assert "liboracle" not in text, f
                assert "oracle/" not in text and "oracle.cpp" not in text, f


def test_work_item_chunks_tile_the_samples_and_ignore_the_partition():
    """nrrt_chunk_starts: the chunks of a pixel cover [0, spp) exactly, none empty, for every spp; the schedule is a
    function of spp and the whole image's size only (it has no rank / world argument at all), and its scratch stays bounded for large images."""
    from nr_ray_tracer_b200 import api
    for pixels in (1, 90000, 1920 * 1080, 3840 * 2160, 2 ** 31 - 1):
        for spp in list(range(1, 260)) + [1000, 1024, 4096, 65535, 1 << 20]:
            st = api.chunk_starts(spp, pixels)
            assert st[0] == 0 and st[-1] == spp and all(a < b for a, b in zip(st, st[1:])), (pixels, spp, st)
            assert len(st) - 1 <= 32
    st = api.chunk_starts(1024, 1920 * 1080)
    sizes = [b - a for a, b in zip(st, st[1:])]
    assert sizes == [37] * 27 + [25]
    assert len(api.chunk_starts(4096, 3840 * 2160)) - 1 == 8           # 4K: 8 equal chunks
    assert api.chunk_starts(0, 100) == [0, 1]                          # spp clamps to 1 (camera.rs:104)


def test_work_items_cover_every_owned_pixel_and_sample_exactly_once():
    """nrrt_work_items runs the kernels' own item decode (decode_item / owned_pixel, compiled for the host as well):
    for every partition the items of all ranks together hold each (pixel, sample) of the image exactly once, a rank's
    items lie on the rows distributed.owned_rows gives it, every pixel is cut at the chunk boundaries of
    nrrt_chunk_starts, and items come chunk-major with the pixels of a chunk in compact tiles."""
    from nr_ray_tracer_b200 import api
    from nr_ray_tracer_b200 import distributed as D
    cases = [(1, 1, 1), (1, 1, 257), (33, 17, 5), (64, 37, 40), (7, 100, 3), (400, 225, 9), (5, 3, 1000), (1, 64, 33),
             (129, 9, 2)]
    for W, H, spp in cases:
        starts = api.chunk_starts(spp, W * H)
        n_chunks = len(starts) - 1
        for world, R in ((1, 8), (1, 1), (2, 1), (2, 8), (3, 2), (3, 5), (8, 1), (8, 8), (5, 12), (4, 16), (70, 1)):
            seen = np.zeros((H, W), dtype=np.int64)          # samples of each pixel covered so far
            for rank in range(world):
                it = api.work_items(W, H, spp, rank, world, R).astype(np.int64)
                rows = D.owned_rows(H, rank, world, R)
                assert len(it) == len(rows) * W * n_chunks, (W, H, spp, world, R, rank)
                if len(it) == 0:
                    continue
                x, y, s0, s1 = it.T
                assert x.min() >= 0 and x.max() < W and np.isin(y, rows).all()
                npx = len(rows) * W
                c = np.arange(len(it)) // npx                 # chunk-major numbering
                assert np.array_equal(s0, np.asarray(starts)[c]) and np.array_equal(s1, np.asarray(starts)[c + 1])
                # every chunk visits the rank's pixels in the same order, each exactly once
                first = y[:npx] * W + x[:npx]
                assert len(np.unique(first)) == npx
                assert np.array_equal((y * W + x).reshape(n_chunks, npx), np.broadcast_to(first, (n_chunks, npx)))
                np.add.at(seen, (y, x), s1 - s0)
            assert (seen == spp).all(), (W, H, spp, world, R)
    # the tiles: on one GPU (row-blocks of 8) a run of 128 items is 16 columns x 8 rows; with single-row blocks
    # (what a shared image uses) it is 128 pixels of one row
    it = api.work_items(400, 225, 9).astype(np.int64)
    blk = it[3 * 128:4 * 128]
    assert np.ptp(blk[:, 0]) == 15 and np.ptp(blk[:, 1]) == 7 and len(np.unique(blk[:, 1] * 400 + blk[:, 0])) == 128
    it = api.work_items(400, 225, 9, rank=3, world=8, rows_per_block=1).astype(np.int64)
    blk = it[:128]
    assert np.ptp(blk[:, 1]) == 0 and blk[0, 1] == 3 and np.array_equal(blk[:, 0], np.arange(128))
    # the last, cut-off strip (225 = 28 * 8 + 1 rows): one row, walked along x
    last = api.work_items(400, 225, 9).astype(np.int64)[224 * 400:225 * 400]
    assert (last[:, 1] == 224).all() and np.array_equal(last[:, 0], np.arange(400))
    # arguments nrrt_render refuses
    assert len(api.work_items(0, 5, 1)) == 0 and len(api.work_items(5, 5, 1, rank=2, world=2)) == 0
