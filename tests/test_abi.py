"""The C-ABI library loads on a CPU-only box, exports every symbol include/nrrt.h declares, its struct
layouts match the ctypes mirror, and it fails loudly (no fallback) when there is no GPU."""
import ctypes as C
import os
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
