#!/usr/bin/env python3
"""Generates tests/golden/golden.npz from the CPU oracle (oracle/oracle.cpp).

The reference ships no golden vectors and cannot be run here (no Rust toolchain), so these fixtures
pin the ORACLE: the CPU-only suite checks that the oracle still reproduces them, and the GPU suite
checks the CUDA path against them as well as against the live oracle.  Regenerate with
    python tests/golden/make_golden.py
only when the oracle is deliberately changed, and say so in the commit.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.chdir(ROOT)

from nr_ray_tracer_b200 import _abi as A  # noqa: E402
from nr_ray_tracer_b200.scene_config import SceneGraph  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import kat  # noqa: E402
from tests.scenes_util import load  # noqa: E402

GOLDEN_SCENES = ["spheres.toml", "earth.toml", "noise.toml", "cornell-box-scene.json", "utah-teapot-scene.json",
                 "scale.json", "simple-lights.toml"]
RENDER_W, RENDER_H, RENDER_SPP, RENDER_SEED = 48, 27, 4, 11


def golden_rays(g):
    sp = kat.special_rays(g)
    rng = np.random.default_rng(5)
    sp = sp[rng.permutation(len(sp))[:500]]
    return np.concatenate([kat.random_rays(g, 500, seed=21), kat.aimed_rays(g, 500, seed=22), sp])


def texture_graph():
    """A synthetic scene exercising every texture kind (checker is used by no shipped scene)."""
    g = SceneGraph()
    t_a = g.add_texture(kind=A.TEX_SOLID, color=(0.9, 0.1, 0.2))
    t_b = g.add_texture(kind=A.TEX_SOLID, color=(0.1, 0.8, 0.3))
    t_chk = g.add_texture(kind=A.TEX_CHECKER, a=t_a, b=t_b, f0=7.5)
    t_noise = g.add_texture(kind=A.TEX_NOISE, seed=3, octaves=5, f0=1.7, f1=2.1, f2=0.45)
    t_marble = g.add_texture(kind=A.TEX_MARBLE, seed=1, octaves=7, f0=0.8)
    t_chk2 = g.add_texture(kind=A.TEX_CHECKER, a=t_chk, b=t_noise, f0=2.0)
    img = (np.arange(16 * 8 * 3, dtype=np.uint32) * 37 % 256).astype(np.uint8).reshape(8, 16, 3)
    t_img = g.add_texture(kind=A.TEX_IMAGE, a=g.add_image(img))
    mats = [g.add_material(A.MAT_LAMBERTIAN, t) for t in (t_chk, t_noise, t_marble, t_chk2, t_img)]
    mats.append(g.add_material(A.MAT_DIFFUSE_LIGHT, t_chk2, 3.0))
    mats.append(g.add_material(A.MAT_METAL, t_img, 0.3))
    mats.append(g.add_material(A.MAT_DIELECTRIC, 0, 1.5))
    objs = []
    for i, m in enumerate(mats):
        objs.append(g.add_object(A.OBJ_SPHERE, m, v=(2.5 * (i % 4) - 3.75, 1.0 + 2.2 * (i // 4), 0.3 * i, 1.0)))
    objs.append(g.add_object(A.OBJ_QUAD, mats[3], v=(-8, 0, -6, 16, 0, 0, 0, 0, 12)))
    g.root = g.add_object(A.OBJ_GROUP, children=objs)
    g.camera.look_from, g.camera.look_at = (0.0, 3.0, 12.0), (0.0, 1.5, 0.0)
    g.camera.field_of_view, g.camera.background_color = 40.0, (0.6, 0.7, 0.9)
    g.camera.ray_max_bounces = 8
    return g


def main():
    out = {}
    graphs = {name: load(name, width=RENDER_W, height=RENDER_H, samples_per_pixel=RENDER_SPP) for name in GOLDEN_SCENES}
    tg = texture_graph()
    tg.camera.width, tg.camera.height, tg.camera.samples_per_pixel = RENDER_W, RENDER_H, RENDER_SPP
    graphs["_textures"] = tg
    for name, g in graphs.items():
        key = name.split(".")[0].replace("-", "_")
        sc = O.OracleScene(g)
        rays = golden_rays(g)
        hits, _ = sc.trace_rays(rays)
        cam = O.camera_build(g.camera.to_builder_config())
        img, cnt = sc.render(cam, seed=RENDER_SEED)
        out[f"{key}__rays"] = rays
        out[f"{key}__hits"] = hits
        out[f"{key}__camera"] = np.frombuffer(bytes(cam), dtype=np.uint8)
        out[f"{key}__image"] = img
        out[f"{key}__segments"] = np.array([cnt["segments"]], dtype=np.uint64)
    # textures: direct evaluation
    sc = O.OracleScene(tg)
    rng = np.random.default_rng(8)
    uvp = np.concatenate([rng.random((400, 2)) * 1.2 - 0.1, rng.normal(size=(400, 3)) * 6.0], axis=1)
    uvp[:4, :2] = [[0, 0], [1, 1], [1, 0], [0, 1]]
    out["tex__uvp"] = uvp
    for ti in range(len(tg.textures)):
        out[f"tex__{ti}"] = sc.texture_eval(ti, uvp)
    out["perm_tables"] = np.stack([O.perm_table(s) for s in range(8)])
    out["philox"] = np.stack([O.philox(0, 0, 0, 0, 0), O.philox(2**64 - 1, 2**32 - 1, 2**32 - 1, 2**32 - 1, 2**32 - 1),
                              O.philox(0x299f31d0a4093822, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)])
    path = os.path.join(ROOT, "tests", "golden", "golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
