#!/usr/bin/env python3
"""Small renders for compute-sanitizer (developer tool, GPU box): every kernel design on a 64x36 Cornell box, the
pooled and fused kernels on the teapot mesh and the sphere field, a fixed-ray batch.
  compute-sanitizer --tool memcheck  python tools/sanitize_run.py
  compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nr_ray_tracer_b200 import _abi as A, api  # noqa: E402
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene  # noqa: E402
from tests import kat  # noqa: E402

ctx = api.Context(0)
for name, modes in (("cornell-box-scene.json", (A.MODE_FUSED, A.MODE_POOL, A.MODE_WAVEFRONT, A.MODE_MEGAKERNEL)),
                    ("utah-teapot-scene.json", (A.MODE_FUSED, A.MODE_POOL)), ("spheres.toml", (A.MODE_FUSED, A.MODE_POOL)),
                    ("noise.toml", (A.MODE_POOL,))):
    g = load_scene("scenes/" + name, camera_override=CameraConfig(width=64, height=36, samples_per_pixel=2, ray_max_bounces=50))
    hs = api.HostScene(g)
    ctx.upload(hs)
    cam = api.camera_build(g.camera.to_builder_config())
    ref = None
    for m in modes:
        img, st = ctx.render(cam, seed=1, mode=m)
        ref = img if ref is None else ref
        assert np.array_equal(img, ref)
        print(name, A.MODE_NAMES[m], st["segments"], flush=True)
    rays = np.concatenate([kat.random_rays(g, 2000), kat.special_rays(g)[:500]])
    hits, _ = ctx.trace_rays(rays)
    print(name, "trace", int((hits["object"] != 0xFFFFFFFF).sum()), flush=True)
print("sanitize_run done")
