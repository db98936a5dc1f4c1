#!/usr/bin/env python3
"""Developer check (GPU box, checked build): medium-size pooled renders of the plane-heavy scenes, so that the in-kernel
verification of the reject-only plane test (NRRT_CHECKED) sees tens of millions of primitive tests."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nr_ray_tracer_b200 import _abi as A, api  # noqa: E402
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene  # noqa: E402
ctx = api.Context(0)
for name in ("utah-teapot-scene.json", "cornell-teapot-scene.json", "cornell-box-scene.json", "triangles.toml", "quads.toml",
             "scale.json", "cube-scene.json"):
    g = load_scene("scenes/" + name, camera_override=CameraConfig(width=960, height=540, samples_per_pixel=4, ray_max_bounces=50))
    ctx.upload(api.HostScene(g))
    cam = api.camera_build(g.camera.to_builder_config())
    _, st = ctx.render(cam, seed=2, mode=A.MODE_POOL)
    print(name, st["segments"], "segments ok", flush=True)
print("checked_big done")
