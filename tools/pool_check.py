#!/usr/bin/env python3
"""Developer check (GPU box): the pooled kernel's image against the fused kernel's on every scene, plus timings."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nr_ray_tracer_b200 import _abi as A, api  # noqa: E402
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene  # noqa: E402

SCENES = ["cornell-box-scene.json", "utah-teapot-scene.json", "spheres.toml", "earth.toml", "noise.toml", "quads.toml",
          "triangles.toml", "simple-lights.toml", "scale.json", "cube-scene.json", "cornell-teapot-scene.json"]
ctx = api.Context(0)
ok = True
for name in SCENES:
    g = load_scene("scenes/" + name, camera_override=CameraConfig(width=160, height=90, samples_per_pixel=8, ray_max_bounces=50))
    hs = api.HostScene(g)
    ctx.upload(hs)
    cam = api.camera_build(g.camera.to_builder_config())
    a, sa = ctx.render(cam, seed=3, mode=A.MODE_FUSED)
    b, sb = ctx.render(cam, seed=3, mode=A.MODE_POOL)
    same = bool(np.array_equal(a, b)) and sa["segments"] == sb["segments"] and sa["paths"] == sb["paths"]
    ok &= same
    print(f"{name:28s} identical={same} segs {sa['segments']} / {sb['segments']} paths {sa['paths']} / {sb['paths']} "
          f"maxdiff={float(np.abs(a - b).max()):.3e}", flush=True)
print("POOL OK" if ok else "POOL MISMATCH")
for name, spp in (("utah-teapot-scene.json", 64), ("cornell-teapot-scene.json", 32), ("cornell-box-scene.json", 64),
                  ("spheres.toml", 32), ("earth.toml", 32), ("noise.toml", 32)):
    g = load_scene("scenes/" + name, camera_override=CameraConfig(width=1920, height=1080, samples_per_pixel=spp, ray_max_bounces=50))
    hs = api.HostScene(g)
    ctx.upload(hs)
    cam = api.camera_build(g.camera.to_builder_config())
    for mode, mname in ((A.MODE_FUSED, "fused"), (A.MODE_POOL, "pool")):
        ctx.render(cam, seed=1, mode=mode, max_slots=4096)
        _, st = ctx.render(cam, seed=1, mode=mode)
        print(f"TIMING {name:28s} {mname:6s} {st['segments'] / st['device_ms'] / 1e3:8.1f} Mseg/s device_ms={st['device_ms']:.1f}", flush=True)
sys.exit(0 if ok else 1)
