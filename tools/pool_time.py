#!/usr/bin/env python3
"""Developer tool (GPU box): pooled-kernel timings at 1080p.  Usage: pool_time.py [mode ...] (default: pool)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nr_ray_tracer_b200 import _abi as A, api  # noqa: E402
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene  # noqa: E402

modes = sys.argv[1:] or ["pool"]
M = {"auto": A.MODE_AUTO, "pool": A.MODE_POOL, "fused": A.MODE_FUSED, "wavefront": A.MODE_WAVEFRONT, "mega": A.MODE_MEGAKERNEL}
ctx = api.Context(0)
out = []
for name, spp in (("utah-teapot-scene.json", 64), ("cornell-teapot-scene.json", 32), ("cornell-box-scene.json", 64),
                  ("spheres.toml", 32), ("noise.toml", 32), ("earth.toml", 64)):
    g = load_scene("scenes/" + name, camera_override=CameraConfig(width=1920, height=1080, samples_per_pixel=spp, ray_max_bounces=50))
    hs = api.HostScene(g, bvh=os.environ.get("BVH", "reference"))
    ctx.upload(hs)
    cam = api.camera_build(g.camera.to_builder_config())
    for m in modes:
        ctx.render(cam, seed=1, mode=M[m], max_slots=4096)
        _, st = ctx.render(cam, seed=1, mode=M[m])
        out.append(f"{name.split('-scene')[0].split('.')[0]}:{m}={st['segments'] / st['device_ms'] / 1e3:.0f}")
print(os.environ.get("NRRT_B200_LIB", "default").split("/")[-1], " ".join(out), flush=True)
