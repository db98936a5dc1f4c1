#!/usr/bin/env python3
"""Developer tool (GPU box): pooled-kernel throughput against resident warps per SM (throttled through max_slots)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nr_ray_tracer_b200 import _abi as A, api  # noqa: E402
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene  # noqa: E402
ctx = api.Context(0)
NS = int(os.environ.get("NS", "64"))
for name, spp in (("utah-teapot-scene.json", 32), ("cornell-box-scene.json", 32)):
    g = load_scene("scenes/" + name, camera_override=CameraConfig(width=1920, height=1080, samples_per_pixel=spp, ray_max_bounces=50))
    hs = api.HostScene(g)
    ctx.upload(hs)
    cam = api.camera_build(g.camera.to_builder_config())
    ctx.render(cam, seed=1, mode=A.MODE_POOL, max_slots=4096)
    for w in (3, 6, 9, 12, 15, 18):
        _, st = ctx.render(cam, seed=1, mode=A.MODE_POOL, max_slots=148 * NS * w)
        print(f"{name} warps/SM<={w}: {st['segments'] / st['device_ms'] / 1e3:.0f} Mseg/s", flush=True)
