#!/usr/bin/env python3
"""Developer tool (GPU box): BASELINE config C1 (spheres.toml 400x225, 100 spp) and C2b (noise, 1080p, 256 spp)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nr_ray_tracer_b200 import _abi as A, api  # noqa: E402
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene  # noqa: E402
import torch  # noqa: E402
ctx = api.Context(0)
out = []
for name, w, h, spp, depth in (("spheres.toml", 400, 225, 100, 50), ("noise.toml", 1920, 1080, 256, None)):
    g = load_scene("scenes/" + name, camera_override=CameraConfig(width=w, height=h, samples_per_pixel=spp, ray_max_bounces=depth))
    ctx.upload(api.HostScene(g))
    cam = api.camera_build(g.camera.to_builder_config())
    fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    best = 0.0
    for _ in range(4):
        _, st = ctx.render(cam, seed=0, out_device_ptr=fb.data_ptr())
        best = max(best, st["segments"] / st["device_ms"] / 1e3)
    out.append(f"{name}={best:.0f}")
print(" ".join(out), flush=True)
