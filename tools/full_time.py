#!/usr/bin/env python3
"""Developer tool (GPU box): C3 and C4 at their full BASELINE sizes (1080p, 1024 spp, depth 50), product path."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nr_ray_tracer_b200 import _abi as A, api  # noqa: E402
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene  # noqa: E402
import torch  # noqa: E402
ctx = api.Context(0)
out = []
for name in ("cornell-box-scene.json", "utah-teapot-scene.json"):
    g = load_scene("scenes/" + name, camera_override=CameraConfig(width=1920, height=1080, samples_per_pixel=1024, ray_max_bounces=50))
    ctx.upload(api.HostScene(g))
    cam = api.camera_build(g.camera.to_builder_config())
    fb = torch.zeros((1080, 1920, 3), dtype=torch.float32, device="cuda")
    best = 0.0
    for _ in range(2):
        _, st = ctx.render(cam, seed=0, out_device_ptr=fb.data_ptr())
        best = max(best, st["segments"] / st["device_ms"] / 1e3)
    out.append(f"{name.split('-scene')[0]}={best:.0f}")
print(os.environ.get("NRRT_CHUNKS", "default"), " ".join(out), flush=True)
