#!/usr/bin/env python3
"""Developer tool (one GPU): the benchmark workload as ONE rank of an 8-GPU partition sees it (rank 0 of world 8),
next to the whole image — the per-GPU rate a shared image can reach, without an 8-GPU box."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nr_ray_tracer_b200 import _abi as A, api  # noqa: E402
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene  # noqa: E402
import torch  # noqa: E402
ctx = api.Context(0)
g = load_scene("scenes/cornell-box-scene.json", camera_override=CameraConfig(width=1920, height=1080, samples_per_pixel=1024, ray_max_bounces=50))
ctx.upload(api.HostScene(g))
cam = api.camera_build(g.camera.to_builder_config())
fb = torch.zeros((1080, 1920, 3), dtype=torch.float32, device="cuda")
out = []
for world, R in ((1, 8), (8, 1), (8, 8), (8, 2), (4, 1), (2, 1)):
    best = 0.0
    for _ in range(3):
        _, st = ctx.render(cam, seed=0, rank=0, world=world, rows_per_block=R, out_device_ptr=fb.data_ptr(), packed=True)
        best = max(best, st["segments"] / st["device_ms"] / 1e3)
    out.append(f"w{world}R{R}={best:.0f}")
print(os.environ.get("NRRT_CHUNKS", "default"), os.environ.get("NRRT_B200_LIB", "").split("/")[-1], " ".join(out), flush=True)
