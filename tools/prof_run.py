#!/usr/bin/env python3
"""Short, deterministic render used under ncu (developer tool, GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nr_ray_tracer_b200 import _abi as A, api  # noqa: E402
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cornell-box-scene.json"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mode = {"mega": A.MODE_MEGAKERNEL, "wavefront": A.MODE_WAVEFRONT, "pool": A.MODE_POOL, "fused": A.MODE_FUSED,
        "auto": A.MODE_AUTO}[sys.argv[3] if len(sys.argv) > 3 else "auto"]
ctx = api.Context(0)
g = load_scene("scenes/" + name, camera_override=CameraConfig(width=1920, height=1080, samples_per_pixel=spp,
                                                              ray_max_bounces=50))
hs = api.HostScene(g, bvh=os.environ.get("BVH", "reference"))
ctx.upload(hs)
cam = api.camera_build(g.camera.to_builder_config())
if os.environ.get("WARMUP", "1") != "0":   # the first launch pays module loading (tens of ms on a cold box)
    warm = api.camera_build(load_scene("scenes/" + name, camera_override=CameraConfig(
        width=64, height=36, samples_per_pixel=1, ray_max_bounces=50)).camera.to_builder_config())
    ctx.render(warm, seed=1, mode=mode)
img, st = ctx.render(cam, seed=1, mode=mode)
print(f"{name} spp={spp}: {st['segments']/st['device_ms']/1e3:.1f} Mseg/s segments={st['segments']} device_ms={st['device_ms']:.1f} "
      f"launches={st['launches']} mean={img.mean():.6f}")
