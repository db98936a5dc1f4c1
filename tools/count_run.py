#!/usr/bin/env python3
"""Per-segment traversal counts of render rays (instrumented megakernel pass) — developer tool, GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nr_ray_tracer_b200 import api
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene
ctx = api.Context(0)
for name in sys.argv[1:] or ["cornell-box-scene.json", "spheres.toml", "utah-teapot-scene.json", "earth.toml", "noise.toml"]:
    g = load_scene("scenes/" + name, camera_override=CameraConfig(width=960, height=540, samples_per_pixel=4, ray_max_bounces=int(os.environ.get("DEPTH", "50"))))
    hs = api.HostScene(g, bvh=os.environ.get("BVH", "reference")); ctx.upload(hs)
    cam = api.camera_build(g.camera.to_builder_config())
    _, st = ctx.render(cam, seed=1, count=True)
    s = st["segments"]
    print(f"{name}: segs/path={s/st['paths']:.2f} nodes/seg={st['node_visits']/s:.2f} prims/seg={st['prim_tests']/s:.2f} "
          f"exact/seg={st['box_exact']/s:.4f} inst_entries/seg={st['inst_entries']/s:.2f} inst_root_miss/seg={st['inst_misses']/s:.2f}")
