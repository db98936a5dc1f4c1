#!/usr/bin/env python3
"""Timing sweep over kernel design x slot count (developer tool, GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nr_ray_tracer_b200 import _abi as A, api  # noqa: E402
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene  # noqa: E402


def main():
    scenes = sys.argv[1:] or ["cornell-box-scene.json:256", "spheres.toml:64", "utah-teapot-scene.json:128"]
    ctx = api.Context(0)
    for spec in scenes:
        name, spp = spec.split(":")
        g = load_scene("scenes/" + name, camera_override=CameraConfig(width=1920, height=1080, samples_per_pixel=int(spp),
                                                                      ray_max_bounces=50))
        hs = api.HostScene(g, bvh=os.environ.get("BVH", "reference"))
        ctx.upload(hs)
        cam = api.camera_build(g.camera.to_builder_config())
        ref = None
        for mode, mname in ((A.MODE_WAVEFRONT, "wavefront"), (A.MODE_MEGAKERNEL, "mega")):
            for slots in (1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22):
                img, st = ctx.render(cam, seed=1, mode=mode, max_slots=slots)
                if ref is None:
                    ref = img
                same = bool((img == ref).all())
                print(f"{name} spp={spp} {mname:9s} slots={slots:8d}: {st['segments']/st['device_ms']/1e3:8.1f} Mseg/s "
                      f"device_ms={st['device_ms']:8.1f} launches={st['launches']:6d} extend_share="
                      f"{st['extend_ms']/st['device_ms']:.2f} identical={same}", flush=True)


if __name__ == "__main__":
    main()
