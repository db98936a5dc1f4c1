#!/usr/bin/env python3
"""Developer check on a GPU box: KATs + small renders of every scene against the oracle, and a timing line.
(The formal versions live in tests/; this prints details for iteration.)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nr_ray_tracer_b200 import _abi as A, api  # noqa: E402
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import kat  # noqa: E402

SCENES = ["spheres.toml", "earth.toml", "noise.toml", "cornell-box-scene.json", "utah-teapot-scene.json",
          "quads.toml", "triangles.toml", "simple-lights.toml", "scale.json", "cube-scene.json"]


def main():
    quick = "--quick" in sys.argv
    ctx = api.Context(0)
    ok_all = True
    for name in SCENES:
        g = load_scene("scenes/" + name, camera_override=CameraConfig(width=160, height=90, samples_per_pixel=16))
        hs = api.HostScene(g)
        ctx.upload(hs)
        osc = O.OracleScene(g)
        n = 20000 if quick else 200000
        rays = np.concatenate([kat.random_rays(g, n), kat.aimed_rays(g, n), kat.special_rays(g)])
        ref, oc = osc.trace_rays(rays)
        for visit_all in (True, False):
            gpu, st = ctx.trace_rays(rays, visit_all=visit_all)
            res = kat.compare_hits(gpu, ref)
            ok = kat.hits_ok(res)
            ok_all &= ok
            print(f"KAT {name:28s} visit_all={int(visit_all)} ok={ok} {res} nodes/ray={st['node_visits']/len(rays):.1f} "
                  f"exact/ray={st['box_exact']/len(rays):.3f} prims/ray={st['prim_tests']/len(rays):.2f} "
                  f"(oracle aabb/ray={oc['aabb_tests']/len(rays):.1f} prims/ray={oc['prim_tests']/len(rays):.2f}) "
                  f"{len(rays)/st['kernel_ms']/1e3:.1f} Mrays/s")
        cam = api.camera_build(g.camera.to_builder_config())
        ocam = O.camera_build(g.camera.to_builder_config())
        assert bytes(cam) == bytes(ocam), "camera mismatch host vs oracle"
        oimg, ocn = osc.render(ocam, seed=5)
        for mode, mname in ((A.MODE_MEGAKERNEL, "mega"), (A.MODE_WAVEFRONT, "wavefront")):
            img, st = ctx.render(cam, seed=5, mode=mode)
            diff = np.abs(img.astype(np.float64) - oimg.astype(np.float64))
            rel = diff / np.maximum(1e-3, np.abs(oimg))
            nbad = int((rel > 1e-4).any(axis=2).sum())
            print(f"RENDER {name:25s} {mname:9s} max_abs={diff.max():.3e} pixels_off={nbad}/{img.shape[0]*img.shape[1]} "
                  f"segs gpu={st['segments']} oracle={ocn['segments']} paths={st['paths']} "
                  f"{st['segments']/st['device_ms']/1e3:.1f} Mseg/s launches={st['launches']}")
            ok_all &= st["segments"] > 0
    print("ALL OK" if ok_all else "SOME FAILED")
    # timing at a realistic size
    if not quick:
        for name, w, h, spp in (("cornell-box-scene.json", 1920, 1080, 64), ("spheres.toml", 1920, 1080, 16),
                                ("utah-teapot-scene.json", 1920, 1080, 32)):
            g = load_scene("scenes/" + name, camera_override=CameraConfig(width=w, height=h, samples_per_pixel=spp))
            hs = api.HostScene(g)
            ctx.upload(hs)
            cam = api.camera_build(g.camera.to_builder_config())
            for mode, mname in ((A.MODE_MEGAKERNEL, "mega"), (A.MODE_WAVEFRONT, "wavefront")):
                t0 = time.time()
                img, st = ctx.render(cam, seed=1, mode=mode)
                dt = time.time() - t0
                print(f"TIMING {name} {w}x{h}x{spp} {mname}: {st['segments']/st['device_ms']/1e3:.1f} Mseg/s "
                      f"device_ms={st['device_ms']:.1f} wall={dt*1e3:.1f}ms segs/path={st['segments']/st['paths']:.2f} "
                      f"launches={st['launches']} mean={img.mean():.5f}")
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
