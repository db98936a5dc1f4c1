#!/usr/bin/env python3
"""Runs the five BASELINE.json configs on one GPU (product path: NRRT_MODE_AUTO) next to a bounded CPU sample of the
oracle, plus the traversal-only microbenchmark bench.py uses as the ceiling (>= 2 M coherent primary rays through
nrrt_trace_rays, NRRT_TRACE_COMPACT: closest hit only, 16 B out per ray) per scene.
Writes gpurun_out/configs.json and prints a markdown table.  Usage: run_configs.py [--quick]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
from nr_ray_tracer_b200 import _abi as A, api  # noqa: E402
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene  # noqa: E402
from oracle import oracle as O  # noqa: E402

CONFIGS = [
    ("C1", "scenes/spheres.toml", 400, 225, 100, 50, 100),
    ("C2a", "scenes/earth.toml", 1920, 1080, 256, None, 8),
    ("C2b", "scenes/noise.toml", 1920, 1080, 256, None, 8),
    ("C3", "scenes/cornell-box-scene.json", 1920, 1080, 1024, 50, 32),
    ("C4", "scenes/utah-teapot-scene.json", 1920, 1080, 1024, 50, 16),
    ("C5", "scenes/cornell-teapot-scene.json", 3840, 2160, 4096, 50, 2),
]


sys.path.insert(0, ROOT)
from bench import primary_rays  # noqa: E402  (the same ray set bench.py traces)


def main():
    quick = "--quick" in sys.argv
    ctx = api.Context(0)
    import torch
    out = []
    for name, scene, W, H, spp, depth, cpu_spp in CONFIGS:
        gpu_spp = max(1, spp // 16) if quick else spp
        g = load_scene(scene, camera_override=CameraConfig(width=W, height=H, samples_per_pixel=gpu_spp, ray_max_bounces=depth))
        host = api.HostScene(g, bvh=os.environ.get("BVH", "reference"))
        ctx.upload(host)
        cam = api.camera_build(g.camera.to_builder_config())
        fb = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
        ctx.render(cam, seed=0, out_device_ptr=fb.data_ptr(), max_slots=1 << 14)  # warm-up (small)
        t0 = time.perf_counter()
        _, st = ctx.render(cam, seed=0, out_device_ptr=fb.data_ptr())
        wall = time.perf_counter() - t0
        # traversal-only microbenchmark: coherent primary rays, no shading
        rays = torch.from_numpy(primary_rays(cam)).cuda()
        hits = torch.zeros((rays.shape[0], 16), dtype=torch.uint8, device="cuda")
        best = 0.0
        for _ in range(6):
            ts = ctx.trace_rays_device(rays.data_ptr(), rays.shape[0], hits.data_ptr(), compact=True)
            best = max(best, rays.shape[0] / ts["kernel_ms"] / 1e3)
        n_micro = rays.shape[0]
        del rays, hits
        # CPU sample
        gc = load_scene(scene, camera_override=CameraConfig(width=W, height=H, samples_per_pixel=cpu_spp, ray_max_bounces=depth))
        osc = O.OracleScene(gc)
        t0 = time.perf_counter()
        _, cnt = osc.render(O.camera_build(gc.camera.to_builder_config()), seed=0, n_threads=len(os.sched_getaffinity(0)))
        cpu_dt = time.perf_counter() - t0
        row = {"config": name, "scene": scene, "width": W, "height": H, "spp": gpu_spp, "depth": cam.ray_max_bounces,
               "paths": st["paths"], "segments": st["segments"], "segments_per_path": st["segments"] / st["paths"],
               "gpu_mrays_s": st["segments"] / st["device_ms"] / 1e3, "gpu_time_to_image_s": st["device_ms"] / 1e3,
               "gpu_wall_s": wall, "launches": st["launches"], "kernel_design": A.MODE_NAMES[st["mode"]],
               "primary_ray_traversal_mrays_s": best, "microbench_rays": int(n_micro),
               "frac_of_traversal_microbench": (st["segments"] / st["device_ms"] / 1e3) / best,
               "cpu_mrays_s": cnt["segments"] / cpu_dt / 1e6, "cpu_cores": len(os.sched_getaffinity(0)), "cpu_sample_spp": cpu_spp,
               "cpu_time_to_image_s_extrapolated": cpu_dt * gpu_spp / cpu_spp,
               "prims": g.count_primitives(), "nodes": host.desc.n_nodes, "instances": host.desc.n_instances}
        row["speedup_vs_cpu"] = row["gpu_mrays_s"] / row["cpu_mrays_s"]
        out.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
    print("\n| config | scene | size | spp | kernel | seg/path | GPU Mrays/s | GPU time-to-image | traversal microbench Mrays/s | fraction | CPU Mrays/s (cores) | CPU time-to-image (extrap.) | speed-up |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in out:
        print(f"| {r['config']} | {os.path.basename(r['scene'])} | {r['width']}x{r['height']} | {r['spp']} | {r['kernel_design']} | {r['segments_per_path']:.2f} | "
              f"{r['gpu_mrays_s']:.0f} | {r['gpu_time_to_image_s']:.2f} s | {r['primary_ray_traversal_mrays_s']:.0f} | {r['frac_of_traversal_microbench']:.2f} | "
              f"{r['cpu_mrays_s']:.1f} ({r['cpu_cores']}) | {r['cpu_time_to_image_s_extrapolated']:.0f} s | {r['speedup_vs_cpu']:.0f}x |")


if __name__ == "__main__":
    main()
