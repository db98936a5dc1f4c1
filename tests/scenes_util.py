"""Shared scene lists / loaders for the tests."""
import os

from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALL_SCENES = ["spheres.toml", "earth.toml", "noise.toml", "cornell-box-scene.json", "utah-teapot-scene.json",
              "quads.toml", "triangles.toml", "simple-lights.toml", "scale.json", "cube-scene.json",
              "cornell-teapot-scene.json"]
BASELINE_SCENES = ["spheres.toml", "earth.toml", "noise.toml", "cornell-box-scene.json", "utah-teapot-scene.json",
                   "cornell-teapot-scene.json"]  # C1, C2 (x2), C3, C4, C5


def load(name, **camera):
    return load_scene(os.path.join(ROOT, "scenes", name), base_dir=ROOT,
                      camera_override=CameraConfig(**camera) if camera else None)
