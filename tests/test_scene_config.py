"""Host scene loader (mirror of ray-tracer/src/scene_config.rs and cli.rs camera handling)."""
import json
import math
import os

import pytest

from nr_ray_tracer_b200 import _abi as A
from nr_ray_tracer_b200.scene_config import (CameraConfig, SceneError, build_scene_graph, load_scene, load_scene_file)
from tests.scenes_util import ALL_SCENES, ROOT, load


def kinds(g, kind):
    return [o for o in g.objects if o[0] == kind]


def test_every_shipped_scene_loads_with_expected_primitive_counts():
    expect = {"spheres.toml": 488, "earth.toml": 3, "noise.toml": 4, "cornell-box-scene.json": 18,
              "quads.toml": 5, "triangles.toml": 5, "simple-lights.toml": 4, "scale.json": 2, "cube-scene.json": 18}
    for name in ALL_SCENES:
        g = load(name)
        if name in expect:
            assert g.count_primitives() == expect[name], name
        else:
            assert g.count_primitives() > 1000  # synthesised teapot
        assert g.objects[g.root][0] == A.OBJ_GROUP


def test_v2_table_schema_and_v3_pair_schema_agree():
    # spheres.toml / earth.toml use tables keyed by id; the current serde schema uses [id, cfg] pairs
    cfg = load_scene_file(os.path.join(ROOT, "scenes", "earth.toml"))
    assert isinstance(cfg["textures"], dict) and isinstance(cfg["materials"], dict)
    g2 = build_scene_graph(cfg, ROOT)
    cfg3 = dict(cfg)
    cfg3["textures"] = [[k, v] for k, v in cfg["textures"].items()]
    cfg3["materials"] = [[k, v] for k, v in cfg["materials"].items()]
    g3 = build_scene_graph(cfg3, ROOT)
    assert [o[3] for o in g2.objects] == [o[3] for o in g3.objects]
    assert [t["kind"] for t in g2.textures] == [t["kind"] for t in g3.textures]
    assert len(g2.images) == 2 and g2.images[0].shape == (1024, 2048, 3)
    mats = {g2.materials[o[1]][1] for o in g2.objects if o[0] == A.OBJ_SPHERE}
    assert len(mats) == 3  # ground / earth / moon textures all distinct


def test_v1_index_schema_triangles():
    g = load("triangles.toml")
    tris = kinds(g, A.OBJ_TRIANGLE)
    assert len(tris) == 5
    assert len({t[1] for t in tris}) == 5  # five different materials addressed by integer index


def test_noise_scene_defaults():
    g = load("noise.toml")
    noise = [t for t in g.textures if t["kind"] == A.TEX_NOISE][0]
    marble = [t for t in g.textures if t["kind"] == A.TEX_MARBLE][0]
    assert noise["octaves"] == 8 and noise["f0"] == 0.2 and noise["f2"] == 0.5
    assert abs(noise["f1"] - 2 * math.pi / 3) < 1e-15      # Fbm::DEFAULT_LACUNARITY
    assert marble["octaves"] == 7 and marble["f0"] == 0.2 and marble.get("seed", 0) == 0
    metals = [m for m in g.materials if m[0] == A.MAT_METAL]
    assert sorted(m[2] for m in metals) == [0.05, 0.8, 0.9]


def test_cornell_instancing_structure():
    g = load("cornell-box-scene.json")
    assert len(kinds(g, A.OBJ_QUAD)) == 18
    assert len(kinds(g, A.OBJ_TRANSLATE)) == 4 and len(kinds(g, A.OBJ_SCALE)) == 2 and len(kinds(g, A.OBJ_ROTATE_Y)) == 2
    root_children = g.objects[g.root][2]
    assert len(root_children) == 3
    assert g.objects[root_children[0]][0] == A.OBJ_GROUP       # Ref to the cornell-box Scene instance
    assert g.objects[root_children[1]][0] == A.OBJ_TRANSLATE
    # ScaleU(0.25) = factor * ONE
    su = [o for o in kinds(g, A.OBJ_SCALE) if o[3][:3] == (0.25, 0.25, 0.25)]
    assert len(su) == 1
    # the light: DiffuseLight intensity 15 on the small quad
    lights = [m for m in g.materials if m[0] == A.MAT_DIFFUSE_LIGHT]
    assert len(lights) == 1 and lights[0][2] == 15.0
    # cube faces take the material passed down through Scene{path, material} (scene_config.rs:335-340)
    cube_mats = {o[1] for o in kinds(g, A.OBJ_QUAD)}
    kinds_used = {g.materials[m][0] for m in cube_mats}
    assert kinds_used == {A.MAT_LAMBERTIAN, A.MAT_METAL, A.MAT_DIFFUSE_LIGHT}


def test_material_and_texture_fallbacks(tmp_path):
    cfg = {"camera": {}, "scene": [{"Sphere": {"center": [0, 0, 0], "radius": 1.0}}]}
    g = build_scene_graph(cfg, str(tmp_path))
    sph = kinds(g, A.OBJ_SPHERE)[0]
    kind, tex, _ = g.materials[sph[1]]
    assert kind == A.MAT_LAMBERTIAN and g.textures[tex]["color"] == (0.5, 0.5, 0.5)  # scene_config.rs:421-443
    cfg["material_fallback"] = {"Metal": {"fuzz": 0.25}}
    cfg["texture_fallback"] = {"SolidColor": {"color": [1, 0, 0]}}
    g = build_scene_graph(cfg, str(tmp_path))
    kind, tex, param = g.materials[kinds(g, A.OBJ_SPHERE)[0][1]]
    assert kind == A.MAT_METAL and param == 0.25 and g.textures[tex]["color"] == (1.0, 0.0, 0.0)


def test_ref_shares_the_object_and_group_material_is_passed_down(tmp_path):
    cfg = {"camera": {},
           "textures": [["t", {"SolidColor": {"color": [0, 1, 0]}}]],
           "materials": [["m", {"Lambertian": {"texture": "t"}}]],
           "instances": [["s", {"Group": {"material": "m", "objects": [
               {"Sphere": {"center": [0, 0, 0], "radius": 1.0}},
               {"Quad": {"point": [0, 0, 0], "u": [1, 0, 0], "v": [0, 1, 0]}}]}}]],
           "scene": [{"Ref": {"id": "s"}}, {"Translate": {"offset": [3, 0, 0], "object": {"Ref": {"id": "s"}}}}]}
    g = build_scene_graph(cfg, str(tmp_path))
    r0, r1 = g.objects[g.root][2]
    assert g.objects[r1][0] == A.OBJ_TRANSLATE and g.objects[r1][2] == [r0]   # same object index: shared
    for o in kinds(g, A.OBJ_SPHERE) + kinds(g, A.OBJ_QUAD):
        assert g.textures[g.materials[o[1]][1]]["color"] == (0.0, 1.0, 0.0)


def test_checker_references_and_table_order(tmp_path):
    cfg = {"camera": {}, "textures": {"zz_chk": {"Checker": {"even": "a", "odd": "b", "scale": 4.0}},
                                      "a": {"SolidColor": {"color": [1, 1, 1]}},
                                      "b": {"SolidColor": {"color": [0, 0, 0]}},
                                      "aa_chk": {"Checker": {"even": "a"}}},
           "materials": {"m": {"Lambertian": {"texture": "zz_chk"}}},
           "scene": [{"Sphere": {"center": [0, 0, 0], "radius": 1.0, "material": "m"}}]}
    g = build_scene_graph(cfg, str(tmp_path))
    chk = [t for t in g.textures if t["kind"] == A.TEX_CHECKER]
    assert len(chk) == 2
    for t in chk:
        idx = g.textures.index(t)
        assert t["a"] < idx and t["b"] < idx  # sub-textures always precede (device loop relies on it)
    default = [t for t in chk if t["f0"] == 0.5][0]     # CheckerBuilder default scale (checker.rs:55)
    assert g.textures[default["b"]]["color"] == (0.0, 0.0, 0.0)
    with pytest.raises(SceneError):
        build_scene_graph({"camera": {}, "textures": [["c", {"Checker": {"even": "nope"}}]], "scene": []}, ".")


def test_error_paths(tmp_path):
    with pytest.raises(SceneError):
        load_scene(str(tmp_path / "scene.yaml"))
    with pytest.raises(SceneError):
        load_scene(str(tmp_path / "missing.toml"))
    bad = {"camera": {}, "scene": [{"Sphere": {"center": [0, 0, 0], "radius": 1.0, "material": "nope"}}]}
    with pytest.raises(SceneError):
        build_scene_graph(bad, ".")
    with pytest.raises(SceneError):
        build_scene_graph({"camera": {}, "scene": [{"Ref": {"id": "ghost"}}]}, ".")
    with pytest.raises(SceneError):
        build_scene_graph({"camera": {}, "scene": [{"Blob": {}}]}, ".")
    p = tmp_path / "s.json"
    p.write_text(json.dumps({"camera": {}, "textures": [["e", {"Image": {"path": "nope.jpg"}}]], "scene": []}))
    with pytest.raises(SceneError):
        load_scene(str(p), base_dir=str(tmp_path))


def test_camera_merge_size_rules_and_units():
    c = CameraConfig.from_dict({"field_of_view": 35, "samples_per_pixel": 200, "ray_max_bounces": 50,
                                "look_from": [0.5, 0.5, -1.625]})
    c.merge_with(CameraConfig(width=1920, height=1080, samples_per_pixel=1024))
    b = c.to_builder_config()
    assert (b.width, b.height, b.samples_per_pixel, b.ray_max_bounces) == (1920, 1080, 1024, 50)
    assert abs(b.field_of_view - 35 * math.pi / 180) < 1e-15            # degrees -> radians (cli.rs:369-371)
    assert list(b.view_up) == [0, 1, 0] and b.focus_dist == 1.0 and b.defocus_angle == 0.0
    d = CameraConfig().to_builder_config()                              # CameraBuilder defaults (camera.rs:162-203)
    assert (d.width, d.height, d.samples_per_pixel, d.ray_max_bounces) == (1200, 800, 10, 10)
    assert abs(d.field_of_view - math.pi / 2) < 1e-15 and list(d.look_from) == [1, 1, 1]
    assert CameraConfig(width=400, aspect_ratio=16 / 9).get_size() == (400, 225)
    assert CameraConfig(height=225, aspect_ratio=16 / 9).get_size() == (400, 225)
    for bad in (CameraConfig(width=10), CameraConfig(height=10), CameraConfig(aspect_ratio=2.0),
                CameraConfig(width=1, height=1, aspect_ratio=1.0)):
        with pytest.raises(SceneError):
            bad.get_size()
    # focal_length is parsed but never applied (cli.rs:229 / absent from merge_with and try_update)
    c2 = CameraConfig()
    c2.merge_with(CameraConfig(focal_length=3.0))
    assert c2.focal_length is None
