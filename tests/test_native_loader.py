"""Native C++ scene loader (csrc/scene_loader.cpp): JSON + TOML parsers, scene builder, camera
merge rules and the baseline JPEG decoder, checked against the Python loader / PIL on every shipped scene."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from nr_ray_tracer_b200 import _abi as A
from nr_ray_tracer_b200 import api
from nr_ray_tracer_b200.scene_config import CameraConfig, load_scene
from tests.scenes_util import ALL_SCENES, ROOT

SCENES = ALL_SCENES + ["cornell-teapot-scene.json", "cornell-box-model.json", "cube-model.toml"]


def graph_tuple_native(g):
    objs = [(o.kind, o.material if o.kind <= A.OBJ_TRIANGLE else 0,
             [g.child_ids[o.first_child + k] for k in range(o.n_children)], tuple(o.v)) for o in
            (g.objects[i] for i in range(g.n_objects))]
    mats = [(m.kind, m.texture, m.param) for m in (g.materials[i] for i in range(g.n_materials))]
    texs = [(t.kind, t.a, t.b, t.seed, t.octaves, tuple(t.color), t.f0, t.f1, t.f2) for t in
            (g.textures[i] for i in range(g.n_textures))]
    return objs, mats, texs


def graph_tuple_python(g):
    objs = [(k, m if k <= A.OBJ_TRIANGLE else 0, list(ch), tuple(v)) for k, m, ch, v in g.objects]
    mats = [(k, t, p) for k, t, p in g.materials]
    texs = [(t["kind"], t.get("a", 0), t.get("b", 0), t.get("seed", 0), t.get("octaves", 0),
             tuple(float(x) for x in t.get("color", (0, 0, 0))), float(t.get("f0", 0)), float(t.get("f1", 0)),
             float(t.get("f2", 0))) for t in g.textures]
    return objs, mats, texs


@pytest.mark.parametrize("name", SCENES)
def test_native_loader_matches_python_loader(name):
    path = os.path.join(ROOT, "scenes", name)
    nat = api.NativeScene(path, base_dir=ROOT)
    py = load_scene(path, base_dir=ROOT)
    on, mn, tn = graph_tuple_native(nat.graph)
    op, mp, tp = graph_tuple_python(py)
    assert mn == mp and tn == tp
    assert on == op
    assert nat.graph.root == py.root
    assert nat.graph.n_images == len(py.images)
    for i, im in enumerate(py.images):
        ni = nat.graph.images[i]
        assert (ni.height, ni.width) == im.shape[:2]
        got = np.ctypeslib.as_array(C.cast(ni.rgb, C.POINTER(C.c_uint8)), shape=im.shape)
        assert np.array_equal(got, im), "baseline JPEG decoder differs from PIL"
    # camera: file section, then the CLI's override merge (render.rs:109) and builder defaults
    assert bytes(nat.camera_config()) == bytes(py.camera.to_builder_config())
    over = A.CameraFile(present=A.CAM_WIDTH | A.CAM_HEIGHT | A.CAM_SPP | A.CAM_BOUNCES, width=1920, height=1080,
                        samples_per_pixel=1024, ray_max_bounces=50)
    py2 = load_scene(path, base_dir=ROOT, camera_override=CameraConfig(width=1920, height=1080, samples_per_pixel=1024,
                                                                       ray_max_bounces=50))
    assert bytes(nat.camera_config(over)) == bytes(py2.camera.to_builder_config())
    # and the flattened device layout is byte-identical whichever loader fed the host layer
    h1, h2 = api.HostScene(nat), api.HostScene(py)
    assert (h1.desc.n_nodes, h1.desc.n_spheres, h1.desc.n_planes, h1.desc.n_instances) == \
           (h2.desc.n_nodes, h2.desc.n_spheres, h2.desc.n_planes, h2.desc.n_instances)
    assert np.array_equal(h1.nodes(), h2.nodes()) and np.array_equal(h1.child_boxes(), h2.child_boxes())


def test_camera_size_rules_and_defaults():
    L = api.lib()
    cfg = A.CameraConfig()
    f = A.CameraFile()
    assert L.nrrt_camera_file_to_config(C.byref(f), C.byref(cfg)) == 0
    assert (cfg.width, cfg.height, cfg.samples_per_pixel, cfg.ray_max_bounces) == (1200, 800, 10, 10)
    assert abs(cfg.field_of_view - math.pi / 2) < 1e-15 and list(cfg.look_from) == [1, 1, 1] and list(cfg.view_up) == [0, 1, 0]
    f = A.CameraFile(present=A.CAM_WIDTH | A.CAM_ASPECT_RATIO, width=400, aspect_ratio=16 / 9)
    assert L.nrrt_camera_file_to_config(C.byref(f), C.byref(cfg)) == 0 and (cfg.width, cfg.height) == (400, 225)
    f = A.CameraFile(present=A.CAM_HEIGHT | A.CAM_ASPECT_RATIO, height=225, aspect_ratio=16 / 9)
    assert L.nrrt_camera_file_to_config(C.byref(f), C.byref(cfg)) == 0 and (cfg.width, cfg.height) == (400, 225)
    for bad in (A.CAM_WIDTH, A.CAM_HEIGHT, A.CAM_ASPECT_RATIO, A.CAM_WIDTH | A.CAM_HEIGHT | A.CAM_ASPECT_RATIO):
        f = A.CameraFile(present=bad, width=10, height=10, aspect_ratio=1.0)
        assert L.nrrt_camera_file_to_config(C.byref(f), C.byref(cfg)) == A.ERR_INVALID


def test_parsers_edge_cases(tmp_path):
    # TOML: comments, dotted headers under an array of tables, inline tables, multi-line arrays, literal strings
    (tmp_path / "a.toml").write_text('''
# comment
[camera]          # trailing comment
look_from = [ 1.0, 2,
              3e0 ]   # ints and floats mix
field_of_view = 4_0.0
[textures.t1.SolidColor]
color = [1, 0.5, 0.25]
[materials]
m1 = { Metal = { fuzz = 0.5, texture = 't1' } }
[[scene]]
[scene.Sphere]
center = [0, -1.5e+0, 0]
radius = 1
material = "m1"
[[scene]]
Quad = { point = [0,0,0], u = [1,0,0], v = [0,1,0] }
''')
    n = api.NativeScene(str(tmp_path / "a.toml"), base_dir=str(tmp_path))
    p = load_scene(str(tmp_path / "a.toml"), base_dir=str(tmp_path))
    assert graph_tuple_native(n.graph) == graph_tuple_python(p)
    assert list(n.camera_file.look_from) == [1.0, 2.0, 3.0] and n.camera_file.field_of_view_deg == 40.0
    # JSON: escapes, nulls, nested arrays
    (tmp_path / "b.json").write_text(json.dumps({
        "camera": {"width": None, "look_at": [0, 0, -1], "samples_per_pixel": 3},
        "textures": [['a\u00e9"x', {"SolidColor": {"color": [0.1, 0.2, 0.3]}}]],
        "materials": [["m", {"Lambertian": {"texture": 'a\u00e9"x'}}]],
        "scene": [{"Group": {"material": "m", "objects": [{"Sphere": {"center": [0, 0, 0], "radius": 1e-1}}]}}]}))
    n = api.NativeScene(str(tmp_path / "b.json"), base_dir=str(tmp_path))
    p = load_scene(str(tmp_path / "b.json"), base_dir=str(tmp_path))
    assert graph_tuple_native(n.graph) == graph_tuple_python(p)
    assert n.camera_file.samples_per_pixel == 3 and not (n.camera_file.present & A.CAM_WIDTH)


def test_loader_errors(tmp_path):
    with pytest.raises(api.NrrtError):
        api.NativeScene(str(tmp_path / "missing.toml"))
    (tmp_path / "x.yaml").write_text("a: 1")
    with pytest.raises(api.NrrtError):
        api.NativeScene(str(tmp_path / "x.yaml"))
    for i, text in enumerate(['[camera\n', 'a = \n', '[[scene]]\nSphere = { center = [0,0,0] radius = 1 }\n',
                              '[[scene]]\n[scene.Blob]\nx = 1\n', '[[scene]]\n[scene.Ref]\nid = "ghost"\n',
                              '[textures.t.Image]\npath = "nope.jpg"\n']):
        f = tmp_path / f"bad{i}.toml"
        f.write_text(text)
        with pytest.raises(api.NrrtError):
            api.NativeScene(str(f), base_dir=str(tmp_path))
    (tmp_path / "bad.json").write_text('{"scene": [}')
    with pytest.raises(api.NrrtError):
        api.NativeScene(str(tmp_path / "bad.json"))
    (tmp_path / "notjpeg.jpg").write_bytes(b"\x89PNG....")
    (tmp_path / "img.toml").write_text('[textures.t.Image]\npath = "notjpeg.jpg"\n')
    with pytest.raises(api.NrrtError):
        api.NativeScene(str(tmp_path / "img.toml"), base_dir=str(tmp_path))
