"""Scene generators: the reference's `nr-ray-tracer create <kind>` authoring commands (§8(f) N4).

Host-side tooling, no GPU involved.  Every generator returns the reference's current `SceneConfig` document
(packages/ray-tracer/src/scene_config.rs:383-404: textures / materials as lists of `[id, {Kind: {...}}]` pairs,
objects as externally tagged enums) and `dumps()` writes it as JSON (serde_json::to_string_pretty layout) or TOML.
The deterministic generators restate the reference's literals, so their output describes the same scenes as the
files the reference ships under scenes/ (tests/test_create.py compares the loaded graphs):

    cornell-box    commands/create/cornell_box.rs   -> scenes/cornell-box-model.json
    cube           commands/create/cube.rs          -> scenes/cube-model.toml
    earth          commands/create/earth.rs         -> scenes/earth.toml
    noise          commands/create/noise.rs         -> scenes/noise.toml
    quads          commands/create/quads.rs         -> scenes/quads.toml
    triangles      commands/create/triangles.rs     -> scenes/triangles.toml
    simple-lights  commands/create/simple_lights.rs -> scenes/simple-lights.toml
    spheres        commands/create/spheres.rs       (random: same layout rules and material mix, but Python's
                                                     generator instead of ChaCha8 — positions/colours differ)
    convert-stl    commands/create/convert_stl.rs   binary STL -> normalised triangle group

    python -m nr_ray_tracer_b200.create quads -o quads.toml
    python -m nr_ray_tracer_b200.create convert-stl model.stl -F json -o model.json
"""
from __future__ import annotations

import argparse
import itertools
import json
import math
import os
import random
import struct
import sys
from typing import Any, Dict, List, Optional, Sequence, Tuple

Vec = Tuple[float, float, float]
X, Y, Z = (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)
ZERO, ONE = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
_CAMERA_FIELDS = ("width", "height", "aspect_ratio", "background_color", "look_at", "look_from", "view_up",
                  "focal_length", "field_of_view", "defocus_angle", "focus_distance", "samples_per_pixel",
                  "ray_max_bounces")


def _mul(k: float, v: Vec) -> Vec:
    return (k * v[0], k * v[1], k * v[2])


def _neg(v: Vec) -> Vec:
    return _mul(-1.0, v)


def _add(a: Vec, b: Vec) -> Vec:
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def _sub(a: Vec, b: Vec) -> Vec:
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def _clean(v: Vec) -> List[float]:
    return [0.0 if x == 0 else float(x) for x in v]   # no "-0.0" in the output, like the shipped files


class _Ids:
    """create.rs:166-179: process-wide counters tex_0000000 / mat_0000000."""

    def __init__(self):
        self.tex = self.mat = 0

    def next_texture(self) -> str:
        self.tex += 1
        return f"tex_{self.tex - 1:07d}"

    def next_material(self) -> str:
        self.mat += 1
        return f"mat_{self.mat - 1:07d}"


class SceneDoc:
    """SceneConfig (scene_config.rs:383-404) as plain data."""

    def __init__(self, **camera):
        self.camera: Dict[str, Any] = {k: None for k in _CAMERA_FIELDS}
        self.merge_camera(camera)
        self.textures: List[list] = []
        self.materials: List[list] = []
        self.scene: List[dict] = []
        self.ids = _Ids()

    def merge_camera(self, other: Optional[Dict[str, Any]]):
        """cli.rs:316-355 merge_with: every Some replaces (focal_length is not merged)."""
        for k, v in (other or {}).items():
            if k not in _CAMERA_FIELDS:
                raise KeyError(f"unknown camera field {k}")
            if v is not None and k != "focal_length":
                self.camera[k] = _clean(v) if isinstance(v, (tuple, list)) else v

    # -- textures / materials
    def solid(self, color: Vec, name: Optional[str] = None) -> str:
        name = name or self.ids.next_texture()
        self.textures.append([name, {"SolidColor": {"color": _clean(color)}}])
        return name

    def texture(self, name: Optional[str], kind: str, **fields) -> str:
        name = name or self.ids.next_texture()
        self.textures.append([name, {kind: fields}])
        return name

    def lambertian(self, texture: str, name: Optional[str] = None) -> str:
        name = name or self.ids.next_material()
        self.materials.append([name, {"Lambertian": {"texture": texture}}])
        return name

    def metal(self, texture: str, fuzz: float, name: Optional[str] = None) -> str:
        name = name or self.ids.next_material()
        self.materials.append([name, {"Metal": {"fuzz": float(fuzz), "texture": texture}}])
        return name

    def dielectric(self, refraction_index: float, name: Optional[str] = None) -> str:
        name = name or self.ids.next_material()
        self.materials.append([name, {"Dielectric": {"refraction_index": float(refraction_index)}}])
        return name

    def light(self, texture: str, intensity: float, name: Optional[str] = None) -> str:
        name = name or self.ids.next_material()
        self.materials.append([name, {"DiffuseLight": {"intensity": float(intensity), "texture": texture}}])
        return name

    # -- objects
    @staticmethod
    def _with_material(d: dict, material: Optional[str]) -> dict:
        if material is not None:                      # skip_serializing_if = "Option::is_none"
            d["material"] = material
        return d

    def quad(self, point: Vec, u: Vec, v: Vec, material: Optional[str] = None, kind: str = "Quad") -> dict:
        o = {kind: self._with_material({"point": _clean(point), "u": _clean(u), "v": _clean(v)}, material)}
        self.scene.append(o)
        return o

    def triangle(self, point: Vec, u: Vec, v: Vec, material: Optional[str] = None) -> dict:
        return self.quad(point, u, v, material, kind="Triangle")

    def sphere(self, center: Vec, radius: float, material: Optional[str] = None) -> dict:
        o = {"Sphere": self._with_material({"center": _clean(center), "radius": float(radius)}, material)}
        self.scene.append(o)
        return o

    def to_dict(self) -> Dict[str, Any]:
        d: Dict[str, Any] = {"camera": dict(self.camera)}
        for key in ("textures", "materials"):          # skip_serializing_if = "Vec::is_empty"
            if getattr(self, key):
                d[key] = getattr(self, key)
        if self.scene:
            d["scene"] = self.scene
        return d


# ------------------------------------------------------------------ generators
def cornell_box(camera=None) -> SceneDoc:   # cornell_box.rs:59-136
    s = SceneDoc(background_color=ZERO, look_from=(0.0, 0.5, 0.0), look_at=ZERO, field_of_view=40.0,
                 ray_max_bounces=50, samples_per_pixel=200)
    s.merge_camera(camera)
    white = s.lambertian(s.solid((0.3450980392, 0.3568627451, 0.4392156863)))
    green = s.lambertian(s.solid((0.6509803922, 0.8901960784, 0.631372549)))
    red = s.lambertian(s.solid((0.9529411765, 0.5450980392, 0.6588235294)))
    light = s.light(s.solid(ONE), 15.0)
    s.quad(ZERO, X, Z, white)
    s.quad(ONE, _neg(X), _neg(Z), white)
    s.quad(Z, X, Y, white)
    s.quad(X, Y, Z, green)
    s.quad(ZERO, Y, Z, red)
    s.quad((0.670, 0.998, 0.598), _mul(0.234, _neg(X)), _mul(0.189, _neg(Z)), light)
    return s


def cube(camera=None) -> SceneDoc:   # cube.rs:14-83
    s = SceneDoc(background_color=ONE, look_from=(0.5, 1.0, 2.0), look_at=(0.5, 0.5, 0.0), field_of_view=40.0,
                 ray_max_bounces=50, samples_per_pixel=200)
    s.merge_camera(camera)
    for point, u, v in ((ZERO, X, Y), (ZERO, X, Z), (ZERO, Z, Y), (X, Z, Y), (Y, X, Z), (Z, X, Y)):
        s.quad(point, u, v)
    return s


def earth(camera=None) -> SceneDoc:   # earth.rs:14-84
    s = SceneDoc(background_color=(0.7, 0.8, 1.0), look_from=(60.0, 20.0, 3.0), look_at=_mul(10.0, Y),
                 field_of_view=20.0, ray_max_bounces=10, samples_per_pixel=10)
    s.merge_camera(camera)
    s.solid(_mul(0.5, ONE), "ground")
    s.texture("earth", "Image", path="scenes/textures/earth.jpg")
    s.texture("moon", "Image", path="scenes/textures/moon.jpg")
    for name in ("ground", "earth", "moon"):
        s.lambertian(name, name)
    s.sphere(_mul(1000.0, _neg(Y)), 1000.0, "ground")
    s.sphere((0.0, 10.0, 0.0), 10.0, "earth")
    s.sphere((-12.0, 12.0, -20.0), 3.0, "moon")
    return s


def noise(camera=None) -> SceneDoc:   # noise.rs:14-122
    s = SceneDoc(background_color=(0.7, 0.8, 1.0), look_from=(30.0, 20.0, -60.0), look_at=(20.0, 10.0, -20.0),
                 field_of_view=30.0, ray_max_bounces=10, samples_per_pixel=10)
    s.merge_camera(camera)
    s.solid(_mul(0.5, ONE), "ground")
    s.lambertian("ground", "ground")
    s.sphere(_mul(1_000_000.0, _neg(Y)), 1_000_000.0, "ground")
    s.texture("sphere1", "Noise", frequency=0.2, octaves=8)      # None fields are skipped by serde
    s.metal("sphere1", 0.05, "sphere1")
    s.sphere((-30.0, 10.0, 10.0), 10.0, "sphere1")
    s.texture("sphere2", "Marble", frequency=0.2)
    s.metal("sphere2", 0.9, "sphere2")
    s.sphere((20.0, 10.0, -20.0), 10.0, "sphere2")
    s.solid((1.0, 0.5, 0.65), "sphere3")
    s.metal("sphere3", 0.8, "sphere3")
    s.sphere((10.0, 10.0, 25.0), 10.0, "sphere3")
    return s


_FIVE = (("red", (1.0, 0.2, 0.2)), ("green", (0.2, 1.0, 0.2)), ("blue", (0.2, 0.2, 1.0)), ("orange", (1.0, 0.5, 0.0)),
         ("cyan", (0.2, 0.8, 0.8)))
_FIVE_SHAPES = (((-3.0, -2.0, 5.0), _mul(4.0, _neg(Z)), _mul(4.0, Y)), ((-2.0, -2.0, 0.0), _mul(4.0, X), _mul(4.0, Y)),
                ((3.0, -2.0, 1.0), _mul(4.0, Z), _mul(4.0, Y)), ((-2.0, 3.0, 1.0), _mul(4.0, X), _mul(4.0, Z)),
                ((-2.0, -3.0, 5.0), _mul(4.0, X), _mul(4.0, _neg(Z))))


def _five(kind: str, camera) -> SceneDoc:   # quads.rs / triangles.rs (identical but for the shape)
    s = SceneDoc(background_color=(0.7, 0.8, 1.0), look_from=_mul(9.0, Z), look_at=ZERO, field_of_view=80.0,
                 ray_max_bounces=10, samples_per_pixel=10)
    s.merge_camera(camera)
    for name, color in _FIVE:
        s.solid(color, f"solid_{name}")
    for name, _ in _FIVE:
        s.lambertian(f"solid_{name}", f"lambertian_{name}")
    for (name, _), (p, u, v) in zip(_FIVE, _FIVE_SHAPES):
        s.quad(p, u, v, f"lambertian_{name}", kind=kind)
    return s


def quads(camera=None) -> SceneDoc:
    return _five("Quad", camera)


def triangles(camera=None) -> SceneDoc:
    return _five("Triangle", camera)


def simple_lights(camera=None) -> SceneDoc:   # simple_lights.rs:14-141
    s = SceneDoc(background_color=_mul(0.001, ONE), look_from=(26.0, 3.0, 6.0), look_at=_mul(2.0, Y),
                 field_of_view=20.0, ray_max_bounces=10, samples_per_pixel=10)
    s.merge_camera(camera)

    def ids():   # texture id first, then material id (simple_lights.rs:18-19)
        return s.ids.next_texture(), s.ids.next_material()
    t, m = ids()
    s.sphere(_mul(1_000_000.0, _neg(Y)), 1_000_000.0, s.lambertian(s.solid(_mul(0.4, ONE), t), m))
    t, m = ids()
    s.sphere(_mul(2.0, Y), 2.0, s.lambertian(s.texture(t, "Marble", frequency=0.2), m))
    t, m = ids()
    s.quad((3.0, 1.0, -2.0), _mul(2.0, X), _mul(2.0, Y), s.light(s.solid((1.00, 0.50, 0.25), t), 4.0, m))
    t, m = ids()
    s.sphere(_mul(7.0, Y), 1.0, s.light(s.solid((0.25, 0.50, 1.00), t), 4.0, m))
    return s


def spheres(camera=None, seed: int = 1) -> SceneDoc:   # spheres.rs:75-171
    """Same construction rules as the reference (ground, three large spheres, a 22x22 grid of small ones with a
    5 / 80 / 15 dielectric / Lambertian / metal mix); the random numbers come from Python's generator, not
    ChaCha8Rng::seed_from_u64, so the individual positions and colours are not the reference's."""
    s = SceneDoc(background_color=(0.7, 0.8, 1.0), look_from=(13.0, 2.0, 3.0), look_at=ZERO, field_of_view=20.0,
                 focus_distance=10.0, defocus_angle=0.5, ray_max_bounces=10, samples_per_pixel=10)
    s.merge_camera(camera)
    rng = random.Random(seed)

    def rand3() -> Vec:
        return (rng.random(), rng.random(), rng.random())

    def lamb(color: Vec) -> str:
        m, t = s.ids.next_material(), s.ids.next_texture()   # material id first (spheres.rs:38-39)
        return s.lambertian(s.solid(color, t), m)

    def metal(color: Vec, fuzz: float) -> str:
        m, t = s.ids.next_material(), s.ids.next_texture()
        return s.metal(s.solid(color, t), fuzz, m)
    s.sphere(_mul(100000.0, _neg(Y)), 100000.0, lamb(_mul(0.5, ONE)))
    s.sphere(Y, 1.0, s.dielectric(1.5))
    s.sphere(_sub(Y, _mul(4.0, X)), 1.0, lamb(rand3()))
    s.sphere(_add(Y, _mul(4.0, X)), 1.0, metal(rand3(), rng.random()))
    for a, b in itertools.product(range(-11, 11), range(-11, 11)):
        center = (a + 0.9 * rng.random(), 0.2, b + 0.9 * rng.random())
        pick = rng.choices((0, 1, 2), weights=(5, 80, 15))[0]
        if pick == 0:
            mat = s.dielectric(1.5)
        elif pick == 1:
            mat = lamb(rand3())
        else:
            mat = metal(rand3(), rng.random())
        s.sphere(center, 0.2, mat)
    return s


def read_binary_stl(path: str) -> List[Tuple[Vec, Vec, Vec]]:
    """convert_stl.rs:17-50: 80-byte header, u32 count, then normal + 3 vertices (f32 LE) + u16 per triangle;
    vertices are re-oriented (x, y, z) -> (x, z, -y)."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 84:
        raise ValueError("not a binary STL file (shorter than its header)")
    (count,) = struct.unpack_from("<I", data, 80)
    if len(data) < 84 + 50 * count:
        raise ValueError("truncated binary STL file")
    tris = []
    for i in range(count):
        v = struct.unpack_from("<12f", data, 84 + 50 * i)

        def vert(k):
            return (float(v[3 * k]), float(v[3 * k + 2]), -float(v[3 * k + 1]))
        tris.append((vert(1), vert(2), vert(3)))
    return tris


def convert_stl(path: str, camera=None) -> Tuple[SceneDoc, str]:   # convert_stl.rs:52-138
    tris = read_binary_stl(path)
    inf = math.inf
    lo, hi = [inf] * 3, [-inf] * 3
    for tri in tris:
        for p in tri:
            for k in range(3):
                lo[k], hi[k] = min(lo[k], p[k]), max(hi[k], p[k])
    l, h, w = hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]
    k = 1.0 / max(max(l, w), h)
    look_at = (k * l / 2.0, k * h / 2.0, 0.0)
    s = SceneDoc(background_color=ONE, look_at=look_at, look_from=_add(look_at, Z), field_of_view=50.0,
                 ray_max_bounces=50, samples_per_pixel=200)
    s.merge_camera(camera)
    p_min = (lo[0], lo[1], lo[2])
    objects = [{"Triangle": {"point": _clean(_mul(k, _sub(a, p_min))), "u": _clean(_mul(k, _sub(b, a))),
                             "v": _clean(_mul(k, _sub(c, a)))}} for a, b, c in tris]
    s.scene.append({"Group": {"objects": objects}})
    return s, f"# model bbox: l={k * l:.4f} h={k * h:.4f} w={k * w:.4f}\n"


GENERATORS = {"cornell-box": cornell_box, "cube": cube, "earth": earth, "noise": noise, "quads": quads,
              "triangles": triangles, "spheres": spheres, "simple-lights": simple_lights}


# ------------------------------------------------------------------ writers
def _toml_value(v: Any) -> str:
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, int):
        return str(v)
    if isinstance(v, float):
        if math.isinf(v) or math.isnan(v):
            return ("-" if v < 0 else "") + ("inf" if math.isinf(v) else "nan")
        r = repr(v)
        return r if any(c in r for c in ".en") else r + ".0"
    if isinstance(v, str):
        return json.dumps(v)
    if isinstance(v, (list, tuple)):
        return "[" + ", ".join(_toml_value(x) for x in v) + "]"
    if isinstance(v, dict):
        return "{ " + ", ".join(f"{k} = {_toml_value(x)}" for k, x in v.items() if x is not None) + " }"
    raise TypeError(f"cannot write {type(v).__name__} as TOML")


def dumps(doc: SceneDoc, fmt: str = "toml") -> str:
    d = doc.to_dict()
    if fmt == "json":
        return json.dumps(d, indent=2)
    if fmt != "toml":
        raise ValueError("format must be 'json' or 'toml'")
    out = []
    for key in ("textures", "materials"):     # top-level keys have to precede the first table header
        if key in d:
            out.append(f"{key} = [")
            out += [f"    {_toml_value(pair)}," for pair in d[key]]
            out += ["]", ""]
    out.append("[camera]")
    out += [f"{k} = {_toml_value(v)}" for k, v in d["camera"].items() if v is not None]   # None is skipped in TOML
    out.append("")
    for obj in d.get("scene", []):
        (kind, fields), = obj.items()
        out += ["[[scene]]", f"{kind} = {_toml_value(fields)}", ""]
    return "\n".join(out)


def get_format(fmt: Optional[str], output: Optional[str]) -> str:
    """create.rs:93-112: explicit format, else the output file's extension, else TOML."""
    if fmt:
        return fmt
    ext = os.path.splitext(output or "")[1].lower()
    return {".json": "json", ".toml": "toml"}.get(ext, "toml")


def main(argv: Optional[Sequence[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog="nr-ray-tracer create", description=__doc__.split("\n")[0])
    ap.add_argument("kind", choices=sorted(GENERATORS) + ["convert-stl"])
    ap.add_argument("stl_file", nargs="?", help="convert-stl: binary STL input")
    ap.add_argument("-f", "--force-overwrite", action="store_true")
    ap.add_argument("-F", "--format", choices=["json", "toml"])
    ap.add_argument("-o", "--output", metavar="FILE")
    ap.add_argument("-s", "--seed", type=int, default=1)
    for name in _CAMERA_FIELDS:   # cli.rs:160-226 camera flags
        flag = "--" + name.replace("_", "-")
        if name in ("background_color", "look_at", "look_from", "view_up"):
            ap.add_argument(flag, type=lambda t: tuple(float(x) for x in t.split(",")), metavar="X,Y,Z")
        elif name in ("width", "height", "samples_per_pixel", "ray_max_bounces"):
            ap.add_argument(flag, type=int)
        else:
            ap.add_argument(flag, type=float)
    a = ap.parse_args(argv)
    camera = {k: getattr(a, k) for k in _CAMERA_FIELDS if getattr(a, k) is not None}
    header = ""
    if a.kind == "convert-stl":
        if not a.stl_file:
            ap.error("convert-stl needs an STL file")
        doc, header = convert_stl(a.stl_file, camera)
        fmt = a.format or "toml"
        if fmt == "json":
            header = ""   # the reference writes the '#' line in front of JSON too, which no JSON parser accepts
    else:
        doc = GENERATORS[a.kind](camera, a.seed) if a.kind == "spheres" else GENERATORS[a.kind](camera)
        fmt = get_format(a.format, a.output)
    text = header + dumps(doc, fmt) + "\n"
    if a.output:
        with open(a.output, "w" if a.force_overwrite else "x") as f:   # create_new unless -f (create.rs:26-35)
            f.write(text)
    else:
        sys.stdout.write(text)
    return 0


if __name__ == "__main__":
    sys.exit(main())
