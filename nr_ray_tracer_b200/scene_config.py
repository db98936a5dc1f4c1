"""Host-side scene loader: TOML/JSON scene file -> object-graph description.

Mirrors the reference CLI's loader so the hot path gets the same input:
  SceneConfig / TextureConfig / MaterialConfig / ObjectConfig   packages/ray-tracer/src/scene_config.rs:27-404
  try_build_aux (textures -> materials -> instances -> objects)  scene_config.rs:411-473
  CameraConfig get_size / merge_with / try_update               packages/ray-tracer/src/cli.rs:272-402

It is tolerant where the reference's current serde schema is not (SURVEY.md note B):
`textures` / `materials` may be the v3 array of [id, {Kind: {...}}] pairs, the v2
table keyed by id (scenes/spheres.toml, earth.toml, noise.toml materials), or the
v1 anonymous array addressed by integer index with objects under `objects`
(scenes/triangles.toml).  Paths inside scene files resolve against `base_dir`
(the reference uses the process CWD, scene_config.rs:88,337).

The output is plain data (nrrt_graph_desc, include/nrrt.h); BVH construction and
flattening happen in the C++ host library (csrc/host_scene.cpp).
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import tomllib
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import _abi as A


class SceneError(Exception):
    pass


# --------------------------------------------------------------------------- camera
_CAMERA_KEYS = ("width", "height", "aspect_ratio", "background_color", "look_at", "look_from", "view_up",
                "focal_length", "field_of_view", "defocus_angle", "focus_distance", "samples_per_pixel",
                "ray_max_bounces")


@dataclass
class CameraConfig:
    """cli.rs:157-270.  fov / defocus angle in DEGREES here, like the CLI and scene files."""
    width: Optional[int] = None
    height: Optional[int] = None
    aspect_ratio: Optional[float] = None
    background_color: Optional[Tuple[float, float, float]] = None
    look_at: Optional[Tuple[float, float, float]] = None
    look_from: Optional[Tuple[float, float, float]] = None
    view_up: Optional[Tuple[float, float, float]] = None
    focal_length: Optional[float] = None  # parsed, never used (cli.rs:229)
    field_of_view: Optional[float] = None
    defocus_angle: Optional[float] = None
    focus_distance: Optional[float] = None
    samples_per_pixel: Optional[int] = None
    ray_max_bounces: Optional[int] = None

    @classmethod
    def from_dict(cls, d: Optional[Dict[str, Any]]) -> "CameraConfig":
        d = d or {}
        kw = {}
        for k in _CAMERA_KEYS:
            v = d.get(k)
            if v is None:
                continue
            if k in ("background_color", "look_at", "look_from", "view_up"):
                v = tuple(float(x) for x in v)
            kw[k] = v
        return cls(**kw)

    def merge_with(self, other: "CameraConfig") -> "CameraConfig":
        """cli.rs:316-355: every Some in `other` replaces (focal_length is not merged)."""
        for k in _CAMERA_KEYS:
            if k == "focal_length":
                continue
            v = getattr(other, k)
            if v is not None:
                setattr(self, k, v)
        return self

    def get_size(self) -> Optional[Tuple[int, int]]:
        """cli.rs:273-312: exactly two of width/height/aspect_ratio, or none."""
        w, h, r = self.width, self.height, self.aspect_ratio
        if w is None and h is None and r is None:
            return None
        if w is not None and h is not None and r is None:
            return (int(w), int(h))
        if w is not None and h is None and r is not None:
            return (int(w), max(int(float(w) / r), 1))  # image.rs:26-32
        if w is None and h is not None and r is not None:
            return (max(int(float(h) * r), 1), int(h))  # image.rs:34-40
        if w is not None and h is not None and r is not None:
            raise SceneError("conflicting image size arguments")
        raise SceneError("image size needs two of width/height/aspect ratio")

    def to_builder_config(self) -> A.CameraConfig:
        """try_update (cli.rs:357-402) applied to CameraBuilder::default() (camera.rs:162-203)."""
        c = A.CameraConfig()
        size = self.get_size() or (1200, 800)
        c.width, c.height = size
        bg = self.background_color or (0.0, 0.0, 0.0)
        lf = self.look_from or (1.0, 1.0, 1.0)
        la = self.look_at or (0.0, 0.0, 0.0)
        vu = self.view_up or (0.0, 1.0, 0.0)
        for i in range(3):
            c.background[i], c.look_from[i], c.look_at[i], c.view_up[i] = bg[i], lf[i], la[i], vu[i]
        c.field_of_view = (self.field_of_view * math.pi) / 180.0 if self.field_of_view is not None else math.pi / 2.0
        c.focus_dist = self.focus_distance if self.focus_distance is not None else 1.0
        c.defocus_angle = (self.defocus_angle * math.pi) / 180.0 if self.defocus_angle is not None else 0.0
        c.samples_per_pixel = self.samples_per_pixel if self.samples_per_pixel is not None else 10
        c.ray_max_bounces = self.ray_max_bounces if self.ray_max_bounces is not None else 10
        return c


# --------------------------------------------------------------------------- graph
@dataclass
class SceneGraph:
    objects: List[Tuple[int, int, List[int], Tuple[float, ...]]] = field(default_factory=list)  # kind, mat, children, v
    materials: List[Tuple[int, int, float]] = field(default_factory=list)
    textures: List[Dict[str, Any]] = field(default_factory=list)
    images: List[np.ndarray] = field(default_factory=list)  # (H, W, 3) uint8
    image_paths: Dict[str, int] = field(default_factory=dict)
    root: int = 0
    camera: CameraConfig = field(default_factory=CameraConfig)

    # -- builders
    def add_texture(self, **kw) -> int:
        self.textures.append(kw)
        return len(self.textures) - 1

    def add_material(self, kind: int, texture: int, param: float = 0.0) -> int:
        self.materials.append((kind, texture, float(param)))
        return len(self.materials) - 1

    def add_object(self, kind: int, material: int = 0, children: Optional[List[int]] = None, v=()) -> int:
        v = tuple(float(x) for x in v) + (0.0,) * (9 - len(v))
        self.objects.append((kind, material, list(children or []), v))
        return len(self.objects) - 1

    def add_image(self, rgb: np.ndarray, key: Optional[str] = None) -> int:
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        if rgb.ndim != 3 or rgb.shape[2] != 3:
            raise SceneError("image must be (H, W, 3) uint8")
        self.images.append(rgb)
        if key is not None:
            self.image_paths[key] = len(self.images) - 1
        return len(self.images) - 1

    def count_primitives(self) -> int:
        return sum(1 for o in self.objects if o[0] in (A.OBJ_SPHERE, A.OBJ_QUAD, A.OBJ_TRIANGLE))

    # -- ctypes view (keeps the backing arrays alive on the returned object)
    def to_desc(self) -> "GraphDescHolder":
        return GraphDescHolder(self)


class GraphDescHolder:
    def __init__(self, g: SceneGraph):
        n = len(g.objects)
        self.objects = (A.Object * max(n, 1))()
        child_ids: List[int] = []
        for i, (kind, mat, children, v) in enumerate(g.objects):
            o = self.objects[i]
            o.kind, o.material = kind, mat
            o.first_child, o.n_children = len(child_ids), len(children)
            child_ids.extend(children)
            for k in range(9):
                o.v[k] = v[k]
        self.child_ids = (A.u32 * max(len(child_ids), 1))(*child_ids)
        self.materials = (A.Material * max(len(g.materials), 1))()
        for i, (kind, tex, param) in enumerate(g.materials):
            self.materials[i].kind, self.materials[i].texture, self.materials[i].param = kind, tex, param
        self.textures = (A.Texture * max(len(g.textures), 1))()
        for i, t in enumerate(g.textures):
            d = self.textures[i]
            d.kind = t["kind"]
            d.a, d.b = t.get("a", 0), t.get("b", 0)
            d.seed, d.octaves = t.get("seed", 0), t.get("octaves", 0)
            col = t.get("color", (0.0, 0.0, 0.0))
            for k in range(3):
                d.color[k] = float(col[k])
            d.f0, d.f1, d.f2 = float(t.get("f0", 0.0)), float(t.get("f1", 0.0)), float(t.get("f2", 0.0))
        self._image_arrays = g.images
        self.images = (A.Image * max(len(g.images), 1))()
        for i, im in enumerate(g.images):
            self.images[i].height, self.images[i].width = im.shape[0], im.shape[1]
            self.images[i].rgb = im.ctypes.data
        d = A.GraphDesc()
        d.n_objects, d.objects = n, self.objects
        d.n_child_ids, d.child_ids = len(child_ids), self.child_ids
        d.n_materials, d.materials = len(g.materials), self.materials
        d.n_textures, d.textures = len(g.textures), self.textures
        d.n_images, d.images = len(g.images), self.images
        d.root = g.root
        self.desc = d

    def ptr(self):
        return C.byref(self.desc)


# --------------------------------------------------------------------------- loader
NOISE_DEFAULT_FREQUENCY = 1.0             # noise 0.9.0 Fbm::DEFAULT_FREQUENCY
NOISE_DEFAULT_LACUNARITY = math.pi * 2.0 / 3.0  # Fbm::DEFAULT_LACUNARITY
NOISE_DEFAULT_PERSISTENCE = 0.5           # Fbm::DEFAULT_PERSISTENCE


def _single_variant(cfg: Any, what: str) -> Tuple[str, Dict[str, Any]]:
    if not isinstance(cfg, dict) or len(cfg) != 1:
        raise SceneError(f"{what}: expected a single-variant table, got {cfg!r}")
    (k, v), = cfg.items()
    return k, (v or {})


def _vec3(v: Any, what: str) -> Tuple[float, float, float]:
    if not isinstance(v, (list, tuple)) or len(v) != 3:
        raise SceneError(f"{what}: expected a 3-vector")
    return (float(v[0]), float(v[1]), float(v[2]))


def _load_image(path: str) -> np.ndarray:
    from PIL import Image  # decode only; the reference uses image 0.25.8 / zune-jpeg (SURVEY note A)
    with Image.open(path) as im:
        return np.asarray(im.convert("RGB"), dtype=np.uint8)


def load_scene_file(path: str) -> Dict[str, Any]:
    """SceneConfig::try_load_scene (scene_config.rs:475-492): format chosen by extension."""
    ext = os.path.splitext(path)[1].lower()
    try:
        if ext == ".json":
            with open(path, "r", encoding="utf-8") as f:
                return json.load(f)
        if ext == ".toml":
            with open(path, "rb") as f:
                return tomllib.load(f)
    except OSError as e:
        raise SceneError(f"cannot read scene file {path}: {e}") from e
    except (json.JSONDecodeError, tomllib.TOMLDecodeError) as e:
        raise SceneError(f"cannot parse scene file {path}: {e}") from e
    raise SceneError("invalid scene file format!")


class _Builder:
    def __init__(self, graph: SceneGraph, base_dir: str):
        self.g = graph
        self.base_dir = base_dir

    def resolve(self, p: str) -> str:
        return p if os.path.isabs(p) else os.path.join(self.base_dir, p)

    # TextureConfig::try_make_texture (scene_config.rs:53-122)
    def make_texture(self, cfg: Any, textures: Dict[Any, int]) -> int:
        kind, p = _single_variant(cfg, "texture")
        g = self.g
        if kind == "SolidColor":
            return g.add_texture(kind=A.TEX_SOLID, color=_vec3(p.get("color"), "SolidColor.color"))
        if kind == "Checker":
            def ref(name, default_color):
                tid = p.get(name)
                if tid is None:  # CheckerBuilder defaults: even = ONE, odd = ZERO (checker.rs:53-54)
                    return g.add_texture(kind=A.TEX_SOLID, color=default_color)
                if tid not in textures:
                    raise SceneError("invalid texture index")
                return textures[tid]
            even = ref("even", (1.0, 1.0, 1.0))
            odd = ref("odd", (0.0, 0.0, 0.0))
            scale = p.get("scale")
            return g.add_texture(kind=A.TEX_CHECKER, a=even, b=odd, f0=0.5 if scale is None else float(scale))
        if kind == "Image":
            path = self.resolve(str(p.get("path")))
            key = os.path.abspath(path)
            if key in g.image_paths:
                idx = g.image_paths[key]
            else:
                try:
                    idx = g.add_image(_load_image(path), key)
                except OSError as e:
                    raise SceneError(f"cannot load image {path}: {e}") from e
            return g.add_texture(kind=A.TEX_IMAGE, a=idx)
        if kind == "Marble":
            seed = p.get("seed") or 0
            freq = p.get("frequency")
            return g.add_texture(kind=A.TEX_MARBLE, seed=int(seed), octaves=7,
                                 f0=NOISE_DEFAULT_FREQUENCY if freq is None else float(freq))
        if kind == "Noise":
            def opt(name, default):
                v = p.get(name)
                return default if v is None else v
            return g.add_texture(kind=A.TEX_NOISE, seed=int(opt("seed", 0)), octaves=int(opt("octaves", 1)),
                                 f0=float(opt("frequency", NOISE_DEFAULT_FREQUENCY)),
                                 f1=float(opt("lacunarity", NOISE_DEFAULT_LACUNARITY)),
                                 f2=float(opt("persistence", NOISE_DEFAULT_PERSISTENCE)))
        raise SceneError(f"unknown texture kind {kind!r}")

    # MaterialConfig::try_make_material (scene_config.rs:163-197)
    def make_material(self, cfg: Any, textures: Dict[Any, int], texture_fallback: int) -> int:
        kind, p = _single_variant(cfg, "material")

        def tex():
            tid = p.get("texture")
            if tid is None:
                return texture_fallback
            if tid not in textures:
                raise SceneError(f"invalid texture id: '{tid}'")
            return textures[tid]
        if kind == "Dielectric":
            return self.g.add_material(A.MAT_DIELECTRIC, 0, float(p["refraction_index"]))
        if kind == "DiffuseLight":
            return self.g.add_material(A.MAT_DIFFUSE_LIGHT, tex(), float(p["intensity"]))
        if kind == "Lambertian":
            return self.g.add_material(A.MAT_LAMBERTIAN, tex(), 0.0)
        if kind == "Metal":
            return self.g.add_material(A.MAT_METAL, tex(), float(p["fuzz"]))
        raise SceneError(f"unknown material kind {kind!r}")

    # ObjectConfig::try_make_object (scene_config.rs:278-380)
    def make_object(self, cfg: Any, instances: Dict[Any, int], materials: Dict[Any, int], material_fallback: int,
                    depth: int = 0) -> int:
        if depth > 64:
            raise SceneError("object nesting too deep")
        kind, p = _single_variant(cfg, "object")
        g = self.g

        def mat():
            mid = p.get("material")
            if mid is None:
                return material_fallback
            if mid not in materials:
                raise SceneError(f"invalid material id: '{mid}'")
            return materials[mid]

        def inner():
            return self.make_object(p.get("object"), instances, materials, material_fallback, depth + 1)

        if kind in ("Quad", "Triangle"):
            v = _vec3(p.get("point"), "point") + _vec3(p.get("u"), "u") + _vec3(p.get("v"), "v")
            return g.add_object(A.OBJ_QUAD if kind == "Quad" else A.OBJ_TRIANGLE, mat(), v=v)
        if kind == "Sphere":
            return g.add_object(A.OBJ_SPHERE, mat(), v=_vec3(p.get("center"), "center") + (float(p["radius"]),))
        if kind == "Group":
            m = mat()
            kids = [self.make_object(o, instances, materials, m, depth + 1) for o in p.get("objects", [])]
            return g.add_object(A.OBJ_GROUP, children=kids)
        if kind == "Scene":
            m = mat()
            sub_path = self.resolve(str(p.get("path")))
            stack = self.__dict__.setdefault("_include_stack", [])
            if sub_path in stack:
                raise SceneError(f"recursive scene include: {sub_path}")
            if len(stack) >= 16:
                raise SceneError(f"scene includes nested deeper than 16 files: {sub_path}")
            sub = load_scene_file(sub_path)
            stack.append(sub_path)
            try:
                return self.build_aux(sub, m)  # scene.objects: a BVH (= GROUP)
            finally:
                stack.pop()
        if kind == "Ref":
            rid = p.get("id")
            if rid not in instances:
                raise SceneError("invalid object id")
            return instances[rid]
        if kind in ("RotateX", "RotateY", "RotateZ"):
            k = {"RotateX": A.OBJ_ROTATE_X, "RotateY": A.OBJ_ROTATE_Y, "RotateZ": A.OBJ_ROTATE_Z}[kind]
            c = inner()
            return g.add_object(k, children=[c], v=(float(p["angle"]),))
        if kind == "ScaleU":
            c = inner()
            f = float(p["factor"])
            return g.add_object(A.OBJ_SCALE, children=[c], v=(f * 1.0, f * 1.0, f * 1.0))  # factor*DVec3::ONE
        if kind == "ScaleV":
            c = inner()
            return g.add_object(A.OBJ_SCALE, children=[c], v=_vec3(p.get("scale"), "scale"))
        if kind == "Translate":
            c = inner()
            return g.add_object(A.OBJ_TRANSLATE, children=[c], v=_vec3(p.get("offset"), "offset"))
        raise SceneError(f"unknown object kind {kind!r}")

    @staticmethod
    def _pairs(section: Any, what: str) -> List[Tuple[Any, Any]]:
        """Accept v3 [[id, cfg], ...], v2 {id: cfg}, v1 [cfg, ...] (ids = integer positions)."""
        if section is None:
            return []
        if isinstance(section, dict):
            return list(section.items())
        if isinstance(section, list):
            out = []
            for i, e in enumerate(section):
                if isinstance(e, (list, tuple)) and len(e) == 2 and isinstance(e[0], str):
                    out.append((e[0], e[1]))
                elif isinstance(e, dict):
                    out.append((i, e))
                else:
                    raise SceneError(f"{what}: malformed entry {e!r}")
            return out
        raise SceneError(f"{what}: expected an array or table")

    # SceneConfig::try_build_aux (scene_config.rs:411-473); returns the GROUP object index of scene.objects
    def build_aux(self, cfg: Dict[str, Any], material_fallback: Optional[int]) -> int:
        g = self.g
        textures: Dict[Any, int] = {}
        tex_pairs = self._pairs(cfg.get("textures"), "textures")
        if isinstance(cfg.get("textures"), dict):
            # table form has no order: build non-Checker first, then Checkers (which look up earlier ids)
            tex_pairs.sort(key=lambda kv: 1 if (isinstance(kv[1], dict) and "Checker" in kv[1]) else 0)
        for tid, tcfg in tex_pairs:
            textures[tid] = self.make_texture(tcfg, textures)
        if cfg.get("texture_fallback") is not None:
            texture_fallback = self.make_texture(cfg["texture_fallback"], textures)
        else:
            texture_fallback = g.add_texture(kind=A.TEX_SOLID, color=(0.5 * 1.0, 0.5 * 1.0, 0.5 * 1.0))
        materials: Dict[Any, int] = {}
        for mid, mcfg in self._pairs(cfg.get("materials"), "materials"):
            materials[mid] = self.make_material(mcfg, textures, texture_fallback)
        if material_fallback is None:
            if cfg.get("material_fallback") is not None:
                material_fallback = self.make_material(cfg["material_fallback"], textures, texture_fallback)
            else:
                material_fallback = g.add_material(A.MAT_LAMBERTIAN, texture_fallback, 0.0)
        instances: Dict[Any, int] = {}
        for iid, icfg in self._pairs(cfg.get("instances"), "instances"):
            instances[iid] = self.make_object(icfg, instances, materials, material_fallback)
        scene_list = cfg.get("scene")
        if scene_list is None:
            scene_list = cfg.get("objects", [])  # v1 schema
        objs = [self.make_object(o, instances, materials, material_fallback) for o in scene_list]
        return g.add_object(A.OBJ_GROUP, children=objs)


def build_scene_graph(cfg: Dict[str, Any], base_dir: str = ".", camera_override: Optional[CameraConfig] = None
                      ) -> SceneGraph:
    """SceneConfig::try_build with the CLI's camera merge (render.rs:107-111)."""
    g = SceneGraph()
    b = _Builder(g, base_dir)
    g.root = b.build_aux(cfg, None)
    g.camera = CameraConfig.from_dict(cfg.get("camera"))
    if camera_override is not None:
        g.camera.merge_with(camera_override)
    return g


def load_scene(path: str, base_dir: Optional[str] = None, camera_override: Optional[CameraConfig] = None
               ) -> SceneGraph:
    """Load a scene file; relative paths inside it resolve against base_dir (default: CWD, like the reference)."""
    cfg = load_scene_file(path)
    return build_scene_graph(cfg, base_dir if base_dir is not None else os.getcwd(), camera_override)
