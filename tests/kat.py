"""Known-answer-test helpers: deterministic ray sets and hit comparison (SURVEY.md §4 level 1)."""
from __future__ import annotations

import numpy as np

from nr_ray_tracer_b200 import _abi as A
from nr_ray_tracer_b200.scene_config import SceneGraph


def scene_bounds(graph: SceneGraph):
    """Loose bounds of the primitives' defining points (ignores wrappers; good enough to aim rays)."""
    pts = []
    for kind, _m, _c, v in graph.objects:
        if kind == A.OBJ_SPHERE:
            c, r = np.array(v[:3]), min(abs(v[3]), 50.0)
            pts += [c - r, c + r] if abs(v[3]) < 1e4 else [c + np.array([0, v[3], 0]) + np.array([-20, -1, -20]),
                                                           c + np.array([0, v[3], 0]) + np.array([20, 1, 20])]
        elif kind in (A.OBJ_QUAD, A.OBJ_TRIANGLE):
            p, u, w = np.array(v[0:3]), np.array(v[3:6]), np.array(v[6:9])
            pts += [p, p + u, p + w, p + u + w]
    pts = np.array(pts)
    return pts.min(axis=0), pts.max(axis=0)


def random_rays(graph: SceneGraph, n: int, seed: int = 1234) -> np.ndarray:
    """Origins uniform in the scene bounds inflated x2, directions uniform on the sphere (not normalised
    to unit length: scaled by a random factor, since the reference never normalises directions)."""
    rng = np.random.default_rng(seed)
    lo, hi = scene_bounds(graph)
    c, h = (lo + hi) / 2, (hi - lo) / 2 + 1e-3
    o = c + (rng.random((n, 3)) * 2 - 1) * h * 2.0
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= np.exp(rng.uniform(-2, 2, size=(n, 1)))
    return np.concatenate([o, d], axis=1)


def aimed_rays(graph: SceneGraph, n: int, seed: int = 99) -> np.ndarray:
    """Rays from outside towards random points inside the bounds (high hit rate)."""
    rng = np.random.default_rng(seed)
    lo, hi = scene_bounds(graph)
    c, h = (lo + hi) / 2, (hi - lo) / 2 + 1e-3
    target = c + (rng.random((n, 3)) * 2 - 1) * h
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    o = target - dirs * (np.linalg.norm(h) * rng.uniform(0.1, 3.0, size=(n, 1)))
    return np.concatenate([o, target - o], axis=1)


def special_rays(graph: SceneGraph, seed: int = 7) -> np.ndarray:
    """Axis-parallel rays (zero direction components -> +-inf / NaN in the slab test), rays through
    primitive corners / edges / bbox corners, and rays starting exactly on surfaces."""
    rng = np.random.default_rng(seed)
    lo, hi = scene_bounds(graph)
    c = (lo + hi) / 2
    rays = []
    axes = np.eye(3)
    for a in range(3):
        for s in (-1.0, 1.0):
            for _ in range(64):
                o = lo + rng.random(3) * (hi - lo)
                o[a] = (lo[a] - 1.0) if s > 0 else (hi[a] + 1.0)
                rays.append(np.concatenate([o, axes[a] * s]))
                d2 = axes[a] * s
                d2[(a + 1) % 3] = rng.normal() * 0.3  # one zero component
                rays.append(np.concatenate([o, d2]))
    targets = []
    for kind, _m, _c, v in graph.objects:
        if kind in (A.OBJ_QUAD, A.OBJ_TRIANGLE):
            p, u, w = np.array(v[0:3]), np.array(v[3:6]), np.array(v[6:9])
            targets += [p, p + u, p + w, p + u + w, p + 0.5 * u, p + 0.5 * w, p + 0.5 * u + 0.5 * w,
                        p + 0.25 * u + 0.25 * w]
        elif kind == A.OBJ_SPHERE and abs(v[3]) < 1e4:
            cc, r = np.array(v[:3]), v[3]
            for a in range(3):
                targets += [cc + axes[a] * r, cc - axes[a] * r]
        if len(targets) > 4000:
            break
    for t in targets:
        for _ in range(2):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            o = t - d * rng.uniform(0.5, 5.0)
            rays.append(np.concatenate([o, t - o]))      # aimed exactly at a corner / edge / pole
            rays.append(np.concatenate([t, d]))          # starting exactly on it
    rays.append(np.concatenate([c, [0.0, 0.0, 0.0]]))    # zero direction: NaNs everywhere
    return np.array(rays)


def ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units in the last place between two f64 arrays (finite values)."""
    ia = np.ascontiguousarray(a, dtype=np.float64).view(np.int64).copy()
    ib = np.ascontiguousarray(b, dtype=np.float64).view(np.int64).copy()
    ia = np.where(ia < 0, np.int64(-2**63) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2**63) - ib, ib)
    return np.abs(ia - ib)


def compare_hits(gpu: np.ndarray, ref: np.ndarray, max_t_ulp: int = 2, vec_tol: float = 1e-9):
    """Returns a dict of mismatch counts.  hit/miss, object id, material, front_face must be identical;
    t within max_t_ulp ulp; point / normal / uv within vec_tol (absolute, relative to magnitude)."""
    hit_g, hit_r = gpu["object"] != 0xFFFFFFFF, ref["object"] != 0xFFFFFFFF
    res = {"n": len(gpu), "hits": int(hit_r.sum())}
    res["hitmiss_mismatch"] = int((hit_g != hit_r).sum())
    both = hit_g & hit_r
    res["object_mismatch"] = int((gpu["object"][both] != ref["object"][both]).sum())
    same = both & (gpu["object"] == ref["object"])
    res["material_mismatch"] = int((gpu["material"][same] != ref["material"][same]).sum())
    res["front_face_mismatch"] = int((gpu["front_face"][same] != ref["front_face"][same]).sum())
    ud = ulp_diff(gpu["t"][same], ref["t"][same])
    res["t_max_ulp"] = int(ud.max()) if ud.size else 0
    res["t_over_ulp"] = int((ud > max_t_ulp).sum())

    def vec_bad(name):
        g, r = gpu[name][same], ref[name][same]
        scale = np.maximum(1.0, np.abs(r))
        return int((np.abs(g - r) > vec_tol * scale).any(axis=1).sum())
    res["point_bad"] = vec_bad("point")
    res["normal_bad"] = vec_bad("normal")
    res["uv_bad"] = vec_bad("uv")
    return res


def hits_ok(res: dict) -> bool:
    return all(res[k] == 0 for k in ("hitmiss_mismatch", "object_mismatch", "material_mismatch",
                                     "front_face_mismatch", "t_over_ulp", "point_bad", "normal_bad", "uv_bad"))
