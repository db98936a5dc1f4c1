#!/usr/bin/env python3
"""Synthesises scenes/utah-teapot-model.toml, which is MISSING from the reference
(scenes/utah-teapot-scene.json:41 references it; upstream's .gitignore hid it).

STAND-IN GEOMETRY: parity with upstream's own model file is unpinned.  The mesh is the classic
Newell teapot (the 10 bicubic Bezier patches + reflections of the public-domain 1975 data set, as
distributed with GLUT's teapot.c), tessellated N x N per patch, then pushed through exactly the
normalisation the reference's STL converter applies
(packages/ray-tracer/src/commands/create/convert_stl.rs:40-109):
    (x, y, z) -> (x, z, -y);  k = 1 / longest bbox extent;
    point = k (a - p_min), u = k (b - a), v = k (c - a);  one Group of Triangles, material = None.
The result has bbox ~ 1.0 x 0.483 x 0.613, consistent with the scene file's
Translate(-0.5, -0.244, -0.311) that centres upstream's model (1.0 x 0.488 x 0.622).

Usage: python tools/make_teapot.py [--n 10] [--out scenes/utah-teapot-model.toml]
"""
import argparse

import numpy as np

PATCHES = [
    # rim
    [102, 103, 104, 105, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15],
    # body
    [12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27],
    [24, 25, 26, 27, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40],
    # lid
    [96, 96, 96, 96, 97, 98, 99, 100, 101, 101, 101, 101, 0, 1, 2, 3],
    [0, 1, 2, 3, 106, 107, 108, 109, 110, 111, 112, 113, 114, 115, 116, 117],
    # bottom
    [118, 118, 118, 118, 124, 122, 119, 121, 123, 126, 125, 120, 40, 39, 38, 37],
    # handle
    [41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56],
    [53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 28, 65, 66, 67],
    # spout
    [68, 69, 70, 71, 72, 73, 74, 75, 76, 77, 78, 79, 80, 81, 82, 83],
    [80, 81, 82, 83, 84, 85, 86, 87, 88, 89, 90, 91, 92, 93, 94, 95],
]
FOUR_FOLD = [0, 1, 2, 3, 4, 5]  # rim, body, lid, bottom: reflect in x and y
TWO_FOLD = [6, 7, 8, 9]         # handle, spout: reflect in y only

CP = [
    (0.2, 0, 2.7), (0.2, -0.112, 2.7), (0.112, -0.2, 2.7), (0, -0.2, 2.7), (1.3375, 0, 2.53125),
    (1.3375, -0.749, 2.53125), (0.749, -1.3375, 2.53125), (0, -1.3375, 2.53125), (1.4375, 0, 2.53125),
    (1.4375, -0.805, 2.53125), (0.805, -1.4375, 2.53125), (0, -1.4375, 2.53125), (1.5, 0, 2.4), (1.5, -0.84, 2.4),
    (0.84, -1.5, 2.4), (0, -1.5, 2.4), (1.75, 0, 1.875), (1.75, -0.98, 1.875), (0.98, -1.75, 1.875),
    (0, -1.75, 1.875), (2, 0, 1.35), (2, -1.12, 1.35), (1.12, -2, 1.35), (0, -2, 1.35), (2, 0, 0.9),
    (2, -1.12, 0.9), (1.12, -2, 0.9), (0, -2, 0.9), (-2, 0, 0.9), (2, 0, 0.45), (2, -1.12, 0.45),
    (1.12, -2, 0.45), (0, -2, 0.45), (1.5, 0, 0.225), (1.5, -0.84, 0.225), (0.84, -1.5, 0.225), (0, -1.5, 0.225),
    (1.5, 0, 0.15), (1.5, -0.84, 0.15), (0.84, -1.5, 0.15), (0, -1.5, 0.15), (-1.6, 0, 2.025), (-1.6, -0.3, 2.025),
    (-1.5, -0.3, 2.25), (-1.5, 0, 2.25), (-2.3, 0, 2.025), (-2.3, -0.3, 2.025), (-2.5, -0.3, 2.25), (-2.5, 0, 2.25),
    (-2.7, 0, 2.025), (-2.7, -0.3, 2.025), (-3, -0.3, 2.25), (-3, 0, 2.25), (-2.7, 0, 1.8), (-2.7, -0.3, 1.8),
    (-3, -0.3, 1.8), (-3, 0, 1.8), (-2.7, 0, 1.575), (-2.7, -0.3, 1.575), (-3, -0.3, 1.35), (-3, 0, 1.35),
    (-2.5, 0, 1.125), (-2.5, -0.3, 1.125), (-2.65, -0.3, 0.9375), (-2.65, 0, 0.9375), (-2, -0.3, 0.9),
    (-1.9, -0.3, 0.6), (-1.9, 0, 0.6), (1.7, 0, 1.425), (1.7, -0.66, 1.425), (1.7, -0.66, 0.6), (1.7, 0, 0.6),
    (2.6, 0, 1.425), (2.6, -0.66, 1.425), (3.1, -0.66, 0.825), (3.1, 0, 0.825), (2.3, 0, 2.1), (2.3, -0.25, 2.1),
    (2.4, -0.25, 2.025), (2.4, 0, 2.025), (2.7, 0, 2.4), (2.7, -0.25, 2.4), (3.3, -0.25, 2.4), (3.3, 0, 2.4),
    (2.8, 0, 2.475), (2.8, -0.25, 2.475), (3.525, -0.25, 2.49375), (3.525, 0, 2.49375), (2.9, 0, 2.475),
    (2.9, -0.15, 2.475), (3.45, -0.15, 2.5125), (3.45, 0, 2.5125), (2.8, 0, 2.4), (2.8, -0.15, 2.4),
    (3.2, -0.15, 2.4), (3.2, 0, 2.4), (0, 0, 3.15), (0.8, 0, 3.15), (0.8, -0.45, 3.15), (0.45, -0.8, 3.15),
    (0, -0.8, 3.15), (0, 0, 2.85), (1.4, 0, 2.4), (1.4, -0.784, 2.4), (0.784, -1.4, 2.4), (0, -1.4, 2.4),
    (0.4, 0, 2.55), (0.4, -0.224, 2.55), (0.224, -0.4, 2.55), (0, -0.4, 2.55), (1.3, 0, 2.55), (1.3, -0.728, 2.55),
    (0.728, -1.3, 2.55), (0, -1.3, 2.55), (1.3, 0, 2.4), (1.3, -0.728, 2.4), (0.728, -1.3, 2.4), (0, -1.3, 2.4),
    (0, 0, 0), (1.425, -0.798, 0), (1.5, 0, 0.075), (1.425, 0, 0), (0.798, -1.425, 0), (0, -1.5, 0.075),
    (0, -1.425, 0), (1.5, -0.84, 0.075), (0.84, -1.5, 0.075),
]


def bernstein(t):
    return np.array([(1 - t) ** 3, 3 * t * (1 - t) ** 2, 3 * t * t * (1 - t), t ** 3])


def tessellate(n):
    cp = np.array(CP, dtype=np.float64)
    assert cp.shape == (127, 3)
    ts = np.linspace(0.0, 1.0, n + 1)
    B = np.stack([bernstein(t) for t in ts])  # (n+1, 4)
    tris = []
    for pi, patch in enumerate(PATCHES):
        ctrl = cp[np.array(patch)].reshape(4, 4, 3)
        mirrors = [(1, 1), (1, -1)] if pi in TWO_FOLD else [(1, 1), (-1, 1), (1, -1), (-1, -1)]
        for sx, sy in mirrors:
            c = ctrl * np.array([sx, sy, 1.0])
            grid = np.einsum("ui,vj,ijk->uvk", B, B, c)  # (n+1, n+1, 3)
            flip = (sx * sy) < 0  # keep a consistent winding under reflection
            for i in range(n):
                for j in range(n):
                    p00, p10, p01, p11 = grid[i, j], grid[i + 1, j], grid[i, j + 1], grid[i + 1, j + 1]
                    for a, b, cc in ((p00, p10, p11), (p00, p11, p01)):
                        if flip:
                            b, cc = cc, b
                        area2 = np.linalg.norm(np.cross(b - a, cc - a))
                        if area2 > 1e-12:  # collapsed patch edges (lid tip, bottom centre) give zero-area triangles
                            tris.append((a, b, cc))
    return tris


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10, help="tessellation steps per patch edge")
    ap.add_argument("--out", default="scenes/utah-teapot-model.toml")
    args = ap.parse_args()
    tris = tessellate(args.n)
    # convert_stl.rs:40-46: STL (x,y,z) -> (x, z, -y)
    conv = lambda p: np.array([p[0], p[2], -p[1]])
    tris = [(conv(a), conv(b), conv(c)) for a, b, c in tris]
    allp = np.array([p for t in tris for p in t])
    p_min, p_max = allp.min(axis=0), allp.max(axis=0)
    ext = p_max - p_min
    k = 1.0 / max(ext[0], ext[2], ext[1])
    with open(args.out, "w") as f:
        f.write(f"# model bbox: l={k * ext[0]:.4f} h={k * ext[1]:.4f} w={k * ext[2]:.4f}\n")
        f.write("# STAND-IN for upstream's missing utah-teapot-model.toml; generated by tools/make_teapot.py "
                f"--n {args.n} ({len(tris)} triangles)\n")
        f.write("[camera]\nbackground_color = [1.0, 1.0, 1.0]\n")
        f.write(f"look_at = [{float(k * ext[0] / 2.0)!r}, {float(k * ext[1] / 2.0)!r}, 0.0]\n")
        f.write(f"look_from = [{float(k * ext[0] / 2.0)!r}, {float(k * ext[1] / 2.0)!r}, 1.0]\n")
        f.write("field_of_view = 50.0\nsamples_per_pixel = 200\nray_max_bounces = 50\n\n")
        f.write("[[scene]]\n[scene.Group]\nobjects = [\n")
        for a, b, c in tris:
            pt, u, v = k * (a - p_min), k * (b - a), k * (c - a)
            fmt = lambda x: "[" + ", ".join(repr(float(t)) for t in x) + "]"
            f.write(f"  {{ Triangle = {{ point = {fmt(pt)}, u = {fmt(u)}, v = {fmt(v)} }} }},\n")
        f.write("]\n")
    print(f"wrote {args.out}: {len(tris)} triangles, bbox {k * ext}")


if __name__ == "__main__":
    main()
