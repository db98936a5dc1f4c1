"""Pure-Python interpreter of the flat device layout (include/nrrt.h) with the REFERENCE's traversal
semantics (visit both children, exact f64 slab test, one primitive per leaf, ties to the later leaf).

Test infrastructure: it lets the CPU-only suite check the product's host flattening (csrc/host_scene.cpp)
against the oracle without a GPU.  Python floats are IEEE doubles and never fuse multiply-add, so the
arithmetic below is bit-faithful to the reference's operation order.  Slow: use a few thousand rays.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from nr_ray_tracer_b200 import _abi as A

INF = float("inf")


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    ct = {np.float64: C.c_double, np.uint32: C.c_uint32}[dtype]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()


class FlatScene:
    def __init__(self, host_scene):
        d = host_scene.desc
        self.d = d
        self.nodes = host_scene.nodes()
        self.child_boxes = host_scene.child_boxes()  # (n, 2 children, lo/hi, 3)
        self.root = d.root
        self.root_box = (list(d.root_box.lo), list(d.root_box.hi))
        ns, npl = d.n_spheres, d.n_planes
        srec = _arr(d.sphere_rec, ns * 4, np.float64).reshape(-1, 4)
        self.sc, self.sr = srec[:, :3].copy(), srec[:, 3].copy()
        self.sm = _arr(d.sphere_material, ns, np.uint32)
        self.so = _arr(d.sphere_order, ns, np.uint32)
        self.sobj = _arr(d.sphere_object, ns, np.uint32)
        prec = _arr(d.plane_rec, npl * 16, np.float64).reshape(-1, 16)  # normal, d, p, w, u, v
        self.pn, self.pd, self.pp = prec[:, 0:3].copy(), prec[:, 3].copy(), prec[:, 4:7].copy()
        self.pw, self.pu, self.pv = prec[:, 7:10].copy(), prec[:, 10:13].copy(), prec[:, 13:16].copy()
        self.pm = _arr(d.plane_material, npl, np.uint32)
        self.po = _arr(d.plane_order, npl, np.uint32)
        self.pobj = _arr(d.plane_object, npl, np.uint32)
        self.inst = [d.instances[i] for i in range(d.n_instances)]
        self.inst_order = _arr(d.instance_order, d.n_instances, np.uint32)
        self.xf = [d.xforms[i] for i in range(d.n_xforms)]
        # the four-slot nodes the kernels walk
        self.wnodes = host_scene.wnodes()
        self.wide_boxes = host_scene.wide_boxes()  # (n, 4 slots, own/gate, lo/hi, 3)
        self.wide_root = d.wide_root
        self.inst_wide_inner = host_scene.instance_wide_inner()


def _dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def _cross(a, b):
    return (a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1])


def _div(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return float("nan")
        neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0)
        return -INF if neg else INF


def _fmax(a, b):
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


def _fmin(a, b):
    if a != a:
        return b
    if b != b:
        return a
    return a if a < b else b


def box_hit(lo, hi, o, d, tmin, tmax):
    a, b = tmin, tmax
    for k in range(3):
        t0, t1 = _div(lo[k] - o[k], d[k]), _div(hi[k] - o[k], d[k])
        mn, mx = (t0, t1) if t0 < t1 else (t1, t0)
        a, b = _fmax(a, mn), _fmin(b, mx)
        if a > b:
            return False
    return True


def _mat3(m, v):
    r = [m[0 + i] * v[0] for i in range(3)]
    r = [r[i] + m[3 + i] * v[1] for i in range(3)]
    return [r[i] + m[6 + i] * v[2] for i in range(3)]


def _mat4(m, v, point):
    r = [m[0 + i] * v[0] for i in range(3)]
    r = [m[3 + i] * v[1] + r[i] for i in range(3)]
    r = [m[6 + i] * v[2] + r[i] for i in range(3)]
    if point:
        r = [m[9 + i] + r[i] for i in range(3)]
    return r


def trace(fs: FlatScene, o, d, tmin=0.001, tmax=INF, tie_by_order=False, wide=False):
    """Returns (t, object index, leaf order in the current space) of the reference's winner, or None.
    wide=True walks the four-slot nodes (nrrt_wnode) with the exact tests the reference makes on the way to each
    slot: the folded-away parent's box for gated slots, the slot's own box for inner nodes, nothing for leaves.
    tie_by_order=False: equal-t ties go to the later child of the flat tree (the reference rule on the reference
    tree).  tie_by_order=True: ties go to the leaf with the larger reference DFS order — what the kernels do, and
    the only rule that is right on an NRRT_BUILD_SAH tree, whose child order is not the reference's."""

    def hit_ref(ref, o, d):
        ty, ix = ref >> A.REF_TYPE_SHIFT, ref & A.REF_INDEX_MASK
        if ref == A.REF_NONE:
            return None
        if ty == A.REF_NODE and wide:
            wn = fs.wnodes[ix]
            res = None
            for sl in range(4):
                cref = int(wn["child"][sl])
                if cref == A.REF_NONE:
                    continue
                if int(wn["meta"][sl]) & A.WNODE_GATED:
                    if not box_hit(fs.wide_boxes[ix, sl, 1, 0], fs.wide_boxes[ix, sl, 1, 1], o, d, tmin, tmax):
                        continue
                if (cref >> A.REF_TYPE_SHIFT) == A.REF_NODE:
                    if not box_hit(fs.wide_boxes[ix, sl, 0, 0], fs.wide_boxes[ix, sl, 0, 1], o, d, tmin, tmax):
                        continue
                h = hit_ref(cref, o, d)
                if h is None:
                    continue
                if res is None or h[0] < res[0]:
                    res = h
                elif h[0] == res[0] and (h[2] > res[2] if tie_by_order else True):  # ties -> later leaf
                    res = h
            return res
        if ty == A.REF_NODE:
            nd = fs.nodes[ix]
            res = None
            for c in range(2):
                cref = int(nd["child"][c])
                if cref == A.REF_NONE:
                    continue
                if (cref >> A.REF_TYPE_SHIFT) == A.REF_NODE:
                    lo, hi = fs.child_boxes[ix, c, 0], fs.child_boxes[ix, c, 1]
                    if not box_hit(lo, hi, o, d, tmin, tmax):
                        continue
                h = hit_ref(cref, o, d)
                if h is None:
                    continue
                if res is None or h[0] < res[0]:
                    res = h
                elif h[0] == res[0] and (h[2] > res[2] if tie_by_order else True):  # ties -> later leaf
                    res = h
            return res
        if ty == A.REF_SPHERE:
            c, r = fs.sc[ix], fs.sr[ix]
            ec = (c[0] - o[0], c[1] - o[1], c[2] - o[2])
            a, h = _dot(d, d), _dot(ec, d)
            cc = _dot(ec, ec) - r * r
            disc = h * h - a * cc
            if disc != disc or disc < 0.0:
                return None
            sq = math.sqrt(disc)
            t = _div(h - sq, a)
            if not (tmin < t < tmax):
                t = _div(h + sq, a)
                if not (tmin < t < tmax):
                    return None
            return (t, int(fs.sobj[ix]), int(fs.so[ix]))
        if ty == A.REF_PLANE:
            n = fs.pn[ix]
            denom = _dot(n, d)
            if abs(denom) < 1e-8:
                return None
            t = _div(fs.pd[ix] - _dot(n, o), denom)
            if not (tmin <= t <= tmax):
                return None
            p = [o[k] + t * d[k] for k in range(3)]
            q = [p[k] - fs.pp[ix][k] for k in range(3)]
            alpha, beta = _dot(fs.pw[ix], _cross(q, fs.pv[ix])), _dot(fs.pw[ix], _cross(fs.pu[ix], q))
            if fs.pm[ix] & 0x80000000:
                ok = alpha > 0.0 and beta > 0.0 and (alpha + beta) < 1.0
            else:
                ok = 0.0 <= alpha <= 1.0 and 0.0 <= beta <= 1.0
            return (t, int(fs.pobj[ix]), int(fs.po[ix])) if ok else None
        if ty == A.REF_INSTANCE:
            ins = fs.inst[ix]
            oo, dd = list(o), list(d)
            for k in range(ins.n_xforms):
                x = fs.xf[ins.first_xform + k]
                m = list(x.to_obj)
                if x.kind == 0:
                    oo = [oo[i] - m[i] for i in range(3)]
                elif x.kind == 1:
                    oo, dd = _mat3(m, oo), _mat3(m, dd)
                else:
                    oo, dd = _mat4(m, oo, True), _mat4(m, dd, False)
            inner = int(fs.inst_wide_inner[ix]) if wide else ins.inner
            if inner == A.REF_NONE:
                return None
            if (inner >> A.REF_TYPE_SHIFT) == A.REF_NODE:
                if not box_hit(list(ins.inner_box.lo), list(ins.inner_box.hi), oo, dd, tmin, tmax):
                    return None
            h = hit_ref(inner, oo, dd)
            return None if h is None else (h[0], h[1], int(fs.inst_order[ix]))
        return None

    root = fs.wide_root if wide else fs.root
    if root != A.REF_NONE and (root >> A.REF_TYPE_SHIFT) == A.REF_NODE:
        if not box_hit(fs.root_box[0], fs.root_box[1], o, d, tmin, tmax):
            return None
    return hit_ref(root, list(o), list(d))
