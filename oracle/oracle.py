"""ctypes wrapper of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from nr_ray_tracer_b200 import _abi as A
from nr_ray_tracer_b200.scene_config import SceneGraph

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "oracle.cpp"), os.path.join(_HERE, "..", "include", "nrrt.h")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_build.restype = C.c_void_p
        L.oracle_build.argtypes = [C.POINTER(A.GraphDesc)]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_camera_build.argtypes = [C.POINTER(A.CameraConfig), C.POINTER(A.Camera)]
        L.oracle_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_void_p,
                                        C.c_void_p, C.c_int]
        L.oracle_set_trace_time.argtypes = [C.c_double]
        L.oracle_render.argtypes = [C.c_void_p, C.POINTER(A.Camera), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_render_chunked.argtypes = [C.c_void_p, C.POINTER(A.Camera), C.c_uint64, C.c_uint32, C.c_uint32,
                                            C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_texture_eval.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
        L.oracle_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.oracle_perm_table.argtypes = [C.c_uint32, C.c_void_p]
        L.oracle_num_threads.restype = C.c_int
        _lib = L
    return _lib


def camera_build(cfg: A.CameraConfig) -> A.Camera:
    cam = A.Camera()
    if lib().oracle_camera_build(C.byref(cfg), C.byref(cam)) != 0:
        raise ValueError("oracle_camera_build failed")
    return cam


def philox(seed: int, c0: int, c1: int, c2: int, c3: int) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox(seed, c0, c1, c2, c3, out.ctypes.data)
    return out


def perm_table(seed: int) -> np.ndarray:
    out = np.zeros(256, dtype=np.uint8)
    lib().oracle_perm_table(seed, out.ctypes.data)
    return out


def num_threads() -> int:
    return lib().oracle_num_threads()


class OracleScene:
    def __init__(self, graph: SceneGraph):
        self._holder = graph.to_desc()
        self._h = lib().oracle_build(self._holder.ptr())
        if not self._h:
            raise ValueError("oracle_build failed (malformed graph)")

    def close(self):
        if self._h:
            lib().oracle_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trace_rays(self, rays: np.ndarray, tmin: float = 0.001, tmax: float = float("inf"), n_threads: int = 0,
                   time: float = 0.0):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(rays.shape[0], dtype=A.HIT_DTYPE)
        counters = np.zeros(2, dtype=np.uint64)
        lib().oracle_set_trace_time(float(time))
        rc = lib().oracle_trace_rays(self._h, rays.ctypes.data, rays.shape[0], tmin, tmax, out.ctypes.data,
                                     counters.ctypes.data, n_threads)
        if rc != 0:
            raise ValueError("oracle_trace_rays failed")
        return out, {"aabb_tests": int(counters[0]), "prim_tests": int(counters[1])}

    def render(self, cam: A.Camera, seed: int = 0, pixel_range=None, sample_range=None, n_threads: int = 0):
        W, H = cam.width, cam.height
        p0, p1 = pixel_range if pixel_range is not None else (0, W * H)
        s0, s1 = sample_range if sample_range is not None else (0, cam.samples_per_pixel)
        out = np.zeros((H, W, 3), dtype=np.float32)
        counters = np.zeros(4, dtype=np.uint64)
        rc = lib().oracle_render(self._h, C.byref(cam), seed, p0, p1, s0, s1, out.ctypes.data, counters.ctypes.data,
                                 n_threads)
        if rc != 0:
            raise ValueError("oracle_render failed")
        return out, {"paths": int(counters[0]), "segments": int(counters[1]), "aabb_tests": int(counters[2]),
                     "prim_tests": int(counters[3])}

    def render_chunked(self, cam: A.Camera, chunk_starts, seed: int = 0, pixel_range=None, n_threads: int = 0):
        """render() with the per-pixel sum associated as given: samples [chunk_starts[c], chunk_starts[c+1]) are summed
        in order into partials, the partials added in order (oracle_render_chunked).  One chunk = render()."""
        W, H = cam.width, cam.height
        p0, p1 = pixel_range if pixel_range is not None else (0, W * H)
        st = np.ascontiguousarray(chunk_starts, dtype=np.uint32)
        out = np.zeros((H, W, 3), dtype=np.float32)
        counters = np.zeros(4, dtype=np.uint64)
        rc = lib().oracle_render_chunked(self._h, C.byref(cam), seed, p0, p1, st.ctypes.data, len(st) - 1,
                                         out.ctypes.data, counters.ctypes.data, n_threads)
        if rc != 0:
            raise ValueError("oracle_render_chunked failed")
        return out, {"paths": int(counters[0]), "segments": int(counters[1]), "aabb_tests": int(counters[2]),
                     "prim_tests": int(counters[3])}

    def texture_eval(self, texture: int, uvp: np.ndarray) -> np.ndarray:
        uvp = np.ascontiguousarray(uvp, dtype=np.float64).reshape(-1, 5)
        out = np.zeros((uvp.shape[0], 3), dtype=np.float64)
        if lib().oracle_texture_eval(self._h, texture, uvp.ctypes.data, uvp.shape[0], out.ctypes.data) != 0:
            raise ValueError("oracle_texture_eval failed")
        return out
