"""ctypes loader of oracle/libpvoracle.so (the CPU restatement).  Test infrastructure: imported only from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libpvoracle.so")

RAY_DTYPE = np.dtype([("org", "<f8", 3), ("dir", "<f8", 3), ("depth", "<f8"), ("obj", "<i4"), ("aux", "<i4")])


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        l = C.CDLL(LIB)
        l.pvo_scene_load.restype = C.c_void_p
        l.pvo_scene_load.argtypes = [C.c_char_p]
        l.pvo_scene_destroy.argtypes = [C.c_void_p]
        l.pvo_render.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_ulonglong)]
        l.pvo_trace_rays.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint32)]
        l.pvo_camera_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_double)]
        l.pvo_solve_polynomial.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_double]
        l.pvo_noise.restype = C.c_double
        l.pvo_noise.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int]
        l.pvo_turbulence.restype = C.c_double
        l.pvo_turbulence.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
        l.pvo_dnoise.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]
        _lib = l
    return _lib


class OracleScene:
    def __init__(self, path):
        self._h = lib().pvo_scene_load(str(path).encode())
        if not self._h:
            raise IOError(f"oracle cannot load {path}")

    def close(self):
        if self._h:
            lib().pvo_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, width, height, rect=None, threads=1):
        """Returns (H x W x 4 float32 image of the rectangle, stats dict)."""
        l, t, r, b = rect if rect is not None else (0, 0, width - 1, height - 1)
        out = np.zeros((b - t + 1, r - l + 1, 4), dtype=np.float32)
        st = (C.c_ulonglong * 3)()
        lib().pvo_render(self._h, width, height, l, t, r, b, out.ctypes.data_as(C.POINTER(C.c_float)), threads, st)
        return out, dict(rays=st[0], shadow_ray_tests=st[1], max_trace_level=st[2])

    def trace_rays(self, org_dir):
        rays = np.ascontiguousarray(org_dir, dtype=np.float64).reshape(-1, 6)
        n = len(rays)
        obj = np.empty(n, dtype=np.int32)
        depth = np.empty(n, dtype=np.float64)
        aux = np.empty(n, dtype=np.uint32)
        lib().pvo_trace_rays(self._h, rays.ctypes.data_as(C.POINTER(C.c_double)), n, obj.ctypes.data_as(C.POINTER(C.c_int32)),
                             depth.ctypes.data_as(C.POINTER(C.c_double)), aux.ctypes.data_as(C.POINTER(C.c_uint32)))
        return obj.astype(np.int64), depth, aux

    def camera_rays(self, width, height, xy):
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        out = np.empty((len(xy), 6), dtype=np.float64)
        lib().pvo_camera_rays(self._h, width, height, xy.ctypes.data_as(C.POINTER(C.c_double)), len(xy), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out


def _setup_aa(l):
    l.pvo_render_aa.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_ulonglong)]


def render_aa(scene, width, height, rects, method, depth=3, threshold=0.3, jitter=1.0, gamma=2.5, threads=1):
    """Anti-aliased render of `rects` [(l, t, r, b)] by the oracle; returns (rect-major [n,4] float32 pixels, stats)."""
    l = lib()
    _setup_aa(l)
    ra = np.ascontiguousarray(np.array(rects, dtype=np.int32).reshape(-1, 4))
    n = int(((ra[:, 2] - ra[:, 0] + 1) * (ra[:, 3] - ra[:, 1] + 1)).sum())
    out = np.zeros((n, 4), dtype=np.float32)
    st = (C.c_ulonglong * 4)()
    l.pvo_render_aa(scene._h, width, height, ra.ctypes.data_as(C.POINTER(C.c_int)), len(ra), method, depth, threshold, jitter, gamma,
                    out.ctypes.data_as(C.POINTER(C.c_float)), threads, st)
    return out, dict(rays=st[0], shadow_ray_tests=st[1], max_trace_level=st[2], samples=st[3])
