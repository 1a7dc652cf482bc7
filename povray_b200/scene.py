"""Host-side handle of a flattened scene and of the render calls (thin ctypes layer over the C ABI).

Mirrors, on the Python side, what the reference's `TraceTask` does with its `ViewData`
(source/backend/render/tracetask.cpp:287-454, source/backend/scene/view.cpp:236-473): get rectangles,
trace them, hand back row-major RGBT pixels.
"""
import ctypes as C

import numpy as np

from . import _abi as A


def tiles(width, height, block=32, left=0, top=0):
    """Rectangles in the order `ViewData::GetNextRectangle` hands them out for the default render
    pattern: `block` x `block` tiles (view.cpp:810), row-major over the image (view.cpp:126-271)."""
    out = []
    for y in range(top, height, block):
        for x in range(left, width, block):
            out.append((x, y, min(x + block, width) - 1, min(y + block, height) - 1))
    return out


def _rect_array(rects):
    arr = (A.Rect * len(rects))()
    for i, (l, t, r, b) in enumerate(rects):
        arr[i].left, arr[i].top, arr[i].right, arr[i].bottom = int(l), int(t), int(r), int(b)
    return arr


def _area(rects):
    return sum((r - l + 1) * (b - t + 1) for l, t, r, b in rects)


def assemble(pixels, rects, width, height):
    """Scatters rect-major RGBT pixels (what `CompletedRectangle` receives) into an H x W x 4 image."""
    img = np.zeros((height, width, 4), dtype=np.float32)
    pos = 0
    for l, t, r, b in rects:
        n = (r - l + 1) * (b - t + 1)
        img[t:b + 1, l:r + 1] = pixels[pos:pos + n].reshape(b - t + 1, r - l + 1, 4)
        pos += n
    return img


class HostBuffer:
    """Page-locked host memory from `pvgpu_host_alloc`, exposed as a float32 numpy array (`.array`)."""

    def __init__(self, n_floats):
        self._p = A.lib().pvgpu_host_alloc(int(n_floats) * 4)
        if not self._p:
            raise MemoryError("pvgpu_host_alloc failed")
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_float)), shape=(int(n_floats),))

    def close(self):
        if self._p:
            self.array = None
            A.lib().pvgpu_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """Owns a `pvgpu_scene*`."""

    def __init__(self, handle):
        self._h = handle
        self.finalized = False

    # -- construction ---------------------------------------------------------------------------
    @classmethod
    def create(cls, globals_):
        h = A.VP()
        A.check(A.lib().pvgpu_scene_create(C.byref(h), C.byref(globals_)))
        return cls(h)

    @classmethod
    def load(cls, path):
        h = A.VP()
        A.check(A.lib().pvgpu_scene_load(C.byref(h), str(path).encode()))
        return cls(h)

    def save(self, path):
        A.check(A.lib().pvgpu_scene_save(self._h, str(path).encode()))

    def close(self):
        if self._h:
            A.lib().pvgpu_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_camera(self, cam):
        A.check(A.lib().pvgpu_scene_set_camera(self._h, C.byref(cam)))

    def get_camera(self):
        cam = A.Camera()
        A.check(A.lib().pvgpu_scene_get_camera(self._h, C.byref(cam)))
        return cam

    def build_tree(self):
        A.check(A.lib().pvgpu_scene_build_tree(self._h))

    def finalize(self, device=0):
        A.check(A.lib().pvgpu_scene_finalize(self._h, int(device)))
        self.finalized = True
        return self

    def finalize_multi(self, devices=None, n_devices=0):
        """Replicates the scene on several CUDA devices (pvgpu_scene_finalize_multi): `devices` = list of device indices, or
        `n_devices` (0 = every visible device); render calls then shard their rectangles over the devices."""
        if devices is not None:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            A.check(A.lib().pvgpu_scene_finalize_multi(self._h, arr, len(devices)))
        else:
            A.check(A.lib().pvgpu_scene_finalize_multi(self._h, None, int(n_devices)))
        self.finalized = True
        return self

    @property
    def device_count(self):
        return int(A.lib().pvgpu_scene_device_count(self._h))

    @property
    def device_bytes(self):
        return int(A.lib().pvgpu_scene_device_bytes(self._h))

    # -- rendering ------------------------------------------------------------------------------
    def render(self, width, height, rects=None, aa=None, out=None):
        """Host-buffer render (pvgpu_render): returns (rect-major pixels [n,4] float32, stats dict).  `out`: optional
        float32 array with room for n * 4 values (e.g. a `HostBuffer(...).array`, which avoids the staging copy)."""
        if rects is None:
            rects = tiles(width, height)
        n = _area(rects)
        if out is None:
            out = np.empty((n, 4), dtype=np.float32)
        else:
            out = out.reshape(-1)[:n * 4].reshape(n, 4)
        st = A.Stats()
        ra = _rect_array(rects)
        A.check(A.lib().pvgpu_render(self._h, C.byref(aa) if aa is not None else None, int(width), int(height), ra, len(rects),
                                     out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st), None, None))
        return out, st.as_dict()

    def render_image(self, width, height, aa=None, block=32):
        rects = tiles(width, height, block)
        px, st = self.render(width, height, rects, aa)
        return assemble(px, rects, width, height), st

    def render_device(self, width, height, rects, out_ptr, stream=0, aa=None):
        """Device-buffer render (pvgpu_render_device): `out_ptr` is a CUDA device pointer (e.g. a torch tensor's
        data_ptr()) with room for area x 4 floats; work is enqueued on `stream` (a cudaStream_t value)."""
        st = A.Stats()
        ra = rects if isinstance(rects, C.Array) else _rect_array(rects)
        A.check(A.lib().pvgpu_render_device(self._h, C.byref(aa) if aa is not None else None, int(width), int(height), ra, len(ra),
                                            C.c_void_p(int(out_ptr)), C.byref(st), C.c_void_p(int(stream))))
        return st.as_dict()

    def trace_rays(self, org_dir):
        """Ray-level harness (pvgpu_trace_rays): org_dir [n,6] float64 -> (object index int64 with -1 = miss, depth, aux)."""
        rays = np.ascontiguousarray(org_dir, dtype=np.float64).reshape(-1, 6)
        n = rays.shape[0]
        obj = np.empty(n, dtype=np.uint32)
        depth = np.empty(n, dtype=np.float64)
        aux = np.empty(n, dtype=np.uint32)
        A.check(A.lib().pvgpu_trace_rays(self._h, rays.ctypes.data_as(C.POINTER(C.c_double)), n,
                                         obj.ctypes.data_as(C.POINTER(C.c_uint32)), depth.ctypes.data_as(C.POINTER(C.c_double)),
                                         aux.ctypes.data_as(C.POINTER(C.c_uint32))))
        o = obj.astype(np.int64)
        o[obj == 0xFFFFFFFF] = -1
        return o, depth, aux

    def camera_rays(self, width, height, xy):
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        out = np.empty((xy.shape[0], 6), dtype=np.float64)
        A.check(A.lib().pvgpu_camera_rays(self._h, int(width), int(height), xy.ctypes.data_as(C.POINTER(C.c_double)),
                                          xy.shape[0], out.ctypes.data_as(C.POINTER(C.c_double))))
        return out
