"""ctypes image of include/pvgpu.h and loader of the in-tree CUDA library.

The library is the product: there is no Python or CPU implementation of the trace path behind it.
If ``libpvgpu.so`` has not been built (``python -c 'import __graft_entry__ as g; g.build()'``) importing
this module raises, and every render entry point fails loudly when no CUDA device is present.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PVGPU_LIB") or os.path.join(HERE, "libpvgpu.so")

ABI_VERSION = 3

# error codes
OK, E_INVALID, E_UNSUPPORTED, E_NO_DEVICE, E_CUDA, E_IO, E_ABORTED, E_OVERFLOW = 0, -1, -2, -3, -4, -5, -6, -7

# object kinds
OBJ_SPHERE, OBJ_BOX, OBJ_PLANE, OBJ_QUADRIC, OBJ_TORUS, OBJ_MESH, OBJ_CSG_UNION, OBJ_CSG_INTERSECTION, OBJ_CSG_MERGE, OBJ_BLOB, OBJ_CONE, OBJ_DISC, OBJ_TRIANGLE, OBJ_POLYGON, OBJ_POLY = range(1, 16)

# object flags (source/core/scene/object.h:88-117)
NO_SHADOW_FLAG = 0x00000001
INVERTED_FLAG = 0x00000004
STURM_FLAG = 0x00000040
OPAQUE_FLAG = 0x00000080
MULTITEXTURE_FLAG = 0x00000100
INFINITE_FLAG = 0x00000200
HOLLOW_FLAG = 0x00000800
UV_FLAG = 0x00002000
DOUBLE_ILLUMINATE_FLAG = 0x00004000
NO_IMAGE_FLAG = 0x00008000
NO_REFLECTION_FLAG = 0x00010000
NO_GLOBAL_LIGHTS_FLAG = 0x00020000
HIERARCHY_FLAG = 0x00000400
CLOSED_FLAG = 0x00000002

NODE_INFINITE = 1
TRI_SMOOTH, TRI_THREETEX = 1, 2
LIGHT_POINT, LIGHT_SPOT, LIGHT_FILL, LIGHT_CYLINDER = 1, 2, 3, 4
LIGHT_AREA, LIGHT_PARALLEL, LIGHT_MEDIA_ATTEN, LIGHT_MEDIA_INTERACT = 0x001, 0x020, 0x040, 0x080
(PAT_PLAIN, PAT_CHECKER, PAT_BOZO, PAT_GRANITE, PAT_GRADIENT, PAT_MARBLE, PAT_WRINKLES, PAT_ONION, PAT_BRICK,
 PAT_HEXAGON, PAT_SPOTTED, PAT_AGATE) = range(1, 13)
WAVE_RAW, WAVE_RAMP, WAVE_SINE, WAVE_TRIANGLE, WAVE_SCALLOP, WAVE_CUBIC, WAVE_POLY = range(7)
WARP_TRANSFORM, WARP_TURBULENCE, WARP_CLASSIC_TURBULENCE = 1, 2, 3
Q_AMBIENT_ONLY, Q_QUICK_COLOUR, Q_SHADOWS, Q_AREA_LIGHTS, Q_REFRACTIONS, Q_REFLECTIONS, Q_NORMALS, Q_MEDIA = 1, 2, 4, 8, 16, 32, 64, 128
Q_DEFAULT = Q_SHADOWS | Q_AREA_LIGHTS | Q_REFRACTIONS | Q_REFLECTIONS | Q_NORMALS | Q_MEDIA
CAMERA_PERSPECTIVE, CAMERA_ORTHOGRAPHIC = 1, 2

u8, u16, u32, i32, u64, f32, f64 = C.c_uint8, C.c_uint16, C.c_uint32, C.c_int32, C.c_uint64, C.c_float, C.c_double


class Object(C.Structure):
    _fields_ = [("type", u32), ("flags", u32), ("texture", i32), ("interior_texture", i32), ("interior", i32),
                ("transform", i32), ("parent", i32), ("child_first", u32), ("child_count", u32),
                ("clip_first", u32), ("clip_count", u32), ("bound_first", u32), ("bound_count", u32),
                ("mesh", i32), ("aux", u32), ("bbox", f32 * 6), ("reserved", u32), ("p", f64 * 10)]


class Transform(C.Structure):
    _fields_ = [("matrix", f64 * 16), ("inverse", f64 * 16)]


class Node(C.Structure):
    _fields_ = [("lo", f32 * 3), ("size", f32 * 3), ("first", u32), ("count", u16), ("flags", u16)]


class Triangle(C.Structure):
    _fields_ = [("perp", f32 * 3), ("distance", f32), ("normal_ind", i32), ("p1", i32), ("p2", i32), ("p3", i32),
                ("n1", i32), ("n2", i32), ("n3", i32), ("texture", i32), ("texture2", i32), ("texture3", i32),
                ("flags", u8), ("dominant_axis", u8), ("v_axis", u8), ("reserved", u8)]


class Mesh(C.Structure):
    _fields_ = [("vertex_first", u32), ("vertex_count", u32), ("normal_first", u32), ("normal_count", u32),
                ("triangle_first", u32), ("triangle_count", u32), ("node_first", u32), ("node_count", u32),
                ("texture_first", u32), ("texture_count", u32), ("has_inside_vector", u32), ("reserved", u32),
                ("inside_vector", f64 * 3)]


class BlobElement(C.Structure):
    _fields_ = [("type", u32), ("transform", i32), ("o", f64 * 3), ("len", f64), ("rad2", f64), ("c", f64 * 3)]


class BlobNode(C.Structure):
    _fields_ = [("c", f64 * 3), ("r2", f64), ("first", u32), ("count", u32)]


class Blob(C.Structure):
    _fields_ = [("element_first", u32), ("element_count", u32), ("node_first", u32), ("node_count", u32), ("threshold", f64)]


class Light(C.Structure):
    _fields_ = [("type", u32), ("flags", u32), ("colour", f32 * 3), ("projected_through", i32),
                ("center", f64 * 3), ("direction", f64 * 3), ("points_at", f64 * 3), ("axis1", f64 * 3), ("axis2", f64 * 3),
                ("coeff", f64), ("radius", f64), ("falloff", f64), ("fade_distance", f64), ("fade_power", f64),
                ("area_size1", i32), ("area_size2", i32), ("adaptive_level", i32), ("object_flags", u32)]


class Warp(C.Structure):
    _fields_ = [("type", u32), ("transform", i32), ("turbulence", f64 * 3), ("octaves", i32), ("lambda_", f32),
                ("omega", f32), ("handled_by_pattern", u32)]


class BlendEntry(C.Structure):
    _fields_ = [("value", f32), ("colour", f32 * 5)]


class BlendMap(C.Structure):
    _fields_ = [("entry_first", u32), ("entry_count", u32), ("blend_mode", i32), ("blend_gamma", f32)]


class Pigment(C.Structure):
    _fields_ = [("pattern", u32), ("wave_type", u32), ("frequency", f32), ("phase", f32), ("exponent", f32),
                ("noise_generator", i32), ("warp_first", u32), ("warp_count", u32), ("blend_map", i32),
                ("colour", f32 * 5), ("quick_colour", f32 * 5), ("data", u32), ("p", f64 * 4)]


class Finish(C.Structure):
    _fields_ = [("diffuse", f32), ("diffuse_back", f32), ("brilliance", f32), ("brilliance_adjust", f32),
                ("brilliance_adjust_rad", f32), ("specular", f32), ("roughness", f32), ("phong", f32), ("phong_size", f32),
                ("irid", f32), ("irid_film_thickness", f32), ("irid_turb", f32), ("reflect_exp", f32), ("crand", f32),
                ("metallic", f32), ("ambient", f32 * 3), ("emission", f32 * 3), ("reflection_max", f32 * 3),
                ("reflection_min", f32 * 3), ("reflection_falloff", f32), ("fresnel", f32), ("reflect_metallic", f32),
                ("reflection_fresnel", i32), ("conserve_energy", i32), ("alpha_knockout", i32), ("use_subsurface", i32)]


class Texture(C.Structure):
    _fields_ = [("type", u32), ("next", i32), ("pigment", i32), ("finish", i32), ("tnormal", i32), ("blend_map", i32)]


class SlopeEntry(C.Structure):
    _fields_ = [("value", f32), ("reserved", u32), ("height", f64), ("slope", f64)]


class TNormal(C.Structure):
    _fields_ = [("type", u32), ("flags", u32), ("pattern", i32), ("slope_first", u32), ("slope_count", u32),
                ("amount", f32), ("delta", f32), ("normal_map", u32)]


class SkySphere(C.Structure):
    _fields_ = [("pigment_first", u32), ("pigment_count", u32), ("transform", i32), ("emission", f32 * 3)]


class Fog(C.Structure):
    _fields_ = [("type", u32), ("turbulence", i32), ("distance", f64), ("alt", f64), ("offset", f64), ("up", f64 * 3),
                ("colour", f32 * 5), ("turb_depth", f32)]


class Interior(C.Structure):
    _fields_ = [("hollow", i32), ("disp_nelems", i32), ("ior", f32), ("dispersion", f32), ("caustics", f32),
                ("old_refract", f32), ("fade_distance", f32), ("fade_power", f32), ("fade_colour", f32 * 3), ("reserved", u32)]


class Globals(C.Structure):
    _fields_ = [("max_trace_level", u32), ("language_version", u32), ("noise_generator", i32), ("bounding_method", u32),
                ("quality_flags", u32), ("output_alpha", i32), ("adc_bailout", f64), ("ambient_light", f32 * 3),
                ("background", f32 * 5), ("atmosphere_ior", f32), ("atmosphere_dispersion", f32),
                ("number_of_waves", u32), ("reserved", u32)]


class Camera(C.Structure):
    _fields_ = [("type", u32), ("reserved", u32), ("location", f64 * 3), ("direction", f64 * 3), ("up", f64 * 3),
                ("right", f64 * 3), ("max_ray_distance", f64)]


class AA(C.Structure):
    _fields_ = [("method", u32), ("depth", u32), ("threshold", f64), ("jitter_scale", f64), ("gamma", f64)]


class Rect(C.Structure):
    _fields_ = [("left", i32), ("top", i32), ("right", i32), ("bottom", i32)]


class Image(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("map_type", u32), ("interpolation", u32), ("flags", u32), ("data_first", u32),
                ("fwidth", f32), ("fheight", f32), ("all_filter", f32), ("all_transmit", f32), ("gradient", f64 * 3), ("offset", f64 * 2)]


class Stats(C.Structure):
    _fields_ = [("rays", u64), ("shadow_ray_tests", u64), ("reflected_rays", u64), ("refracted_rays", u64),
                ("transmitted_rays", u64), ("tir_rays", u64), ("adc_saves", u64), ("samples", u64), ("waves", u64),
                ("kernel_launches", u64), ("max_trace_level", u32), ("overflow", u32), ("node_tests_closest", u64), ("prim_tests_closest", u64), ("node_tests_shadow", u64), ("prim_tests_shadow", u64), ("device_ms", f64),
                ("kernel_ms", f64 * 5), ("kernel_count", u64 * 5), ("kernel_items", u64 * 5)]

    KERNELS = ("primary", "closest", "shade", "shadow", "aa")

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_[:17]}
        for i, k in enumerate(self.KERNELS):
            d[k + "_ms"] = self.kernel_ms[i]
            d[k + "_launches"] = int(self.kernel_count[i])
            d[k + "_rays"] = int(self.kernel_items[i])
        return d


# every symbol include/pvgpu.h declares, with its signature
P = C.POINTER
VP = C.c_void_p
SIGNATURES = {
    "pvgpu_abi_version": (C.c_int, []),
    "pvgpu_last_error": (C.c_char_p, []),
    "pvgpu_scene_create": (C.c_int, [P(VP), P(Globals)]),
    "pvgpu_scene_destroy": (None, [VP]),
    "pvgpu_scene_set_objects": (C.c_int, [VP, P(Object), C.c_size_t, P(u32), C.c_size_t, P(u32), C.c_size_t]),
    "pvgpu_scene_set_transforms": (C.c_int, [VP, P(Transform), C.c_size_t]),
    "pvgpu_scene_set_tree": (C.c_int, [VP, P(Node), C.c_size_t]),
    "pvgpu_scene_build_tree": (C.c_int, [VP]),
    "pvgpu_scene_set_blobs": (C.c_int, [VP, P(Blob), C.c_size_t, P(BlobElement), C.c_size_t, P(BlobNode), C.c_size_t]),
    "pvgpu_scene_set_meshes": (C.c_int, [VP, P(Mesh), C.c_size_t, P(f32), C.c_size_t, P(f32), C.c_size_t,
                                         P(Triangle), C.c_size_t, P(Node), C.c_size_t]),
    "pvgpu_scene_set_images": (C.c_int, [VP, P(Image), C.c_size_t, P(f32), C.c_size_t]),
    "pvgpu_scene_set_blob_textures": (C.c_int, [VP, P(i32), C.c_size_t]),
    "pvgpu_scene_set_mesh_uv": (C.c_int, [VP, P(C.c_double), C.c_size_t, P(C.c_uint32), C.c_size_t]),
    "pvgpu_scene_set_shape_data": (C.c_int, [VP, P(f64), C.c_size_t]),
    "pvgpu_scene_set_lights": (C.c_int, [VP, P(Light), C.c_size_t]),
    "pvgpu_scene_set_materials": (C.c_int, [VP, P(Texture), C.c_size_t, P(Pigment), C.c_size_t, P(Finish), C.c_size_t,
                                            P(BlendMap), C.c_size_t, P(BlendEntry), C.c_size_t, P(Warp), C.c_size_t,
                                            P(Interior), C.c_size_t]),
    "pvgpu_scene_set_normals": (C.c_int, [VP, P(TNormal), C.c_size_t, P(SlopeEntry), C.c_size_t]),
    "pvgpu_scene_set_irid_wavelengths": (C.c_int, [VP, P(f32)]),
    "pvgpu_scene_set_atmosphere": (C.c_int, [VP, P(SkySphere), P(Fog), C.c_size_t]),
    "pvgpu_scene_set_camera": (C.c_int, [VP, P(Camera)]),
    "pvgpu_scene_get_camera": (C.c_int, [VP, P(Camera)]),
    "pvgpu_scene_set_camera_angles": (C.c_int, [VP, C.c_double, C.c_double, C.c_double]),
    "pvgpu_scene_add_mesh2": (C.c_int, [VP, P(f64), C.c_size_t, P(i32), C.c_size_t, P(i32)]),
    "pvgpu_scene_finalize": (C.c_int, [VP, C.c_int]),
    "pvgpu_scene_finalize_multi": (C.c_int, [VP, P(C.c_int), C.c_int]),
    "pvgpu_scene_device_count": (C.c_int, [VP]),
    "pvgpu_prewarm": (None, [C.c_int]),
    "pvgpu_fp64_peak": (C.c_int, [C.c_int, P(C.c_double)]),
    "pvgpu_scene_device_bytes": (C.c_size_t, [VP]),
    "pvgpu_scene_save": (C.c_int, [VP, C.c_char_p]),
    "pvgpu_scene_load": (C.c_int, [P(VP), C.c_char_p]),
    "pvgpu_render": (C.c_int, [VP, P(AA), C.c_int, C.c_int, P(Rect), C.c_size_t, P(f32), P(Stats), VP, VP]),
    "pvgpu_solve_polynomial": (C.c_int, [VP, C.c_size_t, P(i32), P(i32), P(f64), P(f64), P(f64), P(i32)]),
    "pvgpu_noise": (C.c_int, [VP, C.c_size_t, P(f64), P(i32), P(i32), P(f64)]),
    "pvgpu_host_alloc": (VP, [C.c_size_t]),
    "pvgpu_host_free": (None, [VP]),
    "pvgpu_render_device": (C.c_int, [VP, P(AA), C.c_int, C.c_int, P(Rect), C.c_size_t, VP, P(Stats), VP]),
    "pvgpu_trace_rays": (C.c_int, [VP, P(f64), C.c_size_t, P(u32), P(f64), P(u32)]),
    "pvgpu_camera_rays": (C.c_int, [VP, C.c_int, C.c_int, P(f64), C.c_size_t, P(f64)]),
}

_lib = None


class PvgpuError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"pvgpu error {code}: {message}")
        self.code = code


def lib():
    """Loads povray_b200/libpvgpu.so (once) and binds every exported entry point."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                              "(the trace path has no Python / CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.pvgpu_abi_version() != ABI_VERSION:
            raise ImportError("libpvgpu.so ABI version mismatch")
        _lib = l
    return _lib


def check(rc):
    if rc != OK:
        raise PvgpuError(rc, lib().pvgpu_last_error().decode("utf-8", "replace"))
