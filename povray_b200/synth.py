"""Synthetic scenes of BASELINE.json and a small scene builder with two back ends:

  * `to_pov()`    - POV-Ray SDL text for the reference binary / the reference-side adapter;
  * `build()`     - the flat tables of include/pvgpu.h, filled the way POV-Ray's parser fills `SceneData`
                    for these constructs (Parse_Sphere parser.cpp:5715, Parse_Plane :4979, Parse_Box, Parse_Mesh2
                    :4079, Post_Process :9004-9296; Sphere/Box/Plane::Compute_BBox; default FINISH / Interior),
                    followed by the reference's own tree build (pvgpu_scene_build_tree).

Only what the synthetic benchmark scenes need is covered (spheres, planes, boxes, mesh2, plain / checker
pigments, finishes with phong / specular / reflection, point lights, perspective camera).  Everything else
reaches the GPU path through the reference parser and the adapter.  The tests compare the
two back ends table by table.
"""
import ctypes as C
import math

import numpy as np

from . import _abi as A
from .scene import Scene

BOUND_HUGE = 2.0e10


def _f(x):
    return repr(float(x))


def _vec(v):
    return "<" + ", ".join(_f(c) for c in v) + ">"


def _f32(x):
    return float(np.float32(x))


class SceneBuilder:
    def __init__(self, max_trace_level=5, adc_bailout=1.0 / 255.0, ambient_light=(1.0, 1.0, 1.0),
                 background=(0.0, 0.0, 0.0), version=370, noise_generator=2, assumed_gamma=1.0):
        self.g = dict(max_trace_level=max_trace_level, adc_bailout=adc_bailout, ambient_light=ambient_light,
                      background=background, version=version, noise_generator=noise_generator, assumed_gamma=assumed_gamma)
        self.objects = []      # dicts: kind, params, texture
        self.lights = []
        self.cam = None

    # ---- materials ----------------------------------------------------------------------------------
    @staticmethod
    def texture(pigment, ambient=0.1, diffuse=0.6, phong=0.0, phong_size=40.0, specular=0.0, roughness=0.05,
                reflection=0.0, filter=0.0, transmit=0.0, checker=None):
        """pigment: rgb triple; checker=(rgb1, rgb2) makes a checker pigment instead."""
        return dict(pigment=tuple(float(c) for c in pigment), ambient=ambient, diffuse=diffuse, phong=phong, phong_size=phong_size,
                    specular=specular, roughness=roughness, reflection=reflection, filter=filter, transmit=transmit,
                    checker=None if checker is None else (tuple(map(float, checker[0])), tuple(map(float, checker[1]))))

    # ---- objects ------------------------------------------------------------------------------------
    def sphere(self, center, radius, texture):
        self.objects.append(dict(kind="sphere", center=tuple(map(float, center)), radius=float(radius), texture=texture))

    def plane(self, normal, distance, texture):
        self.objects.append(dict(kind="plane", normal=tuple(map(float, normal)), distance=float(distance), texture=texture))

    def box(self, c1, c2, texture):
        self.objects.append(dict(kind="box", c1=tuple(map(float, c1)), c2=tuple(map(float, c2)), texture=texture))

    def mesh2(self, vertices, faces, texture):
        self.objects.append(dict(kind="mesh2", vertices=np.ascontiguousarray(vertices, dtype=np.float64),
                                 faces=np.ascontiguousarray(faces, dtype=np.int32), texture=texture))

    def light(self, position, colour=(1.0, 1.0, 1.0)):
        self.lights.append(dict(position=tuple(map(float, position)), colour=tuple(map(float, colour))))

    def camera(self, location, look_at, angle=50.0, aspect=16.0 / 9.0, sky=(0.0, 1.0, 0.0)):
        """Explicit location / direction / up / right vectors (what the parser's look_at + angle code derives);
        the SDL output states the four vectors verbatim so no parser-side trigonometry is involved."""
        loc = np.array(location, dtype=np.float64)
        d = np.array(look_at, dtype=np.float64) - loc
        d /= np.sqrt((d * d).sum())
        right = np.cross(np.array(sky, dtype=np.float64), d)
        right /= np.sqrt((right * right).sum())
        up = np.cross(d, right)
        dir_len = 0.5 * aspect / math.tan(math.radians(angle) / 2.0)
        z = lambda v: tuple(float(c) + 0.0 for c in v)      # -0.0 -> +0.0 (the SDL parser reads "-0.0" as 0 - 0)
        self.cam = dict(location=z(loc), direction=z(d * dir_len), up=z(up), right=z(right * aspect))

    # ---- SDL ----------------------------------------------------------------------------------------
    @staticmethod
    def _texture_sdl(t):
        if t["checker"] is not None:
            a, b = t["checker"]
            pig = f"checker rgb {_vec(a)}, rgb {_vec(b)}"
        elif t["filter"] != 0.0 or t["transmit"] != 0.0:
            pig = f"rgbft {_vec(t['pigment'] + (t['filter'], t['transmit']))}"
        else:
            pig = f"rgb {_vec(t['pigment'])}"
        fin = f"ambient {_f(t['ambient'])} diffuse {_f(t['diffuse'])}"
        if t["phong"] != 0.0:
            fin += f" phong {_f(t['phong'])} phong_size {_f(t['phong_size'])}"
        if t["specular"] != 0.0:
            fin += f" specular {_f(t['specular'])} roughness {_f(t['roughness'])}"
        if t["reflection"] != 0.0:
            fin += f" reflection {_f(t['reflection'])}"
        return f"texture {{ pigment {{ {pig} }} finish {{ {fin} }} }}"

    def to_pov(self, out):
        """Writes the scene as POV-Ray SDL to the text stream `out`."""
        g = self.g
        out.write(f"#version {g['version'] / 100.0:.1f};\n")
        out.write(f"global_settings {{ assumed_gamma {_f(g['assumed_gamma'])} max_trace_level {g['max_trace_level']} "
                  f"adc_bailout {_f(g['adc_bailout'])} ambient_light rgb {_vec(g['ambient_light'])} }}\n")
        out.write(f"background {{ rgb {_vec(g['background'])} }}\n")
        c = self.cam
        out.write(f"camera {{ perspective location {_vec(c['location'])} direction {_vec(c['direction'])} "
                  f"up {_vec(c['up'])} right {_vec(c['right'])} }}\n")
        for l in self.lights:
            out.write(f"light_source {{ {_vec(l['position'])} rgb {_vec(l['colour'])} }}\n")
        for o in self.objects:
            tex = self._texture_sdl(o["texture"])
            if o["kind"] == "sphere":
                out.write(f"sphere {{ {_vec(o['center'])}, {_f(o['radius'])} {tex} }}\n")
            elif o["kind"] == "plane":
                out.write(f"plane {{ {_vec(o['normal'])}, {_f(o['distance'])} {tex} }}\n")
            elif o["kind"] == "box":
                out.write(f"box {{ {_vec(o['c1'])}, {_vec(o['c2'])} {tex} }}\n")
            elif o["kind"] == "mesh2":
                v, f = o["vertices"], o["faces"]
                # %.17g round-trips every double, like repr(); vectorised because config 2 has ~1.5M lines
                out.write(f"mesh2 {{\n vertex_vectors {{ {len(v)}")
                out.write("".join(",\n<%.17g, %.17g, %.17g>" % (p[0], p[1], p[2]) for p in v.tolist()))
                out.write(f"\n }}\n face_indices {{ {len(f)}")
                out.write("".join(",\n<%d,%d,%d>" % (t[0], t[1], t[2]) for t in f.tolist()))
                out.write(f"\n }}\n {tex}\n}}\n")

    # ---- tables -------------------------------------------------------------------------------------
    def build(self):
        """Returns a (not yet finalized) `Scene` holding the same tables the adapter produces for `to_pov()`."""
        g = A.Globals()
        g.max_trace_level = self.g["max_trace_level"]
        g.language_version = self.g["version"]
        g.noise_generator = self.g["noise_generator"]
        g.quality_flags = A.Q_DEFAULT
        g.output_alpha = 0
        g.adc_bailout = self.g["adc_bailout"]
        for k in range(3):
            g.ambient_light[k] = self.g["ambient_light"][k]
            g.background[k] = self.g["background"][k]
        g.atmosphere_ior = 1.0
        g.atmosphere_dispersion = 1.0
        g.number_of_waves = 10
        # BoundingTask turns bounding off below Bounding_Threshold = 3 objects (boundingtask.cpp:170)
        g.bounding_method = 1 if len(self.objects) >= 3 else 0
        scene = Scene.create(g)
        lib = A.lib()

        n = len(self.objects)
        objs = (A.Object * n)()
        textures = (A.Texture * n)()
        pigments = (A.Pigment * n)()
        finishes = (A.Finish * n)()
        interiors = (A.Interior * n)()
        maps, entries = [], []
        meshes_pending = []
        for i, o in enumerate(self.objects):
            t = o["texture"]
            ob = objs[i]
            ob.texture, ob.interior_texture, ob.interior, ob.transform, ob.parent, ob.mesh = i, -1, i, -1, -1, -1
            opaque = (t["filter"] == 0.0 and t["transmit"] == 0.0)
            ob.flags = A.OPAQUE_FLAG if opaque else 0
            # material tables: one texture / pigment / finish / interior per object, in object order
            tx = textures[i]
            tx.type, tx.next, tx.pigment, tx.finish, tx.tnormal = A.PAT_PLAIN, -1, i, i, -1
            pg = pigments[i]
            pg.wave_type, pg.frequency, pg.phase, pg.exponent, pg.blend_map = A.WAVE_RAMP, 1.0, 0.0, 1.0, -1
            pg.quick_colour[0] = float("nan")
            if t["checker"] is not None:
                pg.pattern = A.PAT_CHECKER
                pg.blend_map = len(maps)
                maps.append((len(entries), 2))
                entries.append((0.0, t["checker"][0] + (0.0, 0.0)))
                entries.append((1.0, t["checker"][1] + (0.0, 0.0)))
            else:
                pg.pattern = A.PAT_PLAIN
                for k in range(3):
                    pg.colour[k] = t["pigment"][k]
                pg.colour[3], pg.colour[4] = t["filter"], t["transmit"]
            fn = finishes[i]
            fn.diffuse, fn.brilliance, fn.brilliance_adjust, fn.brilliance_adjust_rad = t["diffuse"], 1.0, 1.0, 1.0
            fn.specular, fn.roughness = t["specular"], (1.0 / t["roughness"]) if t["roughness"] != 0.0 else 0.0
            fn.phong, fn.phong_size, fn.reflect_exp, fn.reflection_falloff = t["phong"], t["phong_size"], 1.0, 1.0
            for k in range(3):
                fn.ambient[k] = t["ambient"]
                fn.reflection_max[k] = t["reflection"]
                fn.reflection_min[k] = t["reflection"]
            it = interiors[i]
            it.hollow, it.disp_nelems, it.ior, it.dispersion, it.old_refract = 0, 7, 1.0, 1.0, 1.0
            if o["kind"] == "sphere":
                ob.type = A.OBJ_SPHERE
                c, r = o["center"], o["radius"]
                ob.p[0], ob.p[1], ob.p[2], ob.p[3] = c[0], c[1], c[2], r
                for k in range(3):                       # Sphere::Compute_BBox (sphere.cpp:646-658)
                    ob.bbox[k] = c[k] - r
                    ob.bbox[3 + k] = 2.0 * r
            elif o["kind"] == "plane":
                ob.type = A.OBJ_PLANE
                nv = np.array(o["normal"], dtype=np.float64)
                nv = nv / math.sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2])      # Parse_Plane (parser.cpp:4992-4999)
                ob.p[0], ob.p[1], ob.p[2], ob.p[3] = nv[0], nv[1], nv[2], -o["distance"]
                for k in range(3):                       # Plane::Compute_BBox (plane.cpp:618-627)
                    ob.bbox[k] = -BOUND_HUGE / 2
                    ob.bbox[3 + k] = BOUND_HUGE
                ob.flags |= A.INFINITE_FLAG              # Post_Process: bounding volume > INFINITE_VOLUME (parser.cpp:9288-9291)
            elif o["kind"] == "box":
                ob.type = A.OBJ_BOX
                lo = [min(a, b) for a, b in zip(o["c1"], o["c2"])]      # Parse_Box orders the corners
                hi = [max(a, b) for a, b in zip(o["c1"], o["c2"])]
                for k in range(3):
                    ob.p[k], ob.p[3 + k] = lo[k], hi[k]
                    ob.bbox[k] = lo[k]                   # Box::Compute_BBox (box.cpp:947-957)
                    ob.bbox[3 + k] = hi[k] - lo[k]
            elif o["kind"] == "mesh2":
                ob.type = A.OBJ_MESH
                ob.flags |= A.HIERARCHY_FLAG
                meshes_pending.append(i)
            else:
                raise ValueError(o["kind"])

        for i in meshes_pending:
            o = self.objects[i]
            v, f = o["vertices"], o["faces"]
            mid = C.c_int32(-1)
            A.check(lib.pvgpu_scene_add_mesh2(scene.handle, v.ctypes.data_as(C.POINTER(C.c_double)), len(v),
                                              f.ctypes.data_as(C.POINTER(C.c_int32)), len(f), C.byref(mid)))
            objs[i].mesh = mid.value
            v32 = v.astype(np.float32).astype(np.float64)[np.unique(f)]      # Mesh::Compute_BBox over the FP32 vertices in use
            lo, hi = v32.min(axis=0), v32.max(axis=0)
            for k in range(3):
                objs[i].bbox[k] = lo[k]
                objs[i].bbox[3 + k] = hi[k] - lo[k]

        frame = (C.c_uint32 * n)(*range(n))
        A.check(lib.pvgpu_scene_set_objects(scene.handle, objs, n, None, 0, frame, n))
        marr = (A.BlendMap * max(1, len(maps)))()
        for k, (first, cnt) in enumerate(maps):
            marr[k].entry_first, marr[k].entry_count = first, cnt
        earr = (A.BlendEntry * max(1, len(entries)))()
        for k, (val, col) in enumerate(entries):
            earr[k].value = val
            for j in range(5):
                earr[k].colour[j] = col[j]
        A.check(lib.pvgpu_scene_set_materials(scene.handle, textures, n, pigments, n, finishes, n, marr, len(maps), earr, len(entries),
                                              None, 0, interiors, n))
        larr = (A.Light * max(1, len(self.lights)))()
        for k, l in enumerate(self.lights):
            L = larr[k]
            L.type, L.flags, L.projected_through = A.LIGHT_POINT, A.LIGHT_MEDIA_INTERACT, -1
            L.adaptive_level, L.object_flags = 100, A.NO_SHADOW_FLAG | A.INFINITE_FLAG      # LightSource() defaults
            for j in range(3):
                L.colour[j] = l["colour"][j]
                L.center[j] = l["position"][j]
            # LightSource defaults (lightsource.cpp): Points_At <0,0,1>, Axis1 <0,0,1>, Axis2 <0,1,0>, Direction = normalize(Points_At - Center)
            pa = np.array([0.0, 0.0, 1.0])
            d = pa - np.array(l["position"])
            d = d / math.sqrt((d * d).sum())
            for j in range(3):
                L.direction[j], L.points_at[j] = d[j], pa[j]
            L.axis1[2], L.axis2[1] = 1.0, 1.0
        A.check(lib.pvgpu_scene_set_lights(scene.handle, larr, len(self.lights)))
        cam = A.Camera()
        cam.type = A.CAMERA_PERSPECTIVE
        for j in range(3):
            cam.location[j] = self.cam["location"][j]
            cam.direction[j] = self.cam["direction"][j]
            cam.up[j] = self.cam["up"][j]
            cam.right[j] = self.cam["right"][j]
        scene.set_camera(cam)
        if g.bounding_method == 1:
            scene.build_tree()
        return scene


# ------------------------------------------------------------------------------------------------
# BASELINE.json configurations
# ------------------------------------------------------------------------------------------------
def spheres_scene(n_spheres=1024, seed=12345):
    """Config 1: `n_spheres` random spheres over a checker plane, one point light with shadows."""
    rng = np.random.RandomState(seed)
    b = SceneBuilder()
    b.camera((0.0, 6.0, -22.0), (0.0, 2.0, 0.0), angle=50.0)
    b.light((30.0, 40.0, -30.0))
    # plane slightly off y = 0 so that the checker's floor() never sits on a lattice plane
    b.plane((0.0, 1.0, 0.0), -0.0078125, b.texture((1, 1, 1), checker=((1, 1, 1), (0.15, 0.15, 0.15))))
    for _ in range(n_spheres):
        x = rng.uniform(-16.0, 16.0)
        z = rng.uniform(-8.0, 24.0)
        r = rng.uniform(0.15, 0.45)
        y = r + rng.uniform(0.0, 6.0)
        col = rng.uniform(0.1, 1.0, size=3)
        b.sphere((x, y, z), r, b.texture(col, ambient=0.1, diffuse=0.7, phong=0.5))
    return b


def heightfield_mesh(grid=708, size=4.0):
    """Vertex / face arrays of the config-2 tessellated height field: (grid-1)^2 * 2 triangles."""
    xs = np.linspace(-size, size, grid)
    X, Z = np.meshgrid(xs, xs, indexing="xy")
    Y = 0.6 * np.sin(1.7 * X) * np.cos(1.3 * Z) + 0.25 * np.sin(4.1 * X + 1.0) * np.sin(3.7 * Z) + 0.1 * np.cos(9.0 * X * Z / 4.0)
    verts = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    idx = np.arange(grid * grid, dtype=np.int32).reshape(grid, grid)
    a, bq, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    faces = np.concatenate([np.stack([a, bq, d], axis=1), np.stack([a, d, c], axis=1)], axis=0).astype(np.int32)
    return verts, faces


def mesh_scene(grid=708):
    """Config 2: ~1M-triangle mesh2 (grid 708 -> 999,698 triangles) with its BBox tree, 2 lights, reflection 0.3,
    max_trace_level 5, plus a floor plane and two mirror spheres so the top-level tree exists (>= 3 objects)."""
    b = SceneBuilder(max_trace_level=5)
    b.camera((0.0, 5.5, -9.5), (0.0, 0.0, 0.0), angle=50.0)
    b.light((12.0, 18.0, -14.0), (0.9, 0.9, 0.9))
    b.light((-15.0, 12.0, -6.0), (0.5, 0.5, 0.6))
    v, f = heightfield_mesh(grid)
    b.mesh2(v, f, b.texture((0.55, 0.7, 0.9), ambient=0.1, diffuse=0.6, phong=0.4, reflection=0.3))
    b.plane((0.0, 1.0, 0.0), -1.5078125, b.texture((1, 1, 1), checker=((0.9, 0.9, 0.9), (0.2, 0.25, 0.3)), reflection=0.15))
    b.sphere((-2.5, 2.2, 1.0), 0.9, b.texture((0.9, 0.3, 0.2), diffuse=0.5, specular=0.6, roughness=0.02, reflection=0.4))
    b.sphere((2.6, 2.0, -0.5), 0.8, b.texture((0.2, 0.8, 0.3), diffuse=0.5, specular=0.6, roughness=0.02, reflection=0.4))
    return b


# ------------------------------------------------------------------------------------------------
# Configs 3 and 4 as SDL text only: these scenes use constructs (CSG, quadrics, tori, interiors, noise pigments with
# colour maps) whose tables come from the reference parser through the adapter, not from build().
# ------------------------------------------------------------------------------------------------
def csg_scene_pov(n_objects=4096, seed=777, max_trace_level=6):
    """Config 3: `n_objects` CSG objects, each a difference / intersection / merge of 2-4 of {box, sphere, quadric
    (cylinder, cone, ellipsoid forms)}, glass-like (filter 0.7, ior 1.3-1.6), on a jittered grid over a checker floor."""
    rng = np.random.RandomState(seed)
    side = int(math.ceil(math.sqrt(n_objects)))
    out = ["#version 3.7;",
           f"global_settings {{ assumed_gamma 1 max_trace_level {max_trace_level} }}",
           "background { rgb <0.05, 0.07, 0.12> }",
           "camera { perspective location <0, 14, -30> direction <0, 0, 1.6> up <0, 1, 0> right <1.7777777777777777, 0, 0> look_at <0, 0, 2> }",
           "light_source { <40, 60, -40> rgb <1, 1, 1> }",
           "light_source { <-30, 25, -10> rgb <0.4, 0.4, 0.5> }",
           "plane { y, -0.5078125 pigment { checker rgb <0.9, 0.9, 0.9>, rgb <0.25, 0.3, 0.35> } finish { ambient 0.1 diffuse 0.7 } }"]

    def prim(kind, c, s):
        x, y, z = c
        if kind == 0:
            return f"box {{ <{x - s:.6f}, {y - s:.6f}, {z - s:.6f}>, <{x + s:.6f}, {y + s:.6f}, {z + s:.6f}> }}"
        if kind == 1:
            return f"sphere {{ <{x:.6f}, {y:.6f}, {z:.6f}>, {s * 1.2:.6f} }}"
        if kind == 2:      # cylinder along y as a quadric: x^2 + z^2 = r^2
            return f"quadric {{ <1, 0, 1>, <0, 0, 0>, <0, 0, 0>, {-(s * 0.8) ** 2:.6f} translate <{x:.6f}, {y:.6f}, {z:.6f}> }}"
        if kind == 3:      # ellipsoid as a quadric
            return f"quadric {{ <1, 2.5, 1.5>, <0, 0, 0>, <0, 0, 0>, {-(s * 1.1) ** 2:.6f} translate <{x:.6f}, {y:.6f}, {z:.6f}> }}"
        return f"quadric {{ <1, -0.6, 1>, <0, 0, 0>, <0, 0, 0>, 0 translate <{x:.6f}, {y + s:.6f}, {z:.6f}> }}"      # cone

    n = 0
    for iz in range(side):
        for ix in range(side):
            if n >= n_objects:
                break
            n += 1
            cx = (ix - side / 2.0) * 1.1 + rng.uniform(-0.15, 0.15)
            cz = (iz - side / 2.0) * 1.1 + 8.0 + rng.uniform(-0.15, 0.15)
            s = rng.uniform(0.28, 0.42)
            cy = s - 0.3 + rng.uniform(0.0, 0.4)
            op = ("difference", "intersection", "merge")[rng.randint(0, 3)]
            k = rng.randint(2, 5)
            # every object stays finite: a merge gets finite children only (box, sphere, ellipsoid), the first child of a
            # difference / intersection is a box or a sphere (its box bounds the whole object, csg.cpp Compute_BBox)
            finite = (0, 1, 3)
            kinds = [rng.randint(0, 2)] + [(finite[rng.randint(0, 3)] if op == "merge" else rng.randint(0, 5)) for _ in range(k - 1)]
            parts = []
            for j, kd in enumerate(kinds):
                off = (0.0, 0.0, 0.0) if j == 0 else tuple(rng.uniform(-0.5, 0.5) * s for _ in range(3))
                parts.append(prim(kd, (cx + off[0], cy + off[1], cz + off[2]), s * (1.0 if j == 0 else rng.uniform(0.6, 0.95))))
            col = rng.uniform(0.3, 1.0, size=3)
            ior = rng.uniform(1.3, 1.6)
            body = " ".join(parts)
            bound = f" bounded_by {{ sphere {{ <{cx:.6f}, {cy:.6f}, {cz:.6f}>, {s * 2.6:.6f} }} }}" if op != "merge" and rng.rand() < 0.3 else ""
            out.append(f"{op} {{ {body}{bound} pigment {{ rgbf <{col[0]:.4f}, {col[1]:.4f}, {col[2]:.4f}, 0.7> }} "
                       f"finish {{ ambient 0.05 diffuse 0.5 specular 0.4 roughness 0.03 reflection 0.08 }} interior {{ ior {ior:.4f} }} }}")
    return "\n".join(out) + "\n"


def torus_scene_pov(n_tori=2048, seed=4242, n_blobs=64):
    """Config 4: `n_tori` tori (half of them `sturm`), random rotation / translation, granite and bozo pigments with
    colour maps and turbulence, noise_generator 2 for one half of the pigments and 3 for the other; plus `n_blobs` blobs of
    8 components each (spheres and cylinders, every fourth blob `sturm`)."""
    rng = np.random.RandomState(seed)
    out = ["#version 3.7;",
           "global_settings { assumed_gamma 1 max_trace_level 5 noise_generator 2 }",
           "background { rgb <0.1, 0.12, 0.18> }",
           "camera { perspective location <0, 18, -34> direction <0, 0, 1.5> up <0, 1, 0> right <1.7777777777777777, 0, 0> look_at <0, 2, 4> }",
           "light_source { <30, 50, -40> rgb <1, 1, 1> }",
           "plane { y, -0.2578125 pigment { checker rgb <0.8, 0.8, 0.8>, rgb <0.3, 0.3, 0.3> } finish { ambient 0.1 diffuse 0.7 } }"]
    for i in range(n_tori):
        R, r = rng.uniform(0.5, 1.0), rng.uniform(0.1, 0.3)
        rot = rng.uniform(0.0, 360.0, size=3)
        pos = (rng.uniform(-22.0, 22.0), rng.uniform(0.4, 9.0), rng.uniform(-10.0, 34.0))
        c1, c2, c3 = rng.uniform(0.05, 1.0, size=(3, 3))
        pat = "granite" if i % 2 == 0 else "bozo"
        gen = "" if i % 4 < 2 else " noise_generator 3"
        turb = f" turbulence {rng.uniform(0.1, 0.6):.3f}" if i % 3 == 0 else ""
        sturm = " sturm" if i % 2 == 1 else ""
        out.append(f"torus {{ {R:.6f}, {r:.6f}{sturm} pigment {{ {pat}{gen}{turb} scale {rng.uniform(0.15, 0.6):.4f} color_map {{ "
                   f"[0 rgb <{c1[0]:.4f}, {c1[1]:.4f}, {c1[2]:.4f}>] [0.5 rgb <{c2[0]:.4f}, {c2[1]:.4f}, {c2[2]:.4f}>] "
                   f"[1 rgb <{c3[0]:.4f}, {c3[1]:.4f}, {c3[2]:.4f}>] }} }} finish {{ ambient 0.1 diffuse 0.65 phong 0.4 }} "
                   f"rotate <{rot[0]:.4f}, {rot[1]:.4f}, {rot[2]:.4f}> translate <{pos[0]:.6f}, {pos[1]:.6f}, {pos[2]:.6f}> }}")
    for i in range(n_blobs):
        cx, cy, cz = rng.uniform(-20.0, 20.0), rng.uniform(1.0, 6.0), rng.uniform(-8.0, 30.0)
        comps = []
        for k in range(8):
            px, py, pz = cx + rng.uniform(-1.0, 1.0), cy + rng.uniform(-0.8, 0.8), cz + rng.uniform(-1.0, 1.0)
            if k % 3 == 2:
                comps.append(f"cylinder {{ <{px:.5f}, {py:.5f}, {pz:.5f}>, <{px + rng.uniform(-0.9, 0.9):.5f}, {py + rng.uniform(0.3, 0.9):.5f}, "
                             f"{pz + rng.uniform(-0.9, 0.9):.5f}>, {rng.uniform(0.4, 0.7):.5f}, {rng.uniform(0.7, 1.2):.4f} }}")
            else:
                comps.append(f"sphere {{ <{px:.5f}, {py:.5f}, {pz:.5f}>, {rng.uniform(0.7, 1.3):.5f}, {rng.uniform(0.7, 1.2):.4f} }}")
        col = rng.uniform(0.2, 1.0, size=3)
        out.append(f"blob {{ threshold {rng.uniform(0.45, 0.65):.4f} {' '.join(comps)}{' sturm' if i % 4 == 0 else ''} "
                   f"pigment {{ rgb <{col[0]:.4f}, {col[1]:.4f}, {col[2]:.4f}> }} finish {{ ambient 0.1 diffuse 0.65 specular 0.4 roughness 0.03 }} }}")
    return "\n".join(out) + "\n"
