#!/usr/bin/env python3
"""Full-size goldens for BASELINE.json configs 1 and 2: the UNMODIFIED reference's TracePixel (adapter in `stock` mode)
renders the synthetic scenes at 1920x1080; kept is the float RGBT of the pixel lattice x % 8 == 3, y % 8 == 5
(240 x 135 pixels, 518 KB per scene) and, for configs 3 and 4, the ray records (first-hit object, depth) of the same pixels.  Also checks that the Python scene builder's tables are byte-identical to what the
reference parser produced.  Only runs inside the build container (needs oracle/_ref/parity); ~2 minutes.
usage: python tests/golden/make_golden_1080.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from povray_b200 import synth

W, H = 1920, 1080
ADAPTER = os.path.join(ROOT, "oracle", "_ref", "parity", "povray-gpu")
with tempfile.TemporaryDirectory() as d:
    for name, b in (("cfg1", synth.spheres_scene(1024)), ("cfg2", synth.mesh_scene(708))):
        pov = os.path.join(d, name + ".pov")
        with open(pov, "w") as f:
            b.to_pov(f)
        full = os.path.join(d, name + ".rgbt")
        env = dict(os.environ, PVGPU_RENDER="stock", PVGPU_DUMP_SCENE=os.path.join(d, name + "_ref.pvs"), PVGPU_DUMP_RGBT=full)
        r = subprocess.run([ADAPTER, "+I" + pov, "+O" + os.path.join(d, name + ".png"), f"+W{W}", f"+H{H}", "-A", "-D", "+WT1", "-GA"],
                           env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        s = b.build()
        s.save(os.path.join(d, name + "_py.pvs"))
        same = open(os.path.join(d, name + "_py.pvs"), "rb").read() == open(os.path.join(d, name + "_ref.pvs"), "rb").read()
        img = np.fromfile(full, dtype="<f4").reshape(H, W, 4)
        img[5::8, 3::8].astype("<f4").tofile(os.path.join(HERE, f"{name}_1080_lattice8.rgbt"))
        print(name, "tables identical to the parser's:", same)
    # configs 3 and 4 (scene text from the generators, tables from the reference parser): float RGBT of the same lattice plus the
    # reference's own camera ray / first-hit object / depth records (Trace::FindIntersection) of those pixels
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    for name, text in (("cfg3", synth.csg_scene_pov(4096)), ("cfg4", synth.torus_scene_pov(2048))):
        pov = os.path.join(d, name + ".pov")
        open(pov, "w").write(text)
        full, rays = os.path.join(d, name + ".rgbt"), os.path.join(d, name + ".rays")
        env = dict(os.environ, PVGPU_RENDER="stock", PVGPU_DUMP_RGBT=full, PVGPU_DUMP_RAYS=rays)
        r = subprocess.run([ADAPTER, "+I" + pov, "+O" + os.path.join(d, name + ".png"), f"+W{W}", f"+H{H}", "-A", "-D", "+WT1", "-GA"],
                           env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        img = np.fromfile(full, dtype="<f4").reshape(H, W, 4)
        img[5::8, 3::8].astype("<f4").tofile(os.path.join(HERE, f"{name}_1080_lattice8.rgbt"))
        rec = np.fromfile(rays, dtype=oracle_lib.RAY_DTYPE).reshape(H, W)
        rec[5::8, 3::8].copy().tofile(os.path.join(HERE, f"{name}_1080_lattice8.rays"))
        print(name, "lattice written:", (rec[5::8, 3::8]["obj"] >= 0).mean(), "of the lattice rays hit something")
