#!/usr/bin/env python3
"""Known-answer vectors for Solve_Polynomial (polynomialsolver.cpp:1585-1729) and Noise / DNoise / Turbulence
(noise.h:196-202) computed by the UNMODIFIED reference functions: the adapter (oracle/_ref/parity/povray-gpu) runs its
probe hook (PVGPU_PROBE_IN / PVGPU_PROBE_OUT) during a tiny render.  Outputs tests/golden/probe.in and probe.out.
Only runs inside the build container; the outputs are committed.
usage: python tests/golden/make_golden_probe.py"""
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
ADAPTER = os.path.join(ROOT, "oracle", "_ref", "parity", "povray-gpu")
rng = np.random.RandomState(20261017)

polys = []      # (degree, sturm, epsilon, c[5] with leading zeros for degree < 4)
def add(deg, sturm, eps, coeffs):
    c = [0.0] * (4 - deg) + [float(x) for x in coeffs]
    polys.append((deg, sturm, eps, c))

for i in range(3000):
    kind = i % 6
    sturm = (i // 6) % 2
    if kind == 0:      # quartics from four real roots
        r = rng.uniform(-5, 5, 4); c = np.poly(r) * rng.uniform(0.2, 3.0)
        add(4, sturm, 1e-4 if i % 3 else 0.0, c)
    elif kind == 1:    # generic quartics (torus-like magnitudes)
        add(4, sturm, 1e-4, [1.0] + list(rng.uniform(-40, 40, 4)))
    elif kind == 2:    # wide coefficient ranges: trigger the automatic Sturm fallback (difficult_coeffs)
        add(4, 0, 1e-4, [1.0] + list(rng.uniform(-1, 1, 4) * 10.0 ** rng.uniform(-7, 8, 4)))
    elif kind == 3:    # blob-like quartics, tiny epsilon
        add(4, sturm, 1e-11, list(rng.uniform(-3, 3, 5)))
    elif kind == 4:    # cubics and quadratics
        add(3, sturm, 0.0, list(rng.uniform(-10, 10, 4)))
        add(2, sturm, 0.0, list(rng.uniform(-10, 10, 3)))
    else:              # leading coefficients that vanish / double roots
        r = rng.uniform(-3, 3, 2); c = np.poly([r[0], r[0], r[1], r[1]])
        add(4, sturm, 1e-4, c)
        add(4, sturm, 1e-4, [0.0, 0.0] + list(rng.uniform(-5, 5, 3)))

pts = []
for i in range(4096):
    scale = 10.0 ** rng.uniform(-1, 3)
    p = rng.uniform(-1, 1, 3) * scale
    if i % 64 == 0:
        p = np.round(p)                     # lattice points
    pts.append((p[0], p[1], p[2], 1 + i % 3, 1 + i % 8))

with open(os.path.join(HERE, "probe.in"), "wb") as f:
    f.write(struct.pack("<I", len(polys)))
    for deg, sturm, eps, c in polys:
        f.write(struct.pack("<iid5d", deg, sturm, eps, *c))
    f.write(struct.pack("<I", len(pts)))
    for x, y, z, gen, octv in pts:
        f.write(struct.pack("<3dii", x, y, z, gen, octv))
out = os.path.join(HERE, "probe.out")
if os.path.exists(out):
    os.remove(out)
env = dict(os.environ, PVGPU_RENDER="stock", PVGPU_PROBE_IN=os.path.join(HERE, "probe.in"), PVGPU_PROBE_OUT=out)
r = subprocess.run([ADAPTER, "+I" + os.path.join(HERE, "scenes", "spheres64.pov"), "+O/tmp/probe.png", "+W8", "+H8", "-A", "-D", "+WT1", "-GA"],
                   env=env, capture_output=True, text=True, cwd="/tmp")
assert r.returncode == 0 and os.path.exists(out), r.stdout[-2000:] + r.stderr[-2000:]
print(len(polys), "polynomials,", len(pts), "noise points ->", os.path.getsize(out), "bytes")
