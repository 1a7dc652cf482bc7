#!/usr/bin/env python3
"""Generates the golden vectors of tests/golden/ by running the UNMODIFIED reference (oracle/_ref/parity,
built by oracle/build_ref.sh) through the reference-side adapter in `stock` mode:

  <name>.pvs   flattened scene exactly as the reference parser + BoundingTask produced it
  <name>.rays  per pixel: the reference's camera ray and Trace::FindIntersection result (object, depth, aux)
  <name>.rgbt  per pixel: float RGBT from the reference's TracePixel

Only runs inside the build container (needs /root/reference to have been built); the outputs are committed.
usage: python tests/golden/make_golden.py [name ...]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
ADAPTER = os.path.join(ROOT, "oracle", "_ref", "parity", "povray-gpu")
W, H = 96, 54


def synthetic():
    from povray_b200 import synth
    return {"spheres64": synth.spheres_scene(64), "mesh24": synth.mesh_scene(24)}


def main(names):
    os.makedirs(os.path.join(HERE, "scenes"), exist_ok=True)
    for name, b in synthetic().items():
        with open(os.path.join(HERE, "scenes", name + ".pov"), "w") as f:
            b.to_pov(f)
    scenes = sorted(f[:-4] for f in os.listdir(os.path.join(HERE, "scenes")) if f.endswith(".pov"))
    for name in scenes:
        if names and name not in names:
            continue
        pov = os.path.join(HERE, "scenes", name + ".pov")
        env = dict(os.environ, PVGPU_RENDER="stock", PVGPU_DUMP_SCENE=os.path.join(HERE, name + ".pvs"),
                   PVGPU_DUMP_RAYS=os.path.join(HERE, name + ".rays"), PVGPU_DUMP_RGBT=os.path.join(HERE, name + ".rgbt"))
        r = subprocess.run([ADAPTER, "+I" + pov, "+O/tmp/golden_" + name + ".png", f"+W{W}", f"+H{H}", "-A", "-D", "+WT1", "-GA"],
                           env=env, capture_output=True, text=True, cwd=os.path.join(HERE, "scenes"))       # image files live next to the scenes
        if r.returncode != 0:
            print(name, "FAILED\n", r.stdout[-3000:], r.stderr[-3000:])
            sys.exit(1)
        print(name, "ok", {e: os.path.getsize(os.path.join(HERE, name + e)) for e in (".pvs", ".rays", ".rgbt")})


if __name__ == "__main__":
    main(sys.argv[1:])
