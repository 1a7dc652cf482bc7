#!/usr/bin/env python3
"""Anti-aliasing goldens: the UNMODIFIED reference binary (oracle/_ref/parity/povray) renders the golden scenes with
+A0.3 +R3 and sampling methods 1 and 2 to 16-bit linear PPM (File_Gamma=1.0, no dithering), default 32x32 render
blocks.  Outputs tests/golden/aa/<scene>_<mode>.ppm plus the ray / sample counters in tests/golden/aa/counters.json.
Only runs inside the build container; the outputs are committed.
usage: python tests/golden/make_golden_aa.py"""
import json
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "parity", "povray")
W, H = 96, 54
MODES = {"m1_jitter": ["+A0.3", "+AM1", "+R3", "+J"], "m1_nojitter": ["+A0.3", "+AM1", "+R3", "-J"], "m1_r2": ["+A0.1", "+AM1", "+R2", "+J"],
         "m2_jitter": ["+A0.3", "+AM2", "+R3", "+J"], "m2_nojitter": ["+A0.3", "+AM2", "+R3", "-J"], "m2_r2": ["+A0.1", "+AM2", "+R2", "+J"]}
SCENES = ["spheres64", "csg_glass", "torus_noise"]

os.makedirs(os.path.join(HERE, "aa"), exist_ok=True)
counters = {}
for scene in SCENES:
    for mode, flags in MODES.items():
        out = os.path.join(HERE, "aa", f"{scene}_{mode}.ppm")
        r = subprocess.run([REF, "+I" + os.path.join(HERE, "scenes", scene + ".pov"), "+O" + out, "+FP16", "File_Gamma=1.0", f"+W{W}", f"+H{H}",
                            "-D", "+WT1", "-GD", "-GR", "-GW", "-GF", "+GS"] + flags, capture_output=True, text=True, cwd="/tmp")
        txt = (r.stdout + r.stderr).replace("\r", "\n")
        assert r.returncode == 0, txt[-2000:]
        counters[f"{scene}_{mode}"] = {"pixels": int(re.search(r"Pixels:\s+(\d+)", txt).group(1)), "samples": int(re.search(r"Samples:\s+(\d+)", txt).group(1)),
                                       "rays": int(re.search(r"Rays:\s+(\d+)", txt).group(1))}
        print(scene, mode, counters[f"{scene}_{mode}"])
json.dump(counters, open(os.path.join(HERE, "aa", "counters.json"), "w"), indent=1, sort_keys=True)
