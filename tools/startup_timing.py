#!/usr/bin/env python3
"""Where a one-shot render spends its wall clock on the library side: load, finalize (context + upload), first frame (module load,
work buffers), second frame.  usage (GPU box): python tools/startup_timing.py [cfg1|cfg2]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t0 = time.perf_counter()
import povray_b200 as pv
from povray_b200 import synth, _abi as A
A.lib()
t1 = time.perf_counter()
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
s = (synth.spheres_scene(1024) if wl == "cfg1" else synth.mesh_scene(708)).build()
t2 = time.perf_counter()
s.finalize(0)
t3 = time.perf_counter()
hb = pv.HostBuffer(1920 * 1080 * 4)
t4 = time.perf_counter()
rects = pv.tiles(1920, 1080)
s.render(1920, 1080, rects, out=hb.array)
t5 = time.perf_counter()
s.render(1920, 1080, rects, out=hb.array)
t6 = time.perf_counter()
print(f"{wl}: import+dlopen {t1 - t0:.3f} s, host scene build {t2 - t1:.3f} s, finalize (context + upload) {t3 - t2:.3f} s, pinned frame alloc {t4 - t3:.3f} s, "
      f"first frame {t5 - t4:.3f} s, second frame {t6 - t5:.3f} s")
