#!/usr/bin/env python3
"""Wall-clock of the whole program, reference front end included: `povray-gpu` (reference parser, bounding, flattening, upload,
GPU trace path, PNG output) next to the unmodified `povray` binary with +WT<host threads> on the same .pov, 1920x1080.
Parse and bounding stay on the CPU in both, so this puts the floor that the CPU side of the seam sets on the record.

usage (GPU box): python tools/e2e_adapter.py [--workloads cfg1,cfg2,cfg3_noaa,cfg4] [--devices N] > gpurun_out/e2e.json
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def run(binary, pov, d, threads, env=None):
    t0 = time.perf_counter()
    r = subprocess.run([binary, "+I" + pov, "+O" + os.path.join(d, "o.png"), "+W1920", "+H1080", "-A", "-D", f"+WT{threads}",
                        "-GD", "-GR", "-GW", "-GF", "+GS"], capture_output=True, text=True, cwd=d, env=dict(os.environ, **(env or {})))
    wall = time.perf_counter() - t0
    out = (r.stdout + r.stderr).replace("\r", "\n")
    if r.returncode != 0:
        return {"error": out[-400:]}

    def sec(name):
        m = re.search(name + r" Time:.*?\(([\d.]+) seconds\)", out)
        return float(m.group(1)) if m else None
    return {"wall_s": round(wall, 3), "parse_s": sec("Parse"), "bounding_s": sec("Bounding"), "trace_s": sec("Trace")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="cfg1,cfg2,cfg3_noaa,cfg4")
    ap.add_argument("--devices", default="")
    args = ap.parse_args()
    threads = os.cpu_count() or 1
    ref = os.path.join(ROOT, "oracle", "_ref", "fast", "povray")
    out = {"host_threads": threads, "frame": "1920x1080 -A", "rows": {}}
    with tempfile.TemporaryDirectory() as d:
        for wl in args.workloads.split(","):
            pov = bench.write_pov(wl, d)
            env = {"PVGPU_RENDER": "gpu"}
            if args.devices:
                env["PVGPU_DEVICES"] = args.devices
            fast_adapter = os.path.join(ROOT, "oracle", "_ref", "fast", "povray-gpu")       # same -O3 front end as the CPU binary
            row = {"povray_gpu": run(fast_adapter if os.path.exists(fast_adapter) else bench.ADAPTER, pov, d, threads, env),
                   "povray_cpu": run(ref, pov, d, threads)}
            out["rows"][wl] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
