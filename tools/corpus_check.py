#!/usr/bin/env python3
"""Runs the reference's own scene corpus (distribution/scenes/**/*.pov) through the reference-side adapter in `stock` mode:
which scenes does the flattener accept for the GPU trace path, which feature makes it reject the others, and - for accepted
scenes - does the CPU restatement (oracle) reproduce the reference's pixels.  Build-container tool (needs /root/reference and
oracle/_ref/parity/povray-gpu); writes a JSON summary.

usage: python tools/corpus_check.py [--out profiles/r1_corpus.json] [--jobs 8] [subdir ...]
"""
import argparse
import concurrent.futures as cf
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
ADAPTER = os.path.join(ROOT, "oracle", "_ref", "parity", "povray-gpu")
SCENES = "/root/reference/distribution/scenes"
INC = "/root/reference/distribution/include"
W, H = 64, 48


def one(pov):
    rel = os.path.relpath(pov, SCENES)
    with tempfile.TemporaryDirectory() as d:
        env = dict(os.environ, PVGPU_RENDER="stock", PVGPU_DUMP_SCENE=os.path.join(d, "s.pvs"), PVGPU_DUMP_RGBT=os.path.join(d, "s.rgbt"))
        try:
            r = subprocess.run([ADAPTER, "+I" + pov, "+O" + os.path.join(d, "o.png"), f"+W{W}", f"+H{H}", "-A", "-D", "+WT1", "-GA", "+L" + INC,
                                "+L" + os.path.dirname(pov)], env=env, capture_output=True, text=True, timeout=120, cwd=os.path.dirname(pov))
        except subprocess.TimeoutExpired:
            return rel, dict(status="timeout")
        m = re.search(r"feature outside the GPU trace path: (.*)", r.stderr)
        if r.returncode != 0 and not m:
            return rel, dict(status="reference_error", detail=r.stderr.strip().splitlines()[-1][:200] if r.stderr.strip() else "")
        if m:
            return rel, dict(status="rejected", reason=m.group(1).strip())
        if not os.path.exists(os.path.join(d, "s.rgbt")):
            return rel, dict(status="no_dump")
        # the library's own validation (pvgpu_scene_finalize validates before it looks for a device)
        from povray_b200 import _abi as A
        import ctypes as C
        h = A.VP()
        if A.lib().pvgpu_scene_load(C.byref(h), os.path.join(d, "s.pvs").encode()) != 0:
            return rel, dict(status="load_error")
        rc = A.lib().pvgpu_scene_finalize(h, 0)
        msg = A.lib().pvgpu_last_error().decode()
        A.lib().pvgpu_scene_destroy(h)
        if rc not in (A.OK, A.E_NO_DEVICE):
            return rel, dict(status="rejected", reason="library: " + re.sub(r"\b(object|finish|light|texture|pigment|interior|mesh|tnormal|warp|fog) \d+", r"\1 N", msg))
        import oracle_lib
        ref = np.fromfile(os.path.join(d, "s.rgbt"), dtype=np.float32).reshape(H, W, 4)
        try:
            img, _ = oracle_lib.OracleScene(os.path.join(d, "s.pvs")).render(W, H, threads=1)
        except Exception as e:
            return rel, dict(status="oracle_error", detail=str(e)[:200])
        diff = np.abs(img - ref).max(axis=2)
        save = os.environ.get("PVGPU_CORPUS_SAVE")
        if save:        # keep the flattened scene + the reference's pixels as a fixture (tests/golden/corpus, test_gpu_parity.py)
            import shutil
            stem = os.path.join(save, rel[:-4].replace("/", "__"))
            shutil.copy(os.path.join(d, "s.pvs"), stem + ".pvs")
            shutil.copy(os.path.join(d, "s.rgbt"), stem + ".rgbt")
        return rel, dict(status="accepted", max_abs=float(diff.max()), frac_within_1_255=float((diff <= 1 / 255).mean()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r1_corpus.json"))
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("subdirs", nargs="*")
    a = ap.parse_args()
    povs = []
    for base, _, files in os.walk(SCENES):
        if a.subdirs and not any(os.path.relpath(base, SCENES).startswith(s) for s in a.subdirs):
            continue
        povs += [os.path.join(base, f) for f in files if f.endswith(".pov")]
    povs.sort()
    res = {}
    with cf.ProcessPoolExecutor(a.jobs) as ex:
        for rel, r in ex.map(one, povs):
            res[rel] = r
            print(rel, r, flush=True)
    by = {}
    for r in res.values():
        by[r["status"]] = by.get(r["status"], 0) + 1
    reasons = {}
    for r in res.values():
        if r["status"] == "rejected":
            reasons[r["reason"]] = reasons.get(r["reason"], 0) + 1
    acc = [r for r in res.values() if r["status"] == "accepted"]
    summary = dict(scenes=len(res), by_status=by, rejection_reasons=dict(sorted(reasons.items(), key=lambda kv: -kv[1])),
                   accepted_within_contract=sum(1 for r in acc if r["frac_within_1_255"] >= 0.999), accepted=len(acc), width=W, height=H)
    json.dump(dict(summary=summary, scenes=res), open(a.out, "w"), indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
