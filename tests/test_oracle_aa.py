"""Pins the oracle's anti-aliasing (sampling methods 1 and 2: tracetask.cpp:521-657, 838-1074; Jitter2d jitter.h:92)
against 16-bit linear PPM renders of the UNMODIFIED reference binary (tests/golden/make_golden_aa.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_W as W, GOLDEN_H as H

AA = os.path.join(GOLDEN, "aa")
MODES = {"m1_jitter": (1, 3, 0.3, 1.0), "m1_nojitter": (1, 3, 0.3, 0.0), "m1_r2": (1, 2, 0.1, 1.0),
         "m2_jitter": (2, 3, 0.3, 1.0), "m2_nojitter": (2, 3, 0.3, 0.0), "m2_r2": (2, 2, 0.1, 1.0)}
SCENES = ["spheres64", "csg_glass", "torus_noise"]


def read_ppm16(path):
    data = open(path, "rb").read()
    toks, pos = [], 0
    while len(toks) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        toks.append(data[pos:end]); pos = end
    pos += 1
    assert toks[0] == b"P6" and int(toks[3]) == 65535
    return np.frombuffer(data[pos:], dtype=">u2").reshape(int(toks[2]), int(toks[1]), 3).astype(np.float64) / 65535.0


def tiles32(w, h):
    return [(x, y, min(x + 32, w) - 1, min(y + 32, h) - 1) for y in range(0, h, 32) for x in range(0, w, 32)]


def assemble(px, rects, w, h):
    img = np.zeros((h, w, 4), dtype=np.float32)
    pos = 0
    for l, t, r, b in rects:
        n = (r - l + 1) * (b - t + 1)
        img[t:b + 1, l:r + 1] = px[pos:pos + n].reshape(b - t + 1, r - l + 1, 4)
        pos += n
    return img


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("scene", SCENES)
def test_oracle_aa_matches_reference_image(oracle, scene, mode):
    method, depth, thr, jit = MODES[mode]
    ref = read_ppm16(os.path.join(AA, f"{scene}_{mode}.ppm"))
    counters = json.load(open(os.path.join(AA, "counters.json")))[f"{scene}_{mode}"]
    o = oracle.OracleScene(os.path.join(GOLDEN, scene + ".pvs"))
    rects = tiles32(W, H)
    px, st = oracle.render_aa(o, W, H, rects, method, depth, thr, jit, 2.5, threads=4)
    img = np.clip(assemble(px, rects, W, H)[..., :3].astype(np.float64), 0.0, 1.0)
    d = np.abs(img - ref).max(axis=2)
    # 16-bit quantisation of the reference file is 1.5e-5; a differing supersampling decision would show as >= 1e-3
    assert (d > 3e-5).mean() <= 0.001, f"{(d > 3e-5).sum()} pixels differ (max {d.max():.2e})"
    assert st["samples"] == counters["samples"]
    assert st["rays"] == counters["rays"]
