"""Pins the oracle (oracle/pvoracle.cpp, the CPU restatement) against golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py): ray-level object id + depth (bit-exact demanded here, the contract is
1e-9 relative) and float RGBT per pixel."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_W as W, GOLDEN_H as H, GOLDEN_SCENES


def load_golden(name):
    import oracle_lib
    rays = np.fromfile(os.path.join(GOLDEN, name + ".rays"), dtype=oracle_lib.RAY_DTYPE)
    rgbt = np.fromfile(os.path.join(GOLDEN, name + ".rgbt"), dtype="<f4").reshape(H, W, 4)
    return rays, rgbt


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_oracle_first_hits_match_reference(oracle, name):
    rays, _ = load_golden(name)
    o = oracle.OracleScene(os.path.join(GOLDEN, name + ".pvs"))
    od = np.concatenate([rays["org"], rays["dir"]], axis=1)
    obj, depth, aux = o.trace_rays(od)
    assert np.array_equal(obj, rays["obj"]), f"{(obj != rays['obj']).sum()} first-hit object ids differ"
    hit = rays["obj"] >= 0
    assert hit.sum() > 1000
    rel = np.abs(depth[hit] - rays["depth"][hit]) / rays["depth"][hit]
    assert rel.max() <= 1e-9
    assert (rel == 0).mean() > 0.999          # in practice every depth is bit-identical
    cmp = hit & (rays["aux"] != -1)          # (-1: a glyph hit - the reference stores a normal with it, not an index)
    assert np.array_equal(aux[cmp], rays["aux"][cmp].astype(np.uint32))


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_oracle_camera_rays_are_bit_exact(oracle, name):
    rays, _ = load_golden(name)
    o = oracle.OracleScene(os.path.join(GOLDEN, name + ".pvs"))
    xy = np.stack(np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5), axis=-1).reshape(-1, 2)
    od = o.camera_rays(W, H, xy)
    assert np.array_equal(od, np.concatenate([rays["org"], rays["dir"]], axis=1))


@pytest.mark.parametrize("name", GOLDEN_SCENES)
def test_oracle_pixels_match_reference(oracle, name):
    _, rgbt = load_golden(name)
    o = oracle.OracleScene(os.path.join(GOLDEN, name + ".pvs"))
    img, st = o.render(W, H, threads=2)
    d = np.abs(img - rgbt).max(axis=2)
    assert (d > 1.0 / 255.0).mean() <= 0.001   # the contract
    # what the restatement actually achieves on these scenes: identical to ~1e-7 everywhere, except one pixel of the
    # refracting blob (1.4e-4) and, with area lights, the few samples the reference's light-source shadow cache answers
    # differently from its own uncached search (the cached object is tested without the
    # SMALL_TOLERANCE post-condition, trace.cpp:1989-2013, so the result depends on which object was cached last)
    if name == "area_lights":
        assert d.max() < 1e-3 and (d > 1e-5).sum() <= 8
    elif name == "normal_maps":
        # one pixel where the reference's shadow cache flips the second light's contribution; rendered alone
        # (+SC53 +EC53 +SR26 +ER26) the reference gives exactly the oracle's value
        assert d.max() < 0.03 and (d > 1e-5).sum() <= 1
    elif name == "crackle_cells":
        # the reference's per-thread crackle cache is keyed by the cell coordinates only (CrackleCellCoord::operator==,
        # cracklecache.h:74-77) and shared by every crackle pattern of the scene: a pattern with `repeat` leaves wrapped nuclei
        # behind that another pattern then picks up for the same cell (history dependent; one pixel here)
        assert d.max() < 0.05 and (d > 1e-5).sum() <= 2
    else:
        assert d.max() < 2e-4 and (d > 1e-5).sum() <= 1
    # (fisheye / omnimax: pixels outside the image circle trace nothing, tracepixel.cpp:408-411)
    assert st["rays"] >= (W * H if name not in ("cam_fisheye", "cam_omnimax") else 1000)


def test_oracle_rect_and_thread_invariance(oracle):
    o = oracle.OracleScene(os.path.join(GOLDEN, "csg_glass.pvs"))
    full, _ = o.render(W, H, threads=1)
    part, _ = o.render(W, H, rect=(10, 5, 41, 30), threads=3)
    assert np.array_equal(part, full[5:31, 10:42])


def test_oracle_solver_known_answers(oracle):
    import ctypes as C
    l = oracle.lib()

    def solve(c, sturm):
        cc = (C.c_double * len(c))(*c)
        r = (C.c_double * 4)()
        n = l.pvo_solve_polynomial(len(c) - 1, cc, r, sturm, 0.0)
        return sorted(r[i] for i in range(n))
    # (x-1)(x-2)(x-3)(x-4) = x^4 - 10x^3 + 35x^2 - 50x + 24
    for sturm in (0, 1):
        roots = solve([1.0, -10.0, 35.0, -50.0, 24.0], sturm)
        assert np.allclose(roots, [1, 2, 3, 4], atol=1e-7)
    assert np.allclose(solve([1.0, -6.0, 11.0, -6.0], 0), [1, 2, 3], atol=1e-9)
    assert solve([1.0, 0.0, 0.0, 0.0, 1.0], 0) == []          # x^4 + 1 has no real root
    assert np.allclose(solve([0.0, 0.0, 1.0, -3.0, 2.0], 0), [1, 2])   # leading zeros are stripped
