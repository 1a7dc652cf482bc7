"""The sharded render path behind the C ABI (pvgpu_scene_finalize_multi -> pvgpu_render / pvgpu_render_device):
one host thread per work context pulls chunks of rectangles from one atomic counter (the GetNextRectangle contract,
source/backend/scene/view.cpp:236-271) and the finished tiles are gathered into the caller's frame.

Checked here: the N-worker frame equals the 1-worker frame (same pixels up to the FP32 accumulation order, same ray
counters) - with two work contexts on ONE device (always runnable on the GPU box) and with two devices when the box
has them; the shadow-queue clamp + overflow retry (a tiny queue must not change the frame); chunk-count independence.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, has_gpu

pytestmark = pytest.mark.gpu

W, H = 320, 180


@pytest.fixture(scope="module")
def pv():
    if not has_gpu():
        pytest.skip("no CUDA device")
    import povray_b200
    return povray_b200


def n_devices():
    import torch
    return torch.cuda.device_count()


def render(pv, name, finalize, aa=None, env=None):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        s = pv.Scene.load(os.path.join(GOLDEN, name + ".pvs"))
        finalize(s)
        img, st = s.render_image(W, H, aa=aa)
        n = s.device_count
        s.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return img, st, n


def same_frame(a, b, sta, stb):
    assert np.abs(a - b).max() < 2e-5            # FP32 atomics: the order of the additions differs, nothing else
    for k in ("rays", "shadow_ray_tests", "reflected_rays", "refracted_rays", "transmitted_rays", "max_trace_level"):
        assert sta[k] == stb[k], k


@pytest.mark.parametrize("name", ["spheres64", "csg_glass", "mesh24", "layered_lights"])
def test_two_contexts_on_one_device_equal_one(pv, name):
    one, st1, _ = render(pv, name, lambda s: s.finalize(0))
    two, st2, n = render(pv, name, lambda s: s.finalize(0), env={"PVGPU_CTX_PER_DEVICE": "2"})
    assert n == 1
    same_frame(one, two, st1, st2)


@pytest.mark.parametrize("chunks", ["1", "3", "8"])
def test_chunk_count_does_not_change_the_frame(pv, chunks):
    one, st1, _ = render(pv, "csg_glass", lambda s: s.finalize(0))
    two, st2, _ = render(pv, "csg_glass", lambda s: s.finalize(0), env={"PVGPU_CTX_PER_DEVICE": "2", "PVGPU_CHUNKS_PER_WORKER": chunks})
    same_frame(one, two, st1, st2)


def test_two_contexts_with_antialiasing(pv):
    from povray_b200 import _abi as A
    aa = A.AA()
    aa.method, aa.depth, aa.threshold, aa.jitter_scale, aa.gamma = 2, 2, 0.3, 1.0, 2.5
    one, st1, _ = render(pv, "spheres64", lambda s: s.finalize(0), aa=aa)
    two, st2, _ = render(pv, "spheres64", lambda s: s.finalize(0), aa=aa, env={"PVGPU_CTX_PER_DEVICE": "2"})
    # adaptive AA re-traces the one-pixel frame around every tile, so tiles are independent of the sharding
    assert np.abs(one - two).max() < 2e-5
    assert st1["samples"] == st2["samples"]


def test_two_devices_equal_one(pv):
    if n_devices() < 2:
        pytest.skip("needs two CUDA devices")
    for name in ("spheres64", "csg_glass", "mesh24"):
        one, st1, _ = render(pv, name, lambda s: s.finalize(0))
        two, st2, n = render(pv, name, lambda s: s.finalize_multi(devices=[0, 1]))
        assert n == 2
        same_frame(one, two, st1, st2)


def test_all_devices_device_frame(pv):
    """pvgpu_render_device on a multi-device scene: tiles arrive in the first device's memory by peer copies."""
    import torch
    if n_devices() < 2:
        pytest.skip("needs two CUDA devices")
    from povray_b200.scene import tiles, assemble, _area
    s1 = pv.Scene.load(os.path.join(GOLDEN, "csg_glass.pvs")).finalize(0)
    ref, _ = s1.render_image(W, H)
    s = pv.Scene.load(os.path.join(GOLDEN, "csg_glass.pvs")).finalize_multi(n_devices=0)
    assert s.device_count == n_devices()
    rects = tiles(W, H)
    out = torch.zeros(_area(rects) * 4, dtype=torch.float32, device="cuda:0")
    s.render_device(W, H, rects, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    img = assemble(out.cpu().numpy().reshape(-1, 4), rects, W, H)
    assert np.abs(img - ref).max() < 2e-5


def test_tiny_shadow_queue_is_clamped_and_retried(pv):
    """the shadow kernels clamp their count to the queue capacity; an overflowing batch is retried in halves
    and the frame is unchanged.  PVGPU_TEST_QUEUE_CAP forces small queues."""
    one, st1, _ = render(pv, "layered_lights", lambda s: s.finalize(0))
    two, st2, _ = render(pv, "layered_lights", lambda s: s.finalize(0), env={"PVGPU_TEST_QUEUE_CAP": "4096"})
    assert np.abs(one - two).max() < 2e-5
