import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_W, GOLDEN_H = 96, 54
GOLDEN_SCENES = ["spheres64", "mesh24", "csg_glass", "torus_noise", "layered_lights", "clipped_bounded", "blob_mix", "cones_csg", "patches", "normals", "patterns2", "sky_fog", "area_lights", "polys", "pigment_maps", "crackle_cells", "texture_maps", "normal_maps", "irid", "reflect_exp", "csg_children", "blob_textures", "image_maps", "looks_like", "looks_like_flat", "fractals", "text", "prisms", "normals_block", "superquadrics", "uv_mapping", "pigment_pattern", "warps",
                 "cam_normal", "cam_fisheye", "cam_omnimax", "cam_panoramic", "cam_ultrawide", "cam_cyl3", "cam_cyl2", "cam_spherical"]
ADAPTER = os.path.join(ROOT, "oracle", "_ref", "parity", "povray-gpu")
REF_BINARY = os.path.join(ROOT, "oracle", "_ref", "parity", "povray")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def pvlib():
    """The product library; building it is part of __graft_entry__.build()."""
    from povray_b200 import _abi
    if not os.path.exists(_abi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _abi.lib()


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib
