"""povray_b200: B200-native trace path for POV-Ray behind the pvgpu C ABI (include/pvgpu.h).

`povray_b200.Scene` wraps a flattened scene living in HBM; `povray_b200.synth` builds the synthetic
benchmark scenes of BASELINE.json.  All tracing happens in povray_b200/libpvgpu.so (hand-written
sm_100a CUDA); nothing here computes pixels on the CPU.
"""
from . import _abi as abi
from ._abi import PvgpuError
from .scene import Scene, HostBuffer, tiles, assemble

__all__ = ["abi", "PvgpuError", "Scene", "HostBuffer", "tiles", "assemble"]
