"""pathtracerap_b200 - B200-native render core behind PathTracerAP's Scene / Renderer API.

The compute path is libptap.so (hand-written sm_100a CUDA behind the C ABI of include/ptap.h);
this package is the thin host-side mirror of the reference interface used by tests and bench.py.
"""
from ._native import (ACCEL_BVH, ACCEL_BVH_DEVICE, ACCEL_GRID_COMPAT, ACCEL_GRID_EMULATED, COAT, DIFFUSE, EMISSIVE, FLOAT_MAX, HIT, MATERIAL, MESH, METAL, MODEL,
                      PATH_IN, PATH_OUT, REFLECTIVE, REFRACTIVE, SPECULAR, TRIANGLE, VERTEX, PtapError)
from .renderer import Renderer
from .scene import Scene

__all__ = ["Scene", "Renderer", "PtapError", "ACCEL_GRID_COMPAT", "ACCEL_GRID_EMULATED", "ACCEL_BVH", "ACCEL_BVH_DEVICE"]
