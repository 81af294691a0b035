"""cudaraytracing_b200 — B200-native path tracer behind the drop-in surface of guomc9/CudaRayTracing.

The product is cudaraytracing_b200/libcrt.so (hand-written CUDA for sm_100a behind the C-ABI of
include/crt.h) plus the headless CLI `crt`. This package is the thin host-side mirror of the
reference's classes (Scene, Render, Task/config, Camera) over that C-ABI via ctypes. It never
falls back to a CPU implementation: without the built library or without a GPU it raises.
"""
from .api import (CrtError, Config, Scene, Render, RenderGroup, inverse_view_matrix, load_config, write_png, device_count,
                  lib_path, load_library, ESTIMATOR_COMPAT, ESTIMATOR_MIS, RAY_CLOSEST, RAY_ANY, RAY_SORTED, BUILDER_LBVH, BUILDER_LBVH8,
                  BUILDER_PLOC, BUILDER_PLOC8)
from .build import build as build_native

__all__ = ["CrtError", "Config", "Scene", "Render", "RenderGroup", "inverse_view_matrix", "load_config", "write_png", "device_count",
           "lib_path", "load_library", "build_native", "ESTIMATOR_COMPAT", "ESTIMATOR_MIS", "RAY_CLOSEST", "RAY_ANY", "RAY_SORTED", "BUILDER_LBVH", "BUILDER_LBVH8",
           "BUILDER_PLOC", "BUILDER_PLOC8"]
