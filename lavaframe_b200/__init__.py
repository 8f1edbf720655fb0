"""lavaframe_b200 — B200-native implementation of LavaFrame's progressive path-tracing core.

The product is the CUDA library `liblfcuda.so` (C ABI: include/lfcuda.h) plus the C++ `CudaRenderer`
(`liblfhost.so`), a drop-in for the reference's TiledRenderer.  This Python package is only the ctypes
glue used by the tests, `bench.py` and `__graft_entry__.py`; there is no Python or CPU compute path.
"""
from .capi import (LfSceneView, LfParams, LfCamera, LfPostParams, LfCounters, LfStageStats, load_lfcuda, load_lfhost, LfCudaError,
                   STAGE_NAMES, LfBlasInfo, build_blas)
from .scenepack import ScenePack
from .pathtracer import PathTracer, PathTracerGroup
from .host import HostScene, CudaRenderer

__all__ = ["LfSceneView", "LfParams", "LfCamera", "LfPostParams", "LfCounters", "LfStageStats", "load_lfcuda", "load_lfhost",
           "LfCudaError", "STAGE_NAMES", "ScenePack", "PathTracer", "PathTracerGroup", "HostScene", "CudaRenderer", "LfBlasInfo", "build_blas"]
