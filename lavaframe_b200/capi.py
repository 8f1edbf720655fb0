"""ctypes mirror of include/lfcuda.h (the C ABI of liblfcuda.so) and loader of the in-tree shared libraries.

The libraries are built in-tree by `__graft_entry__.build()` (CMake).  Loading fails loudly when they are
missing: there is deliberately no fallback implementation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)


class LfSceneView(C.Structure):
    _fields_ = [
        ("bvh_nodes", c_float_p), ("num_nodes", C.c_int32), ("top_bvh_index", C.c_int32),
        ("vert_indices", c_int32_p), ("num_tri_refs", C.c_int32),
        ("vertices_uvx", c_float_p), ("normals_uvy", c_float_p), ("num_vertices", C.c_int32),
        ("transforms", c_float_p), ("num_instances", C.c_int32),
        ("materials", c_float_p), ("num_materials", C.c_int32),
        ("lights", c_float_p), ("num_lights", C.c_int32),
        ("texture_maps", c_uint8_p), ("tex_width", C.c_int32), ("tex_height", C.c_int32), ("num_textures", C.c_int32),
        ("hdr_cols", c_float_p), ("hdr_marginal", c_float_p), ("hdr_conditional", c_float_p),
        ("hdr_width", C.c_int32), ("hdr_height", C.c_int32),
    ]


class LfParams(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("tile_width", C.c_int32), ("tile_height", C.c_int32),
        ("max_depth", C.c_int32), ("enable_rr", C.c_int32), ("rr_depth", C.c_int32), ("use_envmap", C.c_int32),
        ("use_constant_bg", C.c_int32), ("bg_color", C.c_float * 3), ("hdr_multiplier", C.c_float),
        ("kernel_mode", C.c_int32), ("no_cull", C.c_int32), ("count_work", C.c_int32), ("frames_in_flight", C.c_int32),
    ]


class LfCamera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("right", C.c_float * 3), ("up", C.c_float * 3), ("forward", C.c_float * 3),
                ("fov", C.c_float), ("focal_dist", C.c_float), ("aperture", C.c_float)]


class LfPostParams(C.Structure):
    _fields_ = [("use_ca", C.c_int32), ("use_ca_distortion", C.c_int32), ("ca_distance", C.c_float), ("ca_p1", C.c_float),
                ("ca_p2", C.c_float), ("ca_p3", C.c_float), ("use_vignette", C.c_int32), ("vignette_intensity", C.c_float),
                ("vignette_power", C.c_float)]


class LfCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("samples", "rays_closest", "rays_shadow", "inner_visits", "leaf_visits", "tri_tests",
                                          "tlas_visits", "light_tests", "shaded_hits", "env_nee", "env_miss", "tex_samples",
                                          "inner_visits_shadow", "leaf_visits_shadow", "tri_tests_shadow", "tlas_visits_shadow",
                                          "light_tests_shadow")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


STAGE_NAMES = ("generate", "extend", "shade", "shadow", "accumulate", "megakernel", "sample")


class LfStageStats(C.Structure):
    _fields_ = [("launches", C.c_uint64 * 7), ("ms", C.c_double * 7)]

    def as_dict(self):
        return {n: {"launches": int(self.launches[i]), "ms": float(self.ms[i])} for i, n in enumerate(STAGE_NAMES)}


class LfBlasInfo(C.Structure):
    _fields_ = [("num_nodes", C.c_int32), ("num_indices", C.c_int32), ("height", C.c_int32), ("negative_zero", C.c_int32),
                ("levels", C.c_int32), ("launches", C.c_int32), ("build_ms", C.c_float), ("total_ms", C.c_float)]


class LfCudaError(RuntimeError):
    pass


# every symbol include/lfcuda.h declares: (restype, argtypes)
LFCUDA_SYMBOLS = {
    "lfcuda_abi_version": (C.c_int, []),
    "lfcuda_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "lfcuda_destroy": (None, [C.c_void_p]),
    "lfcuda_last_error": (C.c_char_p, [C.c_void_p]),
    "lfcuda_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lfcuda_synchronize": (C.c_int, [C.c_void_p]),
    "lfcuda_upload_scene": (C.c_int, [C.c_void_p, C.POINTER(LfSceneView)]),
    "lfcuda_update_instances": (C.c_int, [C.c_void_p, c_float_p, C.c_int32, c_float_p, C.c_int32, c_float_p, C.c_int32, C.c_int32]),
    "lfcuda_update_instances_device": (C.c_int, [C.c_void_p, c_float_p, C.c_int32, c_float_p, C.c_int32, C.POINTER(C.c_int32)]),
    "lfcuda_read_tlas_nodes": (C.c_int, [C.c_void_p, c_float_p, C.c_int32, C.POINTER(C.c_int32)]),
    "lfcuda_set_params": (C.c_int, [C.c_void_p, C.POINTER(LfParams)]),
    "lfcuda_set_camera": (C.c_int, [C.c_void_p, C.POINTER(LfCamera)]),
    "lfcuda_set_post": (C.c_int, [C.c_void_p, C.POINTER(LfPostParams)]),
    "lfcuda_clear": (C.c_int, [C.c_void_p]),
    "lfcuda_render_frames": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "lfcuda_render_preview": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "lfcuda_group_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32]),
    "lfcuda_group_destroy": (None, [C.c_void_p]),
    "lfcuda_group_last_error": (C.c_char_p, [C.c_void_p]),
    "lfcuda_group_size": (C.c_int, [C.c_void_p]),
    "lfcuda_group_ctx": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "lfcuda_group_upload_scene": (C.c_int, [C.c_void_p, C.POINTER(LfSceneView)]),
    "lfcuda_group_update_instances": (C.c_int, [C.c_void_p, c_float_p, C.c_int32, c_float_p, C.c_int32, c_float_p, C.c_int32, C.c_int32]),
    "lfcuda_group_update_instances_device": (C.c_int, [C.c_void_p, c_float_p, C.c_int32, c_float_p, C.c_int32, C.POINTER(C.c_int32)]),
    "lfcuda_group_set_params": (C.c_int, [C.c_void_p, C.POINTER(LfParams)]),
    "lfcuda_group_set_camera": (C.c_int, [C.c_void_p, C.POINTER(LfCamera)]),
    "lfcuda_group_set_post": (C.c_int, [C.c_void_p, C.POINTER(LfPostParams)]),
    "lfcuda_group_clear": (C.c_int, [C.c_void_p]),
    "lfcuda_group_synchronize": (C.c_int, [C.c_void_p]),
    "lfcuda_group_render_frames": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "lfcuda_group_read_output": (C.c_int, [C.c_void_p, C.c_float, C.c_int32, C.c_void_p]),
    "lfcuda_group_read_output_u8": (C.c_int, [C.c_void_p, C.c_float, C.c_int32, C.c_void_p]),
    "lfcuda_group_read_accum": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lfcuda_measure_node_fetch": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int32, C.POINTER(C.c_double)]),
    "lfcuda_read_preview": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "lfcuda_read_accum": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lfcuda_read_output": (C.c_int, [C.c_void_p, C.c_float, C.c_int32, C.c_void_p]),
    "lfcuda_read_output_u8": (C.c_int, [C.c_void_p, C.c_float, C.c_int32, C.c_void_p]),
    "lfcuda_accum_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "lfcuda_read_primary_hits": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lfcuda_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "lfcuda_nccl_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "lfcuda_reduce": (C.c_int, [C.c_void_p]),
    "lfcuda_reset_counters": (C.c_int, [C.c_void_p]),
    "lfcuda_get_counters": (C.c_int, [C.c_void_p, C.POINTER(LfCounters)]),
    "lfcuda_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "lfcuda_get_stage_stats": (C.c_int, [C.c_void_p, C.POINTER(LfStageStats)]),
    "lfcuda_get_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "lfcuda_measure_read_bandwidth": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int32, C.POINTER(C.c_double)]),
    "lfcuda_build_blas": (C.c_int, [C.c_int32, c_float_p, C.c_int32, C.c_float, C.c_int32, c_float_p, C.POINTER(C.c_int32), C.POINTER(LfBlasInfo)]),
}

_lfcuda = None
_lfhost = None


def lib_path(name):
    return os.path.join(_HERE, name)


def load_lfcuda():
    """Load liblfcuda.so and bind every declared entry point.  Raises if the CUDA library was not built."""
    global _lfcuda
    if _lfcuda is not None:
        return _lfcuda
    path = os.environ.get("LF_LFCUDA_SO") or lib_path("liblfcuda.so")   # the override is for A/B experiments only
    if not os.path.exists(path):
        raise LfCudaError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the path-tracing core)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in LFCUDA_SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.lfcuda_abi_version() != 2:
        raise LfCudaError("liblfcuda.so ABI version mismatch")
    _lfcuda = lib
    return lib


def load_lfhost():
    """Load liblfhost.so (CudaRenderer + the reference's unchanged scene loader / BVH builder)."""
    global _lfhost
    if _lfhost is not None:
        return _lfhost
    load_lfcuda()
    path = lib_path("liblfhost.so")
    if not os.path.exists(path):
        raise LfCudaError(f"{path} is missing: it is built by __graft_entry__.build() where /root/reference exists")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.lfhost_load_scene.restype = C.c_void_p
    lib.lfhost_load_scene.argtypes = [C.c_char_p, C.c_int, C.c_int]
    lib.lfhost_free_scene.argtypes = [C.c_void_p]
    lib.lfhost_scene_view.argtypes = [C.c_void_p, C.POINTER(LfSceneView), C.POINTER(LfParams), C.POINTER(LfCamera)]
    lib.lfhost_write_pack.argtypes = [C.c_void_p, C.c_char_p]
    lib.lfhost_set_render_options.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p]
    lib.lfhost_move_instance.argtypes = [C.c_void_p, C.c_int, c_float_p]
    lib.lfhost_set_camera_moving.argtypes = [C.c_void_p, C.c_int]
    lib.lfhost_renderer_create.restype = C.c_void_p
    lib.lfhost_renderer_create.argtypes = [C.c_void_p, C.c_int]
    lib.lfhost_renderer_create_multi.restype = C.c_void_p
    lib.lfhost_renderer_create_multi.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
    lib.lfhost_renderer_destroy.argtypes = [C.c_void_p]
    lib.lfhost_renderer_ok.argtypes = [C.c_void_p]
    lib.lfhost_renderer_error.restype = C.c_char_p
    lib.lfhost_renderer_error.argtypes = [C.c_void_p]
    lib.lfhost_renderer_update.argtypes = [C.c_void_p, C.c_float]
    lib.lfhost_renderer_render.argtypes = [C.c_void_p]
    lib.lfhost_renderer_sample_count.argtypes = [C.c_void_p]
    lib.lfhost_renderer_progress.restype = C.c_float
    lib.lfhost_renderer_progress.argtypes = [C.c_void_p]
    lib.lfhost_renderer_flush.argtypes = [C.c_void_p]
    lib.lfhost_renderer_ctx.restype = C.c_void_p
    lib.lfhost_renderer_ctx.argtypes = [C.c_void_p]
    lib.lfhost_renderer_output_hdr.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.lfhost_renderer_output_u8.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.lfhost_renderer_run.argtypes = [C.c_void_p, C.c_int]
    lib.lfhost_renderer_preview_hdr.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.lfhost_set_preview.argtypes = [C.c_float, C.c_int]
    lib.lfhost_set_device_blas.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.lfhost_blas_stats.argtypes = [C.POINTER(C.c_double), C.c_int]
    _lfhost = lib
    return lib


def build_blas(prim_bounds, device=0, traversal_cost=2.0, num_bins=64):
    """lfcuda_build_blas on an [n, 6] float32 array of triangle boxes: (boxes [k, 6] float32, lr [k, 3] int32, indices [n] int32, info dict).
    The arguments default to what Mesh.h:18 constructs: SplitBvh(2.0f, 64, 0, 0.001f, 0)."""
    import numpy as np
    lib = load_lfcuda()
    b = np.ascontiguousarray(prim_bounds, dtype=np.float32).reshape(-1, 6)
    n = b.shape[0]
    nodes = np.zeros(9 * max(2 * n - 1, 1), np.float32)
    idx = np.zeros(max(n, 1), np.int32)
    info = LfBlasInfo()
    rc = lib.lfcuda_build_blas(device, b.ctypes.data_as(c_float_p), n, traversal_cost, num_bins, nodes.ctypes.data_as(c_float_p),
                               idx.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(info))
    if rc != 0:
        raise LfCudaError(f"lfcuda_build_blas: {rc}: {lib.lfcuda_last_error(None).decode()}")
    a = nodes[:9 * info.num_nodes].reshape(info.num_nodes, 9)
    return a[:, :6].copy(), a[:, 6:].copy().view(np.int32), idx[:n], {f: getattr(info, f) for f, _ in LfBlasInfo._fields_}

