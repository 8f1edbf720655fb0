"""Thin object wrapper over the C ABI of liblfcuda.so (include/lfcuda.h).  No compute happens in Python."""
import ctypes as C

import numpy as np

from .capi import load_lfcuda, LfCudaError, LfParams, LfCamera, LfCounters, LfStageStats


class PathTracer:
    """One lfcuda context on one CUDA device: upload a scene, set uniforms, render frames, read results."""

    def __init__(self, device=0):
        self.lib = load_lfcuda()
        h = C.c_void_p()
        rc = self.lib.lfcuda_create(C.byref(h), int(device))
        if rc != 0:
            raise LfCudaError(f"lfcuda_create({device}) failed ({rc}): {self.lib.lfcuda_last_error(None).decode()}")
        self.h = h
        self.params = None
        self.camera = None
        self._keep = None

    def _ck(self, rc, what):
        if rc != 0:
            raise LfCudaError(f"{what} failed ({rc}): {self.lib.lfcuda_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.lfcuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene / uniforms
    def upload_pack(self, pack, **overrides):
        """Upload a ScenePack and set its params/camera; keyword overrides patch LfParams fields."""
        self._keep = pack
        view = pack.view()
        self._ck(self.lib.lfcuda_upload_scene(self.h, C.byref(view)), "lfcuda_upload_scene")
        p = pack.params()
        for k, v in overrides.items():
            setattr(p, k, v)
        self.set_params(p)
        self.set_camera(pack.camera())

    def upload_view(self, view, params, camera):
        self._ck(self.lib.lfcuda_upload_scene(self.h, C.byref(view)), "lfcuda_upload_scene")
        self.set_params(params)
        self.set_camera(camera)

    def set_params(self, p):
        self._ck(self.lib.lfcuda_set_params(self.h, C.byref(p)), "lfcuda_set_params")
        q = LfParams()
        C.memmove(C.byref(q), C.byref(p), C.sizeof(LfParams))
        self.params = q

    def update_params(self, **kw):
        p = self.params
        for k, v in kw.items():
            setattr(p, k, v)
        self.set_params(p)

    def set_camera(self, c):
        self._ck(self.lib.lfcuda_set_camera(self.h, C.byref(c)), "lfcuda_set_camera")
        self.camera = c

    def set_post(self, post=None):
        self._ck(self.lib.lfcuda_set_post(self.h, C.byref(post) if post is not None else None), "lfcuda_set_post")

    def set_stream(self, cuda_stream_handle):
        self._ck(self.lib.lfcuda_set_stream(self.h, C.c_void_p(cuda_stream_handle or 0)), "lfcuda_set_stream")

    def update_instances(self, transforms, materials, tlas_nodes, first_node):
        t = np.ascontiguousarray(transforms, np.float32).reshape(-1)
        m = np.ascontiguousarray(materials, np.float32).reshape(-1)
        n = np.ascontiguousarray(tlas_nodes, np.float32).reshape(-1)
        fp = C.POINTER(C.c_float)
        self._ck(self.lib.lfcuda_update_instances(self.h, t.ctypes.data_as(fp), t.size // 16, m.ctypes.data_as(fp), m.size // 28,
                                                  n.ctypes.data_as(fp), int(first_node), n.size // 9), "lfcuda_update_instances")

    def update_instances_device(self, transforms, materials=None, instance_material_ids=None):
        """Instance edit with the TLAS rebuilt on the device (lfcuda_update_instances_device)."""
        t = np.ascontiguousarray(transforms, np.float32).reshape(-1)
        fp = C.POINTER(C.c_float)
        m = np.ascontiguousarray(materials, np.float32).reshape(-1) if materials is not None else None
        ids = np.ascontiguousarray(instance_material_ids, np.int32) if instance_material_ids is not None else None
        self._ck(self.lib.lfcuda_update_instances_device(self.h, t.ctypes.data_as(fp), t.size // 16, m.ctypes.data_as(fp) if m is not None else None,
                                                         m.size // 28 if m is not None else 0,
                                                         ids.ctypes.data_as(C.POINTER(C.c_int32)) if ids is not None else None), "lfcuda_update_instances_device")

    def read_tlas_nodes(self, num_instances):
        out = np.empty((2 * num_instances - 1, 9), np.float32)
        n = C.c_int32()
        self._ck(self.lib.lfcuda_read_tlas_nodes(self.h, out.ctypes.data_as(C.POINTER(C.c_float)), out.shape[0], C.byref(n)), "lfcuda_read_tlas_nodes")
        return out[:n.value]

    # ---- hot path
    def clear(self):
        self._ck(self.lib.lfcuda_clear(self.h), "lfcuda_clear")

    def render_frames(self, first_frame, nframes, frame_stride=1, tile_x=0, tile_y=0):
        self._ck(self.lib.lfcuda_render_frames(self.h, first_frame, nframes, frame_stride, tile_x, tile_y), "lfcuda_render_frames")

    def render_preview(self, pv_width, pv_height, max_depth=2, use_dof=False, tonemap_index=0, is_in_preview=False):
        """The preview engine (preview_flareon.glsl): renders and returns the pv_height x pv_width x 3 image, rows bottom-up."""
        self._ck(self.lib.lfcuda_render_preview(self.h, pv_width, pv_height, max_depth, int(use_dof)), "lfcuda_render_preview")
        out = np.empty((pv_height, pv_width, 3), np.float32)
        self._ck(self.lib.lfcuda_read_preview(self.h, tonemap_index, 1 if is_in_preview else 0, out.ctypes.data_as(C.c_void_p)), "lfcuda_read_preview")
        return out

    def synchronize(self):
        self._ck(self.lib.lfcuda_synchronize(self.h), "lfcuda_synchronize")

    def _shape(self):
        return (self.params.height, self.params.width, 3)

    def read_accum(self, out=None):
        if out is None:
            out = np.empty(self._shape(), np.float32)
        self._ck(self.lib.lfcuda_read_accum(self.h, out.ctypes.data_as(C.c_void_p)), "lfcuda_read_accum")
        return out

    def read_output(self, inv_sample_counter, tonemap_index=0):
        out = np.empty(self._shape(), np.float32)
        self._ck(self.lib.lfcuda_read_output(self.h, float(inv_sample_counter), int(tonemap_index), out.ctypes.data_as(C.c_void_p)), "lfcuda_read_output")
        return out

    def read_output_u8(self, inv_sample_counter, tonemap_index=0):
        out = np.empty(self._shape(), np.uint8)
        self._ck(self.lib.lfcuda_read_output_u8(self.h, float(inv_sample_counter), int(tonemap_index), out.ctypes.data_as(C.c_void_p)), "lfcuda_read_output_u8")
        return out

    def accum_device_ptr(self):
        p = C.c_void_p()
        n = C.c_size_t()
        self._ck(self.lib.lfcuda_accum_device_ptr(self.h, C.byref(p), C.byref(n)), "lfcuda_accum_device_ptr")
        return p.value, n.value

    def primary_hits(self, frame=2):
        H, W = self.params.height, self.params.width
        t = np.empty((H, W), np.float32)
        tri = np.empty((H, W), np.int32)
        mat = np.empty((H, W), np.int32)
        em = np.empty((H, W), np.int32)
        self._ck(self.lib.lfcuda_read_primary_hits(self.h, int(frame), t.ctypes.data_as(C.c_void_p), tri.ctypes.data_as(C.c_void_p),
                                                   mat.ctypes.data_as(C.c_void_p), em.ctypes.data_as(C.c_void_p)), "lfcuda_read_primary_hits")
        return t, tri, mat, em

    # ---- multi-GPU
    def nccl_unique_id(self):
        buf = (C.c_char * 128)()
        rc = self.lib.lfcuda_nccl_unique_id(buf)
        if rc != 0:
            raise LfCudaError(f"lfcuda_nccl_unique_id failed: {self.lib.lfcuda_last_error(None).decode()}")
        return bytes(buf)

    def nccl_init(self, unique_id, rank, nranks):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        self._ck(self.lib.lfcuda_nccl_init(self.h, buf, rank, nranks), "lfcuda_nccl_init")

    def reduce(self):
        self._ck(self.lib.lfcuda_reduce(self.h), "lfcuda_reduce")

    # ---- instrumentation
    def reset_counters(self):
        self._ck(self.lib.lfcuda_reset_counters(self.h), "lfcuda_reset_counters")

    def counters(self):
        c = LfCounters()
        self._ck(self.lib.lfcuda_get_counters(self.h, C.byref(c)), "lfcuda_get_counters")
        return c.as_dict()

    def set_profiling(self, on):
        self._ck(self.lib.lfcuda_set_profiling(self.h, 1 if on else 0), "lfcuda_set_profiling")

    def stage_stats(self):
        s = LfStageStats()
        self._ck(self.lib.lfcuda_get_stage_stats(self.h, C.byref(s)), "lfcuda_get_stage_stats")
        return s.as_dict()

    def measure_read_bandwidth(self, nbytes, iters=20):
        g = C.c_double()
        self._ck(self.lib.lfcuda_measure_read_bandwidth(self.h, int(nbytes), int(iters), C.byref(g)), "lfcuda_measure_read_bandwidth")
        return float(g.value)

    def measure_node_fetch(self, table_bytes=8 << 20, iters=5):
        g = C.c_double()
        self._ck(self.lib.lfcuda_measure_node_fetch(self.h, int(table_bytes), int(iters), C.byref(g)), "lfcuda_measure_node_fetch")
        return float(g.value)

    def launch_count(self):
        n = C.c_uint64()
        self._ck(self.lib.lfcuda_get_launch_count(self.h, C.byref(n)), "lfcuda_get_launch_count")
        return int(n.value)


def algorithmic_bytes(c):
    """bytes the traversal must touch for the visit counts `c` (reference layout figures)."""
    return (60 * c["inner_visits"] + 12 * c["leaf_visits"] + 60 * c["tri_tests"] + 76 * c["tlas_visits"] + 60 * c["light_tests"])


def algorithmic_bytes_split(c):
    """(closest-hit bytes, shadow-ray bytes) of the traversal formula, for the extend and shadow kernels."""
    sh = (60 * c["inner_visits_shadow"] + 12 * c["leaf_visits_shadow"] + 60 * c["tri_tests_shadow"] + 76 * c["tlas_visits_shadow"]
          + 60 * c["light_tests_shadow"])
    return algorithmic_bytes(c) - sh, sh


def algorithmic_bytes_total(c):
    """Traversal bytes + shading (160 B per shaded hit, 16 B per texture sample), env NEE (80 B), env miss (48 B)
    and the accumulate read+write (24 B per pixel-sample)."""
    return (algorithmic_bytes(c) + 160 * c["shaded_hits"] + 16 * c["tex_samples"] + 80 * c["env_nee"] + 48 * c["env_miss"] + 24 * c["samples"])


class PathTracerGroup:
    """Several GPUs behind one renderer in ONE process (lfcuda_group_*): one context + host thread per device, frames dealt
    round-robin, the devices' accumulation buffers summed inside the post-process kernel of the first device over peer access."""

    def __init__(self, devices):
        self.lib = load_lfcuda()
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self.lib.lfcuda_group_create(C.byref(h), devs, len(devices))
        if rc != 0:
            raise LfCudaError(f"lfcuda_group_create({list(devices)}) failed ({rc}): {self.lib.lfcuda_group_last_error(None).decode()}")
        self.h = h
        self.devices = list(devices)
        self.params = None
        self._keep = None

    def _ck(self, rc, what):
        if rc != 0:
            raise LfCudaError(f"{what} failed ({rc}): {self.lib.lfcuda_group_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.lfcuda_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_pack(self, pack, **overrides):
        self._keep = pack
        view = pack.view()
        self._ck(self.lib.lfcuda_group_upload_scene(self.h, C.byref(view)), "lfcuda_group_upload_scene")
        p = pack.params()
        for k, v in overrides.items():
            setattr(p, k, v)
        self.set_params(p)
        cam = pack.camera()
        self._ck(self.lib.lfcuda_group_set_camera(self.h, C.byref(cam)), "lfcuda_group_set_camera")

    def set_params(self, p):
        self._ck(self.lib.lfcuda_group_set_params(self.h, C.byref(p)), "lfcuda_group_set_params")
        q = LfParams()
        C.memmove(C.byref(q), C.byref(p), C.sizeof(LfParams))
        self.params = q

    def set_post(self, post=None):
        self._ck(self.lib.lfcuda_group_set_post(self.h, C.byref(post) if post is not None else None), "lfcuda_group_set_post")

    def clear(self):
        self._ck(self.lib.lfcuda_group_clear(self.h), "lfcuda_group_clear")

    def synchronize(self):
        self._ck(self.lib.lfcuda_group_synchronize(self.h), "lfcuda_group_synchronize")

    def render_frames(self, first_frame, nframes, frame_stride=1, tile_x=0, tile_y=0):
        self._ck(self.lib.lfcuda_group_render_frames(self.h, first_frame, nframes, frame_stride, tile_x, tile_y), "lfcuda_group_render_frames")

    def _shape(self):
        return (self.params.height, self.params.width, 3)

    def read_accum(self):
        out = np.empty(self._shape(), np.float32)
        self._ck(self.lib.lfcuda_group_read_accum(self.h, out.ctypes.data_as(C.c_void_p)), "lfcuda_group_read_accum")
        return out

    def read_output(self, inv_sample_counter, tonemap_index=0):
        out = np.empty(self._shape(), np.float32)
        self._ck(self.lib.lfcuda_group_read_output(self.h, float(inv_sample_counter), int(tonemap_index), out.ctypes.data_as(C.c_void_p)), "lfcuda_group_read_output")
        return out

    def read_output_u8(self, inv_sample_counter, tonemap_index=0):
        out = np.empty(self._shape(), np.uint8)
        self._ck(self.lib.lfcuda_group_read_output_u8(self.h, float(inv_sample_counter), int(tonemap_index), out.ctypes.data_as(C.c_void_p)), "lfcuda_group_read_output_u8")
        return out
