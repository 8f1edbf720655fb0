"""ctypes access to the CPU oracle (oracle/liblforacle.so) — TEST INFRASTRUCTURE, used by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_lib = None


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)


def load_oracle():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(ORACLE_DIR, "liblforacle.so")
    src_newer = any(os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(path)
                    for f in ("lf_oracle.cpp", "lf_oracle.h", "lf_math_oracle.h", "lf_oracle_capi.cpp")) if os.path.exists(path) else True
    if src_newer:
        build_oracle()
    lib = C.CDLL(path)
    lib.lforacle_open_pack.restype = C.c_void_p
    lib.lforacle_open_pack.argtypes = [C.c_char_p]
    lib.lforacle_close.argtypes = [C.c_void_p]
    lib.lforacle_get_params.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lforacle_set_params.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lforacle_set_options.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.lforacle_render_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.lforacle_render_preview.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.lforacle_primary_hits.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lforacle_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.lforacle_rand_kat.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.lforacle_bsdf_kat.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.lforacle_fp_flags.restype = C.c_int
    lib.lforacle_fp_flags.argtypes = [C.c_int]
    lib.lforacle_post_process.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    lib.lforacle_builtin_kat.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.lforacle_set_threads.restype = C.c_int
    lib.lforacle_set_threads.argtypes = [C.c_int]
    lib.lforacle_get_counters.argtypes = [C.c_void_p, C.c_void_p]
    lib.lforacle_reset_counters.argtypes = [C.c_void_p]
    _lib = lib
    return lib


class Oracle:
    """CPU restatement of the reference shader over a .lfpack scene."""

    def __init__(self, pack_path, cull=False, count=False):
        from lavaframe_b200.capi import LfParams, LfCamera
        self.lib = load_oracle()
        self.h = self.lib.lforacle_open_pack(str(pack_path).encode())
        if not self.h:
            raise RuntimeError(f"oracle cannot open {pack_path}")
        self.params = LfParams()
        self.camera = LfCamera()
        self.lib.lforacle_get_params(self.h, C.byref(self.params), C.byref(self.camera))
        self.lib.lforacle_set_options(self.h, int(cull), int(count))

    def set_options(self, cull=False, count=False):
        self.lib.lforacle_set_options(self.h, int(cull), int(count))

    def update_params(self, **kw):
        for k, v in kw.items():
            setattr(self.params, k, v)
        self.lib.lforacle_set_params(self.h, C.byref(self.params), None)

    def render_frames(self, first_frame, nframes, frame_stride=1, tile_x=0, tile_y=0, accum=None):
        if accum is None:
            accum = np.zeros((self.params.height, self.params.width, 3), np.float32)
        self.lib.lforacle_render_frames(self.h, first_frame, nframes, frame_stride, tile_x, tile_y, accum.ctypes.data_as(C.c_void_p))
        return accum

    def render_preview(self, pv_w, pv_h, max_depth=2, use_dof=False):
        """preview_flareon.glsl: one sample per pixel of a pv_w x pv_h viewport, frame 1, rows bottom-up."""
        out = np.zeros((pv_h, pv_w, 3), np.float32)
        self.lib.lforacle_render_preview(self.h, pv_w, pv_h, max_depth, int(use_dof), out.ctypes.data_as(C.c_void_p))
        return out

    def primary_hits(self, frame=2):
        H, W = self.params.height, self.params.width
        t = np.empty((H, W), np.float32)
        tri = np.empty((H, W), np.int32)
        mat = np.empty((H, W), np.int32)
        em = np.empty((H, W), np.int32)
        self.lib.lforacle_primary_hits(self.h, frame, t.ctypes.data_as(C.c_void_p), tri.ctypes.data_as(C.c_void_p),
                                       mat.ctypes.data_as(C.c_void_p), em.ctypes.data_as(C.c_void_p))
        return t, tri, mat, em

    def counters(self):
        from lavaframe_b200.capi import LfCounters
        c = LfCounters()
        self.lib.lforacle_get_counters(self.h, C.byref(c))
        return c.as_dict()

    def reset_counters(self):
        self.lib.lforacle_reset_counters(self.h)

    def close(self):
        if self.h:
            self.lib.lforacle_close(self.h)
            self.h = None


def set_threads(n=0):
    """Set (n > 0) and return the OpenMP team size the oracle uses (torchrun exports OMP_NUM_THREADS=1)."""
    return int(load_oracle().lforacle_set_threads(int(n)))


def rand_kat(px, py, frame, n=4):
    lib = load_oracle()
    seed = np.zeros(n, np.uint32)
    vals = np.zeros(n, np.float32)
    lib.lforacle_rand_kat(px, py, frame, n, seed.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p))
    return seed, vals


def builtin_kat(op, args, tex=None):
    """The oracle's restatement of GLSL built-ins on (n, 4) float32 arguments (expression group `op`, lf_oracle.cpp BuiltinKat)."""
    lib = load_oracle()
    a = np.ascontiguousarray(args, np.float32)
    out = np.empty_like(a)
    if tex is not None:
        tex = np.ascontiguousarray(tex, np.uint8)
        lib.lforacle_builtin_kat(op, a.ctypes.data_as(C.c_void_p), a.shape[0], out.ctypes.data_as(C.c_void_p), tex.ctypes.data_as(C.c_void_p),
                                 tex.shape[2], tex.shape[1], tex.shape[0])
    else:
        lib.lforacle_builtin_kat(op, a.ctypes.data_as(C.c_void_p), a.shape[0], out.ctypes.data_as(C.c_void_p), None, 0, 0, 0)
    return out


def post_process(accum, inv, tonemap_index, post=None):
    """postprocess.glsl as the oracle restates it: accum (H, W, 3) float32 rows bottom-up -> output image.  `post` = lavaframe_b200.LfPostParams or None."""
    lib = load_oracle()
    a = np.ascontiguousarray(accum, np.float32)
    out = np.empty_like(a)
    lib.lforacle_post_process(a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0], float(inv), int(tonemap_index),
                              C.cast(C.byref(post), C.c_void_p) if post is not None else None, out.ctypes.data_as(C.c_void_p))
    return out


def bsdf_kat(op, items):
    """The oracle's DisneyEval / DisneySample / microfacet helpers on (n, 9, 4) float32 items (layout: lf_oracle.cpp BsdfKat)."""
    lib = load_oracle()
    a = np.ascontiguousarray(items, np.float32)
    assert a.ndim == 3 and a.shape[1:] == (9, 4)
    out = np.empty((a.shape[0], 4), np.float32)
    lib.lforacle_bsdf_kat(op, a.ctypes.data_as(C.c_void_p), a.shape[0], out.ctypes.data_as(C.c_void_p))
    return out
