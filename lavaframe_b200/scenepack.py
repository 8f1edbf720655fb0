"""Reader of the .lfpack scene dump (lavaframe_b200/host/scenepack.h): the flattened arrays the reference
uploads to the GPU (LavaFrame/Renderer.cpp:87-185) + RenderOptions/Camera uniforms."""
import ctypes as C

import numpy as np

from .capi import LfSceneView, LfParams, LfCamera, c_float_p, c_int32_p, c_uint8_p

I = dict(num_nodes=0, top_index=1, num_tri_refs=2, num_vertices=3, num_instances=4, num_materials=5, num_lights=6,
         tex_w=7, tex_h=8, num_tex=9, hdr_w=10, hdr_h=11, width=12, height=13, tile_w=14, tile_h=15, max_depth=16,
         enable_rr=17, rr_depth=18, use_envmap=19, use_const_bg=20, tonemap=21)


class ScenePack:
    def __init__(self, path):
        self.path = str(path)
        raw = np.fromfile(self.path, dtype=np.uint8)
        if raw[:8].tobytes() != b"LFPACK01":
            raise ValueError(f"{path}: not an LFPACK01 file")
        self.ihdr = raw[8:8 + 128].view(np.int32).copy()
        self.fhdr = raw[136:136 + 128].view(np.float32).copy()
        h = self.ihdr
        off = 264

        def take(dtype, count):
            nonlocal off
            n = int(count) * np.dtype(dtype).itemsize
            a = raw[off:off + n].view(dtype).copy()
            if a.size != count:
                raise ValueError(f"{path}: truncated pack")
            off += n
            return a

        hdrn = int(h[I["hdr_w"]]) * int(h[I["hdr_h"]])
        self.nodes = take(np.float32, 9 * h[I["num_nodes"]])
        self.vert_indices = take(np.int32, 3 * h[I["num_tri_refs"]])
        self.vertices = take(np.float32, 4 * h[I["num_vertices"]])
        self.normals = take(np.float32, 4 * h[I["num_vertices"]])
        self.transforms = take(np.float32, 16 * h[I["num_instances"]])
        self.materials = take(np.float32, 28 * h[I["num_materials"]])
        self.lights = take(np.float32, 15 * h[I["num_lights"]])
        self.textures = take(np.uint8, 4 * int(h[I["tex_w"]]) * int(h[I["tex_h"]]) * int(h[I["num_tex"]]))
        self.hdr_cols = take(np.float32, 3 * hdrn)
        self.hdr_marginal = take(np.float32, 2 * h[I["hdr_h"]] if hdrn else 0)
        self.hdr_conditional = take(np.float32, 2 * hdrn)

    def __getattr__(self, name):
        if name in I:
            return int(self.ihdr[I[name]])
        raise AttributeError(name)

    @staticmethod
    def _ptr(a, ptype):
        return a.ctypes.data_as(ptype) if a.size else C.cast(None, ptype)

    def view(self):
        """LfSceneView pointing into this object's arrays (keep the ScenePack alive while it is used)."""
        v = LfSceneView()
        v.bvh_nodes = self._ptr(self.nodes, c_float_p); v.num_nodes = self.num_nodes; v.top_bvh_index = self.top_index
        v.vert_indices = self._ptr(self.vert_indices, c_int32_p); v.num_tri_refs = self.num_tri_refs
        v.vertices_uvx = self._ptr(self.vertices, c_float_p); v.normals_uvy = self._ptr(self.normals, c_float_p)
        v.num_vertices = self.num_vertices
        v.transforms = self._ptr(self.transforms, c_float_p); v.num_instances = self.num_instances
        v.materials = self._ptr(self.materials, c_float_p); v.num_materials = self.num_materials
        v.lights = self._ptr(self.lights, c_float_p); v.num_lights = self.num_lights
        v.texture_maps = self._ptr(self.textures, c_uint8_p)
        v.tex_width = self.tex_w; v.tex_height = self.tex_h; v.num_textures = self.num_tex
        v.hdr_cols = self._ptr(self.hdr_cols, c_float_p); v.hdr_marginal = self._ptr(self.hdr_marginal, c_float_p)
        v.hdr_conditional = self._ptr(self.hdr_conditional, c_float_p)
        v.hdr_width = self.hdr_w; v.hdr_height = self.hdr_h
        return v

    def params(self):
        p = LfParams()
        p.width, p.height, p.tile_width, p.tile_height = self.width, self.height, self.tile_w, self.tile_h
        p.max_depth, p.enable_rr, p.rr_depth = self.max_depth, self.enable_rr, self.rr_depth
        p.use_envmap, p.use_constant_bg = self.use_envmap, self.use_const_bg
        for k in range(3):
            p.bg_color[k] = float(self.fhdr[k])
        p.hdr_multiplier = float(self.fhdr[3])
        return p

    def camera(self):
        c = LfCamera()
        f = self.fhdr
        for k in range(3):
            c.position[k] = float(f[4 + k]); c.right[k] = float(f[7 + k]); c.up[k] = float(f[10 + k]); c.forward[k] = float(f[13 + k])
        c.fov, c.focal_dist, c.aperture = float(f[16]), float(f[17]), float(f[18])
        return c
