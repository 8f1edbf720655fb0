"""ctypes handles over liblfhost.so: the reference's own scene loader / BVH builder (compiled unchanged) and
the C++ CudaRenderer that drops in for TiledRenderer behind the reference's Renderer interface."""
import ctypes as C

import numpy as np

from .capi import load_lfhost, LfSceneView, LfParams, LfCamera, LfCudaError


class HostScene:
    """A LavaFrame::Scene loaded by LoadSceneFromFile (LavaFrame/Loader.cpp:45-386)."""

    def __init__(self, path, keep_tonemap=False, verbose=False):
        self.lib = load_lfhost()
        self.h = self.lib.lfhost_load_scene(str(path).encode(), 1 if keep_tonemap else 0, 1 if verbose else 0)
        if not self.h:
            raise LfCudaError(f"cannot load scene {path}")

    def views(self):
        v, p, c = LfSceneView(), LfParams(), LfCamera()
        self.lib.lfhost_scene_view(self.h, C.byref(v), C.byref(p), C.byref(c))
        return v, p, c

    def write_pack(self, path):
        if self.lib.lfhost_write_pack(self.h, str(path).encode()) != 0:
            raise LfCudaError(f"cannot write {path}")

    def set_render_options(self, max_depth=0, tile_w=0, tile_h=0, use_constant_bg=-1, bg=None):
        bgp = None
        if bg is not None:
            arr = (C.c_float * 3)(*bg)
            bgp = C.cast(arr, C.POINTER(C.c_float))
        self.lib.lfhost_set_render_options(self.h, max_depth, tile_w, tile_h, use_constant_bg, bgp)

    def move_instance(self, index, mat4):
        m = np.ascontiguousarray(mat4, np.float32).reshape(16)
        if self.lib.lfhost_move_instance(self.h, index, m.ctypes.data_as(C.POINTER(C.c_float))) != 0:
            raise LfCudaError("bad instance index")

    def set_camera_moving(self, moving):
        self.lib.lfhost_set_camera_moving(self.h, 1 if moving else 0)

    def set_preview(self, scale=1.0, use_dof=False):
        """GlobalState.previewScale / useDofInPreview (Main.cpp:525-529); read by the renderer's Init like TiledRenderer.cpp:61,90."""
        self.lib.lfhost_set_preview(float(scale), 1 if use_dof else 0)

    def close(self):
        if self.h:
            self.lib.lfhost_free_scene(self.h)
            self.h = None


class CudaRenderer:
    """The C++ LavaFrame::CudaRenderer (lavaframe_b200/host/CudaRenderer.h), same methods as the reference's
    Renderer interface (LavaFrame/Renderer.h:71-117)."""

    def __init__(self, scene, device=0, devices=None):
        self.lib = scene.lib
        self.scene = scene
        if devices is None:
            self.h = self.lib.lfhost_renderer_create(scene.h, device)
        else:   # several GPUs behind the one renderer (lfcuda_group_*)
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            self.h = self.lib.lfhost_renderer_create_multi(scene.h, arr, len(devices))
        if not self.lib.lfhost_renderer_ok(self.h):
            msg = self.lib.lfhost_renderer_error(self.h).decode()
            self.lib.lfhost_renderer_destroy(self.h)
            self.h = None
            raise LfCudaError(f"CudaRenderer::Init failed: {msg}")

    def SetDeviceTlasRebuild(self, on=True):
        self.lib.lfhost_renderer_set_device_tlas(self.h, 1 if on else 0)

    def Update(self, dt=0.0):
        self.lib.lfhost_renderer_update(self.h, dt)

    def Render(self):
        self.lib.lfhost_renderer_render(self.h)

    def GetSampleCount(self):
        return self.lib.lfhost_renderer_sample_count(self.h)

    def GetProgress(self):
        return self.lib.lfhost_renderer_progress(self.h)

    def Flush(self):
        self.lib.lfhost_renderer_flush(self.h)

    def Run(self, spp):
        """Main.cpp's loop: Update -> Render until maxSamples + 1 == GetSampleCount()."""
        return self.lib.lfhost_renderer_run(self.h, spp)

    def _dims(self):
        w, h = C.c_int(), C.c_int()
        self.lib.lfhost_renderer_output_hdr(self.h, None, C.byref(w), C.byref(h))
        return w.value, h.value

    def GetOutputBufferHDR(self):
        _, p, _ = self.scene.views()
        out = np.empty((p.height, p.width, 3), np.float32)
        w, h = C.c_int(), C.c_int()
        self.lib.lfhost_renderer_output_hdr(self.h, out.ctypes.data_as(C.c_void_p), C.byref(w), C.byref(h))
        return out

    def GetOutputBuffer(self):
        _, p, _ = self.scene.views()
        out = np.empty((p.height, p.width, 3), np.uint8)
        w, h = C.c_int(), C.c_int()
        self.lib.lfhost_renderer_output_u8(self.h, out.ctypes.data_as(C.c_void_p), C.byref(w), C.byref(h))
        return out

    def GetPreviewBufferHDR(self):
        """What Present()/SetViewport() show while the camera moves: the preview engine's image, or None."""
        w, h = C.c_int(), C.c_int()
        if self.lib.lfhost_renderer_preview_hdr(self.h, None, C.byref(w), C.byref(h)) != 0:
            return None
        out = np.empty((h.value, w.value, 3), np.float32)
        self.lib.lfhost_renderer_preview_hdr(self.h, out.ctypes.data_as(C.c_void_p), C.byref(w), C.byref(h))
        return out

    def context_handle(self):
        return self.lib.lfhost_renderer_ctx(self.h)

    def close(self):
        if self.h:
            self.lib.lfhost_renderer_destroy(self.h)
            self.h = None
