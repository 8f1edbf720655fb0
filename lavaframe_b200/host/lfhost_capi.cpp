// lfhost_capi.cpp — extern "C" handles over the C++ host side (the reference's Scene/Loader, unchanged, and
// CudaRenderer) so that the Python tests and bench.py can drive the real drop-in class through ctypes.
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "Scene.h"
#include "Loader.h"
#include "Camera.h"
#include "GlobalState.h"
#include "CudaRenderer.h"
#include "scene_view.h"
#include "scenepack.h"
#include "DeviceBvh.h"

using namespace LavaFrame;
extern LavaFrameState GlobalState;

static int quiet_log(const char*, ...) { return 0; }

extern "C" {

// LoadSceneFromFile + `scene->renderOptions = renderOptions` exactly as Main.cpp:961-969 does, with the linear
// tonemap unless keep_tonemap is set (parity work compares radiance).
void* lfhost_load_scene(const char* path, int keep_tonemap, int verbose) {
    LavaFrame::Log = verbose ? printf : quiet_log;
    RenderOptions ro;
    ro.useVignette = false; ro.vignetteIntensity = 0.f; ro.vignettePower = 1.f;   // uninitialised by the ctor (Renderer.h:41-43)
    if (!keep_tonemap) ro.tonemapIndex = 0;
    Scene* scene = new Scene();
    GlobalState.scene = scene;
    bool ok = false;
    try { ok = LoadSceneFromFile(path, scene, ro); }
    catch (const std::exception& e) {   // a device BLAS build that was asked for and failed (DeviceBvh.cpp) must not unwind through the C boundary
        fprintf(stderr, "lfhost_load_scene: %s\n", e.what());
    }
    if (!ok) { delete scene; return nullptr; }
    if (!keep_tonemap) ro.tonemapIndex = 0;
    scene->renderOptions = ro;
    scene->camera->isMoving = false;   // never initialised by Camera's ctor (Camera.cpp:101-117)
    return scene;
}
void lfhost_free_scene(void* s) { delete static_cast<Scene*>(s); }

// Mesh BVHs built on the GPU (DeviceBvh.h) for the scenes loaded from now on: enable 0 / 1 (-1: keep), meshes below min_prims triangles stay on
// the host (-1: keep), CUDA device (-1: keep).  Off by default: north_star keeps the reference's host build as the source of the node arrays.
void lfhost_set_device_blas(int enable, int min_prims, int device) { lfhost::SetDeviceBlas(enable, min_prims, device); }
// out[8]: device builds, host builds, device-built meshes with a -0.0 bound, device triangles, host triangles, device ms (kernels), device ms (whole calls), host ms
void lfhost_blas_stats(double* out, int reset) {
    lfhost::BlasStats s = lfhost::GetBlasStats(reset != 0);
    out[0] = s.device_builds; out[1] = s.host_builds; out[2] = s.negative_zero_meshes; out[3] = (double)s.device_prims; out[4] = (double)s.host_prims;
    out[5] = s.device_ms; out[6] = s.device_total_ms; out[7] = s.host_ms;
}

void lfhost_scene_view(void* s, LfSceneView* v, LfParams* p, LfCamera* c) {
    Scene* scene = static_cast<Scene*>(s);
    if (v) lfhost::MakeSceneView(scene, v);
    if (p) lfhost::MakeParams(scene, p);
    if (c) lfhost::MakeCamera(scene, c);
}
int lfhost_write_pack(void* s, const char* path) {
    Scene* scene = static_cast<Scene*>(s);
    LfSceneView v; LfParams p; LfCamera c;
    lfhost::MakeSceneView(scene, &v); lfhost::MakeParams(scene, &p); lfhost::MakeCamera(scene, &c);
    lfpack::ScenePack pack;
    lfpack::from_views(pack, v, p, c, scene->renderOptions.tonemapIndex);
    return lfpack::write(path, pack) ? 0 : -1;
}
void lfhost_set_render_options(void* s, int max_depth, int tile_w, int tile_h, int use_constant_bg, const float* bg) {
    Scene* scene = static_cast<Scene*>(s);
    if (max_depth > 0) scene->renderOptions.maxDepth = max_depth;
    if (tile_w > 0) scene->renderOptions.tileWidth = tile_w;
    if (tile_h > 0) scene->renderOptions.tileHeight = tile_h;
    if (use_constant_bg >= 0) scene->renderOptions.useConstantBg = use_constant_bg != 0;
    if (bg) scene->renderOptions.bgColor = Vec3(bg[0], bg[1], bg[2]);
}
// Instance edit path (Scene::RebuildInstances, Scene.cpp:165-178): new transform for one instance, TLAS rebuilt by the
// reference's own code, scene->instancesModified raised for the renderer's next Update.
int lfhost_move_instance(void* s, int index, const float* m16) {
    Scene* scene = static_cast<Scene*>(s);
    if (index < 0 || index >= (int)scene->meshInstances.size()) return -1;
    std::memcpy(scene->meshInstances[index].transform.data, m16, 64);
    scene->RebuildInstances();
    return 0;
}
void lfhost_set_camera_moving(void* s, int moving) { static_cast<Scene*>(s)->camera->isMoving = moving != 0; }

// ---- the drop-in renderer ----
void* lfhost_renderer_create(void* s, int device) {
    Scene* scene = static_cast<Scene*>(s);
    CudaRenderer* r = new CudaRenderer(scene, GlobalState.shadersDir, device);   // replaces `new TiledRenderer(...)`, Main.cpp:91
    GlobalState.renderer = r;
    r->Init();
    return r;
}
// ... on several GPUs of this process (CudaRenderer(scene, dir, devices)); still the single construction of Main.cpp:91
void* lfhost_renderer_create_multi(void* s, const int* devices, int ndev) {
    Scene* scene = static_cast<Scene*>(s);
    CudaRenderer* r = new CudaRenderer(scene, GlobalState.shadersDir, std::vector<int>(devices, devices + ndev));
    GlobalState.renderer = r;
    r->Init();
    return r;
}
void lfhost_renderer_set_device_tlas(void* r, int on) { static_cast<CudaRenderer*>(r)->SetDeviceTlasRebuild(on != 0); }
void lfhost_renderer_destroy(void* r) { delete static_cast<CudaRenderer*>(r); }
int  lfhost_renderer_ok(void* r) { return static_cast<CudaRenderer*>(r)->Ok() ? 1 : 0; }
const char* lfhost_renderer_error(void* r) { return static_cast<CudaRenderer*>(r)->LastError(); }
void lfhost_renderer_update(void* r, float dt) { static_cast<Renderer*>(static_cast<CudaRenderer*>(r))->Update(dt); }
void lfhost_renderer_render(void* r) { static_cast<Renderer*>(static_cast<CudaRenderer*>(r))->Render(); }
int  lfhost_renderer_sample_count(void* r) { return static_cast<Renderer*>(static_cast<CudaRenderer*>(r))->GetSampleCount(); }
float lfhost_renderer_progress(void* r) { return static_cast<Renderer*>(static_cast<CudaRenderer*>(r))->GetProgress(); }
void lfhost_renderer_flush(void* r) { static_cast<CudaRenderer*>(r)->Flush(); }
void* lfhost_renderer_ctx(void* r) { return static_cast<CudaRenderer*>(r)->Context(); }
// GetOutputBufferHDR / GetOutputBuffer into caller memory (the interface allocates with new[]; copied and freed here)
int lfhost_renderer_output_hdr(void* r, float* out, int* w, int* h) {
    float* data = nullptr;
    static_cast<Renderer*>(static_cast<CudaRenderer*>(r))->GetOutputBufferHDR(&data, *w, *h);
    if (out) std::memcpy(out, data, (size_t)(*w) * (*h) * 3 * sizeof(float));
    delete[] data;
    return 0;
}
int lfhost_renderer_output_u8(void* r, unsigned char* out, int* w, int* h) {
    unsigned char* data = nullptr;
    static_cast<Renderer*>(static_cast<CudaRenderer*>(r))->GetOutputBuffer(&data, *w, *h);
    if (out) std::memcpy(out, data, (size_t)(*w) * (*h) * 3);
    delete[] data;
    return 0;
}
// The preview image (CudaRenderer::GetPreviewBufferHDR); with out == NULL only the size is returned.  -1 = no preview drawn.
int lfhost_renderer_preview_hdr(void* r, float* out, int* w, int* h) {
    float* data = nullptr;
    static_cast<CudaRenderer*>(r)->GetPreviewBufferHDR(&data, *w, *h);
    if (!data) return -1;
    if (out) std::memcpy(out, data, (size_t)(*w) * (*h) * 3 * sizeof(float));
    delete[] data;
    return 0;
}
void lfhost_set_preview(float scale, int use_dof) { GlobalState.previewScale = scale; GlobalState.useDofInPreview = use_dof != 0; }   // Main.cpp:525-529
// Main.cpp's loop (MainLoop :313-755 -> Update :160-234 -> Render :99-158) for `spp` samples: the auto-stop test
// `maxSamples + 1 == GetSampleCount()` comes first, then renderer->Update, then renderer->Render.
int lfhost_renderer_run(void* rv, int spp) {
    if (!static_cast<CudaRenderer*>(rv)->Ok()) return -1;   // a renderer whose Init failed never advances its sample counter
    Renderer* r = static_cast<Renderer*>(static_cast<CudaRenderer*>(rv));
    int steps = 0;
    while (true) {
        if (r->GetSampleCount() == spp + 1) break;
        r->Update(0.f);
        if (r->GetSampleCount() == spp + 1) break;
        r->Render();
        steps++;
    }
    return steps;
}

}  // extern "C"
