// lf_scenepack — load a LavaFrame scene file with the reference's own loader + BVH builder
// (LoadSceneFromFile, LavaFrame/Loader.cpp:45-386; Scene::CreateAccelerationStructures, Scene.cpp:180-231)
// and dump the flattened arrays as an .lfpack (scenepack.h).
//
//   lf_scenepack <scene file> <out.lfpack> [--info]
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "Scene.h"
#include "Loader.h"
#include "GlobalState.h"
#include "../scene_view.h"
#include "../scenepack.h"
#include "../DeviceBvh.h"

using namespace LavaFrame;
extern LavaFrameState GlobalState;

static int quiet_log(const char*, ...) { return 0; }

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: lf_scenepack <scene> <out.lfpack> [--info]\n");
        return 2;
    }
    bool info = argc > 3 && !strcmp(argv[3], "--info");
    if (!info) LavaFrame::Log = quiet_log;
    RenderOptions ro;
    ro.tonemapIndex = 0;          // linear output: parity work compares radiance, not tonemapped values
    ro.useVignette = false;       // the ctor leaves the vignette members uninitialised (Renderer.h:41-43)
    ro.vignetteIntensity = 0.f;
    ro.vignettePower = 1.f;
    Scene* scene = new Scene();
    GlobalState.scene = scene;
    bool loaded = false;
    try { loaded = LoadSceneFromFile(argv[1], scene, ro); }
    catch (const std::exception& e) { fprintf(stderr, "lf_scenepack: %s\n", e.what()); }      // LF_DEVICE_BLAS=1 without a usable GPU
    if (!loaded) {
        fprintf(stderr, "lf_scenepack: cannot load %s\n", argv[1]);
        return 1;
    }
    scene->renderOptions = ro;    // Main.cpp:966-969
    scene->camera->isMoving = false;

    LfSceneView view; LfParams params; LfCamera cam;
    lfhost::MakeSceneView(scene, &view);
    lfhost::MakeParams(scene, &params);
    lfhost::MakeCamera(scene, &cam);
    lfpack::ScenePack pack;
    lfpack::from_views(pack, view, params, cam, ro.tonemapIndex);
    if (!lfpack::write(argv[2], pack)) {
        fprintf(stderr, "lf_scenepack: cannot write %s\n", argv[2]);
        return 1;
    }
    printf("%s: %dx%d depth %d rr %d/%d env %d | meshes %zu instances %d materials %d lights %d textures %d | "
           "nodes %d top %d tri_refs %d vertices %d | hdr %dx%d\n",
           argv[1], params.width, params.height, params.max_depth, params.enable_rr, params.rr_depth, params.use_envmap,
           scene->meshes.size(), view.num_instances, view.num_materials, view.num_lights, view.num_textures,
           view.num_nodes, view.top_bvh_index, view.num_tri_refs, view.num_vertices, view.hdr_width, view.hdr_height);
    const lfhost::BlasStats bs = lfhost::GetBlasStats(false);     // LF_DEVICE_BLAS=1: mesh BVHs built on the GPU (DeviceBvh.h)
    printf("mesh BVH builds: %d on the device (%lld triangles, %.2f ms in kernels, %.2f ms with allocation and copies; %d with -0.0 bounds), %d on the host (%lld triangles, %.2f ms)\n",
           bs.device_builds, bs.device_prims, bs.device_ms, bs.device_total_ms, bs.negative_zero_meshes, bs.host_builds, bs.host_prims, bs.host_ms);
    return 0;
}
