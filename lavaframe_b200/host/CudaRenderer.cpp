// CudaRenderer.cpp — see CudaRenderer.h.  The state machine restates TiledRenderer::Init/Update/Render
// (LavaFrame/TiledRenderer.cpp:48-64, 317-352, 421-521); GPU work is deferred and batched: Render() only
// queues its (frame, tile) step, and queued steps of COMPLETED samples are launched together (consecutive
// frames of one tile become a single lfcuda_render_frames call), which is unobservable through the interface
// because the reference also exposes only the image of the last completed sample (tileOutputTexture[1 - currentBuffer]).
#include "CudaRenderer.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "Camera.h"
#include "GlobalState.h"
#include "Scene.h"
#include "scene_view.h"

extern LavaFrameState GlobalState;

namespace LavaFrame
{
    CudaRenderer::CudaRenderer(Scene* scene, const std::string& shadersDirectory, int device_)
        : CudaRenderer(scene, shadersDirectory, std::vector<int>(1, device_))
    {
    }

    CudaRenderer::CudaRenderer(Scene* scene, const std::string& shadersDirectory, const std::vector<int>& devices_)
        : Renderer(scene, shadersDirectory)
        , group(nullptr), ctx(nullptr), devices(devices_)
        , tileX(-1), tileY(-1), numTilesX(-1), numTilesY(-1)
        , tileWidth(scene->renderOptions.tileWidth), tileHeight(scene->renderOptions.tileHeight)
        , currentBuffer(0), frameCounter(1), sampleCounter(0)
        , pixelRatio(1.0f), deviceTlas(getenv("LF_DEVICE_TLAS") && atoi(getenv("LF_DEVICE_TLAS")) != 0), haveUniforms(false), previewDof(false), previewDepth(2), previewW(0), previewH(0)
    {
    }

    CudaRenderer::~CudaRenderer()
    {
        this->Finish();                                      // frees the context whether or not Init completed
    }

    const char* CudaRenderer::LastError() const { return group ? lfcuda_group_last_error(group) : error.c_str(); }

    // Init failed after the context was created: keep the message, release the device, leave the renderer not initialised
    // (every entry point then returns at once, like the reference after a failed shader compile, Renderer.cpp:81-85).
    void CudaRenderer::FailInit()
    {
        error = lfcuda_group_last_error(group);
        printf("CudaRenderer: %s\n", error.c_str());
        if (group) lfcuda_group_destroy(group);
        group = nullptr;
        ctx = nullptr;
        initialized = false;
    }

    void CudaRenderer::Init()
    {
        if (initialized) return;
        Renderer::Init();
        if (!initialized) return;
        initialized = false;

        sampleCounter = 1;                                   // TiledRenderer.cpp:55-64
        currentBuffer = 0;
        frameCounter = 1;
        numTilesX = (int)ceil((float)screenSize.x / tileWidth);
        numTilesY = (int)ceil((float)screenSize.y / tileHeight);
        tileX = -1;
        tileY = numTilesY - 1;
        pixelRatio = GlobalState.previewScale;               // :61
        previewDof = GlobalState.useDofInPreview;            // :90-91
        previewDepth = scene->renderOptions.maxDepth;
        previewW = previewH = 0;

        haveUniforms = false;
        if (lfcuda_group_create(&group, devices.data(), (int)devices.size()) != 0) { group = nullptr; FailInit(); return; }
        ctx = lfcuda_group_ctx(group, 0);
        LfSceneView view;
        lfhost::MakeSceneView(scene, &view);
        if (lfcuda_group_upload_scene(group, &view) != 0) { FailInit(); return; }   // e.g. a BVH deeper than the 64-entry stack
        if (!UploadUniforms()) { FailInit(); return; }                               // e.g. path state does not fit the device
        if (lfcuda_group_clear(group) != 0) { FailInit(); return; }
        initialized = true;
    }

    bool CudaRenderer::UploadUniforms()
    {
        LfParams params;
        LfCamera cam;
        memset(&params, 0, sizeof params); memset(&cam, 0, sizeof cam);      // compared bytewise below
        lfhost::MakeParams(scene, &params);
        lfhost::MakeCamera(scene, &cam);
        const RenderOptions& ro = scene->renderOptions;      // postShader uniforms, TiledRenderer.cpp:539-553
        LfPostParams post;
        memset(&post, 0, sizeof post);
        post.use_ca = ro.useCA ? 1 : 0; post.use_ca_distortion = ro.useCADistortion ? 1 : 0;
        post.ca_distance = ro.caDistance; post.ca_p1 = ro.caP1; post.ca_p2 = ro.caP2; post.ca_p3 = ro.caP3;
        post.use_vignette = ro.useVignette ? 1 : 0; post.vignette_intensity = ro.vignetteIntensity; post.vignette_power = ro.vignettePower;
        // Update() runs once per tile step (4096 times for a 4096-spp single-tile render); the uniforms hardly ever change between
        // two of them, and every group call is a round trip to one worker thread per device: only what changed is sent.
        if (!haveUniforms || memcmp(&params, &lastParams, sizeof params) != 0) {
            if (lfcuda_group_set_params(group, &params) != 0) { printf("CudaRenderer: %s\n", lfcuda_group_last_error(group)); return false; }
            lastParams = params;
        }
        if (!haveUniforms || memcmp(&cam, &lastCam, sizeof cam) != 0) {
            if (lfcuda_group_set_camera(group, &cam) != 0) { printf("CudaRenderer: %s\n", lfcuda_group_last_error(group)); return false; }
            lastCam = cam;
        }
        if (!haveUniforms || memcmp(&post, &lastPost, sizeof post) != 0) {
            if (lfcuda_group_set_post(group, &post) != 0) { printf("CudaRenderer: %s\n", lfcuda_group_last_error(group)); return false; }
            lastPost = post;
        }
        haveUniforms = true;
        return true;
    }

    void CudaRenderer::Finish()
    {
        pending.clear();
        if (group) lfcuda_group_destroy(group);
        group = nullptr;
        ctx = nullptr;
        if (initialized) Renderer::Finish();
        initialized = false;
    }

    void CudaRenderer::Execute(size_t count)
    {
        // consecutive frames of the same tile -> one batched call
        size_t i = 0;
        while (i < count) {
            size_t j = i + 1;
            while (j < count && pending[j].tileX == pending[i].tileX && pending[j].tileY == pending[i].tileY &&
                   pending[j].frame == pending[j - 1].frame + 1) j++;
            if (lfcuda_group_render_frames(group, pending[i].frame, (int)(j - i), 1, pending[i].tileX, pending[i].tileY) != 0)
                printf("CudaRenderer: %s\n", lfcuda_group_last_error(group));
            i = j;
        }
        pending.erase(pending.begin(), pending.begin() + count);
    }

    void CudaRenderer::FlushCompletedSamples()
    {
        size_t n = 0;
        while (n < pending.size() && pending[n].sample < sampleCounter) n++;
        if (n) Execute(n);
    }

    void CudaRenderer::Flush() { if (!pending.empty()) Execute(pending.size()); }

    void CudaRenderer::Render()
    {
        if (!initialized) { printf("Renderer is not initialized.\n"); return; }   // TiledRenderer.cpp:319-323
        if (scene->camera->isMoving || scene->instancesModified) {
            // previewEngineShader into previewFBO (:327-333); glViewport truncates the float sizes to GLsizei
            const int pw = (int)(screenSize.x * pixelRatio), ph = (int)(screenSize.y * pixelRatio);
            if (lfcuda_render_preview(ctx, pw, ph, previewDepth, previewDof ? 1 : 0) != 0)
                printf("CudaRenderer: %s\n", lfcuda_last_error(ctx));
            else { previewW = pw; previewH = ph; }
            scene->instancesModified = false;
            return;
        }
        pending.push_back(Step{frameCounter, tileX, tileY, sampleCounter});
        if (pending.size() >= 256) FlushCompletedSamples();   // launches are asynchronous: the devices render while this loop goes on
    }

    float CudaRenderer::GetProgress() const
    {
        return float((numTilesY - tileY - 1) * numTilesX + tileX) / float(numTilesX * numTilesY);   // TiledRenderer.cpp:377-380
    }

    int CudaRenderer::GetSampleCount() const { return sampleCounter; }

    void CudaRenderer::Update(float secondsElapsed)
    {
        if (!initialized) return;
        if (scene->instancesModified && deviceTlas) {
            std::vector<int32_t> mats(scene->meshInstances.size());
            for (size_t i = 0; i < mats.size(); i++) mats[i] = scene->meshInstances[i].materialID;
            if (lfcuda_group_update_instances_device(group, reinterpret_cast<const float*>(scene->transforms.data()), (int)scene->transforms.size(),
                                                     reinterpret_cast<const float*>(scene->materials.data()), (int)scene->materials.size(), mats.data()) != 0)
                printf("CudaRenderer: %s\n", lfcuda_group_last_error(group));
        } else if (scene->instancesModified) {   // Renderer::Update, Renderer.cpp:190-205
            int index = scene->bvhTranslator.topLevelIndex;
            int total = (int)scene->bvhTranslator.nodes.size();
            if (lfcuda_group_update_instances(group, reinterpret_cast<const float*>(scene->transforms.data()), (int)scene->transforms.size(),
                                              reinterpret_cast<const float*>(scene->materials.data()), (int)scene->materials.size(),
                                              reinterpret_cast<const float*>(&scene->bvhTranslator.nodes[index]), index, total - index) != 0)
                printf("CudaRenderer: %s\n", lfcuda_group_last_error(group));
        }
        if (scene->camera->isMoving || scene->instancesModified) {   // TiledRenderer.cpp:471-484
            tileX = -1;
            tileY = numTilesY - 1;
            sampleCounter = 1;
            frameCounter = 1;
            pending.clear();
            lfcuda_group_clear(group);
        } else {                                                     // :485-501
            frameCounter++;
            tileX++;
            if (tileX >= numTilesX) {
                tileX = 0;
                tileY--;
                if (tileY < 0) {
                    tileX = 0;
                    tileY = numTilesY - 1;
                    sampleCounter++;
                    currentBuffer = 1 - currentBuffer;
                }
            }
        }
        UploadUniforms();                                            // :505-521 (camera, maxDepth, hdrMultiplier, bgColor ...)
        previewDepth = (scene->camera->isMoving || scene->instancesModified) ? 2 : scene->renderOptions.maxDepth;   // :532
    }

    void CudaRenderer::GetPreviewBufferHDR(float** data, int& w, int& h)
    {
        w = previewW; h = previewH;
        *data = nullptr;
        if (!initialized || previewW < 1) return;
        *data = new float[(size_t)w * h * 3];
        if (lfcuda_read_preview(ctx, scene->renderOptions.tonemapIndex, scene->camera->isMoving ? 1 : 0, *data) != 0)
            printf("CudaRenderer: %s\n", lfcuda_last_error(ctx));
    }

    void CudaRenderer::GetOutputBufferHDR(float** data, int& w, int& h)
    {
        w = scene->renderOptions.resolution.x;                       // TiledRenderer.cpp:399-414
        h = scene->renderOptions.resolution.y;
        *data = new float[(size_t)w * h * 3]();                     // zero-filled: a renderer that failed to initialise shows black
        if (!initialized) return;
        FlushCompletedSamples();
        int completed = sampleCounter - 1;                           // what tileOutputTexture[1 - currentBuffer] holds
        float inv = 1.0f / (float)(completed > 0 ? completed : 1);   // invSampleCounter, TiledRenderer.cpp:542
        if (lfcuda_group_read_output(group, inv, scene->renderOptions.tonemapIndex, *data) != 0)
            printf("CudaRenderer: %s\n", lfcuda_group_last_error(group));
    }

    void CudaRenderer::GetOutputBuffer(unsigned char** data, int& w, int& h)
    {
        w = scene->renderOptions.resolution.x;                       // TiledRenderer.cpp:382-397
        h = scene->renderOptions.resolution.y;
        *data = new unsigned char[(size_t)w * h * 3]();
        if (!initialized) return;
        FlushCompletedSamples();
        int completed = sampleCounter - 1;
        float inv = 1.0f / (float)(completed > 0 ? completed : 1);
        if (lfcuda_group_read_output_u8(group, inv, scene->renderOptions.tonemapIndex, *data) != 0)
            printf("CudaRenderer: %s\n", lfcuda_group_last_error(group));
    }
}
