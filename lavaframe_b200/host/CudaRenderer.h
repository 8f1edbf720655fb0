// CudaRenderer.h — drop-in replacement for LavaFrame::TiledRenderer behind the reference's own Renderer
// interface (LavaFrame/Renderer.h:71-117).  Construct it where the reference constructs its renderer:
//
//     GlobalState.renderer = new CudaRenderer(GlobalState.scene, GlobalState.shadersDir);   // Main.cpp:91
//
// Same call protocol (Init once; every loop iteration Update(dt) then Render(); GetSampleCount() ==
// completed samples + 1; GetOutputBuffer[HDR] returns the image of the last COMPLETED sample, bottom row
// first, allocated with new[]), same tile / frame / sample counters as TiledRenderer.cpp:55-64,485-501, so the
// `frame` uniform that seeds the RNG takes the same values.  All rendering goes through the C ABI of
// include/lfcuda.h; there is no OpenGL and no CPU fallback.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "Renderer.h"
#include "lfcuda.h"

namespace LavaFrame
{
    class Scene;

    class CudaRenderer : public Renderer
    {
    public:
        CudaRenderer(Scene* scene, const std::string& shadersDirectory, int device = 0);
        // Several GPUs behind the one renderer the application constructs (Main.cpp:91): one context + host thread per device,
        // the frames of every batch dealt round-robin (spp split, disjoint RNG streams), the devices' accumulation buffers summed
        // inside the post-process pass of devices[0] over NVLink peer access when an output buffer is asked for (lfcuda_group_*).
        CudaRenderer(Scene* scene, const std::string& shadersDirectory, const std::vector<int>& devices);
        ~CudaRenderer();

        void Init() override;
        void Finish() override;
        void Render() override;
        void Present() const override {}                         // no window in the headless build
        void Update(float secondsElapsed) override;
        float GetProgress() const override;
        int GetSampleCount() const override;
        void GetOutputBuffer(unsigned char**, int& w, int& h) override;
        void GetOutputBufferHDR(float** data, int& w, int& h) override;
        uint32_t SetViewport(int, int) override { return 0; }    // returned a GL texture id
        uint32_t Denoise() override { return 0; }                // OIDN is not part of the path

        // Extras for drivers/tests (not part of the reference interface)
        lfcuda_ctx* Context() const { return ctx; }                 // the context of the first device (preview, probes)
        lfcuda_group* Group() const { return group; }
        int NumDevices() const { return (int)devices.size(); }
        // Instance edits: rebuild the TLAS on the device(s) from the matrices (lfcuda_update_instances_device) instead of uploading
        // the TLAS Scene::RebuildInstances built on the host.  Same nodes, bit for bit; off by default (also: LF_DEVICE_TLAS=1).
        void SetDeviceTlasRebuild(bool on) { deviceTlas = on; }
        bool Ok() const { return initialized && ctx != nullptr; }   // Init completed: scene uploaded, uniforms set, accumulation cleared
        const char* LastError() const;
        void Flush();                                            // execute every queued tile step now
        // The image Present()/SetViewport() display while the camera moves or before the first sample completes
        // (pathTraceTextureLowRes through postShader, TiledRenderer.cpp:361-364,558-562): w x h x 3 floats, bottom row
        // first, allocated with new[] like GetOutputBufferHDR.  *data is nullptr when no preview has been drawn.
        void GetPreviewBufferHDR(float** data, int& w, int& h);

    private:
        struct Step { int frame, tileX, tileY, sample; };
        void FlushCompletedSamples();
        void Execute(size_t count);
        bool UploadUniforms();
        void FailInit();

        lfcuda_group* group;             // one context per device; a single-GPU renderer is a group of one
        lfcuda_ctx* ctx;                 // = lfcuda_group_ctx(group, 0)
        std::vector<int> devices;
        std::vector<Step> pending;       // tile steps requested by Render() and not yet launched

        int tileX, tileY, numTilesX, numTilesY, tileWidth, tileHeight;
        int currentBuffer, frameCounter, sampleCounter;
        float pixelRatio;                // GlobalState.previewScale at Init (TiledRenderer.cpp:61)
        bool deviceTlas;
        bool haveUniforms;               // the uniforms last sent to the devices (Update sends only what changed)
        LfParams lastParams; LfCamera lastCam; LfPostParams lastPost;
        bool previewDof;                 // GlobalState.useDofInPreview at Init (#define USE_DOF, :90-91)
        int previewDepth;                // the preview shader's maxDepth uniform as Update last set it (:532)
        int previewW, previewH;          // size of the last preview drawn, 0 = none
        std::string error;
    };
}
