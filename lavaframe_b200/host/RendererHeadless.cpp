// RendererHeadless.cpp — GL-free definitions of the out-of-line members of the reference's abstract
// LavaFrame::Renderer (declared in LavaFrame/Renderer.h:71-117, defined with OpenGL calls in
// LavaFrame/Renderer.cpp:23-205).  The headless build links THIS file instead of the reference's Renderer.cpp,
// so that CudaRenderer derives from the reference's own class without any OpenGL object being created.
// Inside the real LavaFrame tree this file is not needed: CudaRenderer.cpp is simply added next to
// TiledRenderer.cpp and keeps using Renderer.cpp.
#include "Config.h"
#include "Renderer.h"
#include "Scene.h"

namespace LavaFrame
{
    Renderer::Renderer(Scene* scene, const std::string& shadersDirectory)
        : BVHTex(0), vertexIndicesTex(0), verticesTex(0), normalsTex(0), materialsTex(0), transformsTex(0), lightsTex(0)
        , textureMapsArrayTex(0), hdrTex(0), hdrMarginalDistTex(0), hdrConditionalDistTex(0)
        , scene(scene), quad(nullptr), initialized(false)
        , numOfLights(scene->lights.size())
        , screenSize(scene->renderOptions.resolution)
        , shadersDirectory(shadersDirectory)
    {
    }

    Renderer::~Renderer() {}

    void Renderer::Finish() { initialized = false; }

    void Renderer::Init()
    {
        if (initialized) return;
        if (scene == nullptr) { printf("Error: No Scene Found\n"); return; }   // Renderer.cpp:81-85
        initialized = true;
    }

    void Renderer::Update(float) {}
}
