// scene_view.h — raw-pointer view of a loaded LavaFrame::Scene for the C ABI (include/lfcuda.h).
// Reads only public members of Scene (LavaFrame/Scene.h:68-101); copies nothing.
#pragma once
#include "lfcuda.h"

namespace LavaFrame { class Scene; }

namespace lfhost {
// Pointers stay valid while the Scene is alive and its vectors are not resized.
void MakeSceneView(const LavaFrame::Scene* scene, LfSceneView* view);
// RenderOptions + the #define selection of TiledRenderer::Init (TiledRenderer.cpp:78-91).
void MakeParams(const LavaFrame::Scene* scene, LfParams* params);
// The camera uniforms TiledRenderer::Update sets (TiledRenderer.cpp:507-513).
void MakeCamera(const LavaFrame::Scene* scene, LfCamera* camera);
}
