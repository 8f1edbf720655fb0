// scene_view.cpp — see scene_view.h.  Compiled against the reference's own headers.
#include "scene_view.h"

#include <cstring>

#include "Scene.h"
#include "Camera.h"

namespace lfhost {

using LavaFrame::Scene;

static_assert(sizeof(RadeonRays::BvhTranslator::Node) == 36, "BvhTranslator::Node must be 9 floats");
static_assert(sizeof(LavaFrame::Material) == 112, "Material must be 7 x vec4");
static_assert(sizeof(LavaFrame::Light) == 60, "Light must be 5 x vec3");
static_assert(sizeof(LavaFrame::Mat4) == 64, "Mat4 must be 16 floats");
static_assert(sizeof(LavaFrame::Indices) == 12, "Indices must be 3 ints");
static_assert(sizeof(LavaFrame::Vec4) == 16, "Vec4 must be 4 floats");

void MakeSceneView(const Scene* s, LfSceneView* v) {
    std::memset(v, 0, sizeof(*v));
    v->bvh_nodes = reinterpret_cast<const float*>(s->bvhTranslator.nodes.data());
    v->num_nodes = (int32_t)s->bvhTranslator.nodes.size();
    v->top_bvh_index = s->bvhTranslator.topLevelIndex;
    v->vert_indices = reinterpret_cast<const int32_t*>(s->vertIndices.data());
    v->num_tri_refs = (int32_t)s->vertIndices.size();
    v->vertices_uvx = reinterpret_cast<const float*>(s->verticesUVX.data());
    v->normals_uvy = reinterpret_cast<const float*>(s->normalsUVY.data());
    v->num_vertices = (int32_t)s->verticesUVX.size();
    v->transforms = reinterpret_cast<const float*>(s->transforms.data());
    v->num_instances = (int32_t)s->transforms.size();
    v->materials = reinterpret_cast<const float*>(s->materials.data());
    v->num_materials = (int32_t)s->materials.size();
    v->lights = s->lights.empty() ? nullptr : reinterpret_cast<const float*>(s->lights.data());
    v->num_lights = (int32_t)s->lights.size();
    if (!s->textures.empty()) {
        v->texture_maps = s->textureMapsArray.data();
        v->tex_width = s->texWidth;
        v->tex_height = s->texHeight;
        v->num_textures = (int32_t)s->textures.size();
    }
    if (s->hdrData != nullptr) {
        v->hdr_cols = s->hdrData->cols;
        v->hdr_marginal = reinterpret_cast<const float*>(s->hdrData->marginalDistData);
        v->hdr_conditional = reinterpret_cast<const float*>(s->hdrData->conditionalDistData);
        v->hdr_width = s->hdrData->width;
        v->hdr_height = s->hdrData->height;
    }
}

void MakeParams(const Scene* s, LfParams* p) {
    std::memset(p, 0, sizeof(*p));
    const LavaFrame::RenderOptions& ro = s->renderOptions;
    p->width = ro.resolution.x;
    p->height = ro.resolution.y;
    p->tile_width = ro.tileWidth;
    p->tile_height = ro.tileHeight;
    p->max_depth = ro.maxDepth;
    p->enable_rr = ro.enableRR ? 1 : 0;
    p->rr_depth = ro.RRDepth;
    p->use_envmap = (ro.useEnvMap && s->hdrData != nullptr) ? 1 : 0;
    p->use_constant_bg = ro.useConstantBg ? 1 : 0;
    p->bg_color[0] = ro.bgColor.x; p->bg_color[1] = ro.bgColor.y; p->bg_color[2] = ro.bgColor.z;
    p->hdr_multiplier = ro.hdrMultiplier;
}

void MakeCamera(const Scene* s, LfCamera* c) {
    const LavaFrame::Camera* cam = s->camera;
    c->position[0] = cam->position.x; c->position[1] = cam->position.y; c->position[2] = cam->position.z;
    c->right[0] = cam->right.x; c->right[1] = cam->right.y; c->right[2] = cam->right.z;
    c->up[0] = cam->up.x; c->up[1] = cam->up.y; c->up[2] = cam->up.z;
    c->forward[0] = cam->forward.x; c->forward[1] = cam->forward.y; c->forward[2] = cam->forward.z;
    c->fov = cam->fov;
    c->focal_dist = cam->focalDist;
    c->aperture = cam->aperture;
}

}  // namespace lfhost
