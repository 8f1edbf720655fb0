// scenepack.h — ".lfpack": a flat binary dump of exactly what the reference uploads to the GPU.
//
// The arrays are the ones Renderer::Init turns into GL textures (LavaFrame/Renderer.cpp:87-185),
// byte for byte as Scene::CreateAccelerationStructures leaves them (LavaFrame/Scene.cpp:180-231),
// followed by the RenderOptions / Camera values TiledRenderer passes as uniforms
// (LavaFrame/TiledRenderer.cpp:222-227,505-521).  A pack lets the CUDA path, the CPU oracle and the
// tests consume the same inputs on a machine where the reference sources do not exist.
//
// Layout (little endian):  char magic[8] "LFPACK01"; int32 ihdr[32]; float fhdr[32]; then the
// arrays in the order of LfSceneView.  Header-only, no dependency on the reference.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "lfcuda.h"

namespace lfpack {

enum IHdr {
    I_NUM_NODES = 0, I_TOP_INDEX, I_NUM_TRI_REFS, I_NUM_VERTICES, I_NUM_INSTANCES, I_NUM_MATERIALS,
    I_NUM_LIGHTS, I_TEX_W, I_TEX_H, I_NUM_TEX, I_HDR_W, I_HDR_H,
    I_WIDTH, I_HEIGHT, I_TILE_W, I_TILE_H, I_MAX_DEPTH, I_ENABLE_RR, I_RR_DEPTH, I_USE_ENVMAP,
    I_USE_CONST_BG, I_TONEMAP, I_COUNT = 32
};
enum FHdr {
    F_BG = 0 /*3*/, F_HDR_MULT = 3, F_CAM_POS = 4 /*3*/, F_CAM_RIGHT = 7, F_CAM_UP = 10, F_CAM_FWD = 13,
    F_CAM_FOV = 16, F_CAM_FOCAL = 17, F_CAM_APERTURE = 18, F_COUNT = 32
};

struct ScenePack {
    int32_t ihdr[I_COUNT] = {0};
    float   fhdr[F_COUNT] = {0};
    std::vector<float>   nodes, vertices, normals, transforms, materials, lights;
    std::vector<int32_t> vert_indices;
    std::vector<uint8_t> textures;
    std::vector<float>   hdr_cols, hdr_marginal, hdr_conditional;

    LfSceneView view() const {
        LfSceneView v;
        std::memset(&v, 0, sizeof(v));
        v.bvh_nodes = nodes.data();        v.num_nodes = ihdr[I_NUM_NODES];  v.top_bvh_index = ihdr[I_TOP_INDEX];
        v.vert_indices = vert_indices.data(); v.num_tri_refs = ihdr[I_NUM_TRI_REFS];
        v.vertices_uvx = vertices.data();  v.normals_uvy = normals.data();   v.num_vertices = ihdr[I_NUM_VERTICES];
        v.transforms = transforms.data();  v.num_instances = ihdr[I_NUM_INSTANCES];
        v.materials = materials.data();    v.num_materials = ihdr[I_NUM_MATERIALS];
        v.lights = lights.empty() ? nullptr : lights.data(); v.num_lights = ihdr[I_NUM_LIGHTS];
        v.texture_maps = textures.empty() ? nullptr : textures.data();
        v.tex_width = ihdr[I_TEX_W]; v.tex_height = ihdr[I_TEX_H]; v.num_textures = ihdr[I_NUM_TEX];
        v.hdr_cols = hdr_cols.empty() ? nullptr : hdr_cols.data();
        v.hdr_marginal = hdr_marginal.empty() ? nullptr : hdr_marginal.data();
        v.hdr_conditional = hdr_conditional.empty() ? nullptr : hdr_conditional.data();
        v.hdr_width = ihdr[I_HDR_W]; v.hdr_height = ihdr[I_HDR_H];
        return v;
    }
    LfParams params() const {
        LfParams p;
        std::memset(&p, 0, sizeof(p));
        p.width = ihdr[I_WIDTH]; p.height = ihdr[I_HEIGHT];
        p.tile_width = ihdr[I_TILE_W]; p.tile_height = ihdr[I_TILE_H];
        p.max_depth = ihdr[I_MAX_DEPTH]; p.enable_rr = ihdr[I_ENABLE_RR]; p.rr_depth = ihdr[I_RR_DEPTH];
        p.use_envmap = ihdr[I_USE_ENVMAP]; p.use_constant_bg = ihdr[I_USE_CONST_BG];
        for (int k = 0; k < 3; k++) p.bg_color[k] = fhdr[F_BG + k];
        p.hdr_multiplier = fhdr[F_HDR_MULT];
        return p;
    }
    LfCamera camera() const {
        LfCamera c;
        for (int k = 0; k < 3; k++) {
            c.position[k] = fhdr[F_CAM_POS + k]; c.right[k] = fhdr[F_CAM_RIGHT + k];
            c.up[k] = fhdr[F_CAM_UP + k];        c.forward[k] = fhdr[F_CAM_FWD + k];
        }
        c.fov = fhdr[F_CAM_FOV]; c.focal_dist = fhdr[F_CAM_FOCAL]; c.aperture = fhdr[F_CAM_APERTURE];
        return c;
    }
};

// Fill a pack from raw views (copies).
inline void from_views(ScenePack& p, const LfSceneView& v, const LfParams& prm, const LfCamera& cam, int tonemap_index) {
    p.ihdr[I_NUM_NODES] = v.num_nodes; p.ihdr[I_TOP_INDEX] = v.top_bvh_index; p.ihdr[I_NUM_TRI_REFS] = v.num_tri_refs;
    p.ihdr[I_NUM_VERTICES] = v.num_vertices; p.ihdr[I_NUM_INSTANCES] = v.num_instances;
    p.ihdr[I_NUM_MATERIALS] = v.num_materials; p.ihdr[I_NUM_LIGHTS] = v.num_lights;
    p.ihdr[I_TEX_W] = v.tex_width; p.ihdr[I_TEX_H] = v.tex_height; p.ihdr[I_NUM_TEX] = v.num_textures;
    p.ihdr[I_HDR_W] = v.hdr_width; p.ihdr[I_HDR_H] = v.hdr_height;
    p.ihdr[I_WIDTH] = prm.width; p.ihdr[I_HEIGHT] = prm.height; p.ihdr[I_TILE_W] = prm.tile_width; p.ihdr[I_TILE_H] = prm.tile_height;
    p.ihdr[I_MAX_DEPTH] = prm.max_depth; p.ihdr[I_ENABLE_RR] = prm.enable_rr; p.ihdr[I_RR_DEPTH] = prm.rr_depth;
    p.ihdr[I_USE_ENVMAP] = prm.use_envmap; p.ihdr[I_USE_CONST_BG] = prm.use_constant_bg; p.ihdr[I_TONEMAP] = tonemap_index;
    for (int k = 0; k < 3; k++) {
        p.fhdr[F_BG + k] = prm.bg_color[k];
        p.fhdr[F_CAM_POS + k] = cam.position[k]; p.fhdr[F_CAM_RIGHT + k] = cam.right[k];
        p.fhdr[F_CAM_UP + k] = cam.up[k];        p.fhdr[F_CAM_FWD + k] = cam.forward[k];
    }
    p.fhdr[F_HDR_MULT] = prm.hdr_multiplier;
    p.fhdr[F_CAM_FOV] = cam.fov; p.fhdr[F_CAM_FOCAL] = cam.focal_dist; p.fhdr[F_CAM_APERTURE] = cam.aperture;
    p.nodes.assign(v.bvh_nodes, v.bvh_nodes + (size_t)9 * v.num_nodes);
    p.vert_indices.assign(v.vert_indices, v.vert_indices + (size_t)3 * v.num_tri_refs);
    p.vertices.assign(v.vertices_uvx, v.vertices_uvx + (size_t)4 * v.num_vertices);
    p.normals.assign(v.normals_uvy, v.normals_uvy + (size_t)4 * v.num_vertices);
    p.transforms.assign(v.transforms, v.transforms + (size_t)16 * v.num_instances);
    p.materials.assign(v.materials, v.materials + (size_t)28 * v.num_materials);
    if (v.num_lights > 0) p.lights.assign(v.lights, v.lights + (size_t)15 * v.num_lights);
    if (v.num_textures > 0)
        p.textures.assign(v.texture_maps, v.texture_maps + (size_t)4 * v.tex_width * v.tex_height * v.num_textures);
    if (v.hdr_cols) {
        size_t n = (size_t)v.hdr_width * v.hdr_height;
        p.hdr_cols.assign(v.hdr_cols, v.hdr_cols + 3 * n);
        p.hdr_marginal.assign(v.hdr_marginal, v.hdr_marginal + (size_t)2 * v.hdr_height);
        p.hdr_conditional.assign(v.hdr_conditional, v.hdr_conditional + 2 * n);
    }
}

template <class T> inline bool wr(FILE* f, const std::vector<T>& v) {
    return v.empty() || fwrite(v.data(), sizeof(T), v.size(), f) == v.size();
}
template <class T> inline bool rd(FILE* f, std::vector<T>& v, size_t n) {
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}

inline bool write(const std::string& path, const ScenePack& p) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = fwrite("LFPACK01", 1, 8, f) == 8 && fwrite(p.ihdr, 4, I_COUNT, f) == I_COUNT && fwrite(p.fhdr, 4, F_COUNT, f) == F_COUNT;
    ok = ok && wr(f, p.nodes) && wr(f, p.vert_indices) && wr(f, p.vertices) && wr(f, p.normals) && wr(f, p.transforms) &&
         wr(f, p.materials) && wr(f, p.lights) && wr(f, p.textures) && wr(f, p.hdr_cols) && wr(f, p.hdr_marginal) &&
         wr(f, p.hdr_conditional);
    fclose(f);
    return ok;
}

inline bool read(const std::string& path, ScenePack& p, std::string* err = nullptr) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { if (err) *err = "cannot open " + path; return false; }
    char magic[8];
    bool ok = fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "LFPACK01", 8) == 0 &&
              fread(p.ihdr, 4, I_COUNT, f) == I_COUNT && fread(p.fhdr, 4, F_COUNT, f) == F_COUNT;
    if (ok) {
        const int32_t* h = p.ihdr;
        size_t hdrn = (size_t)h[I_HDR_W] * h[I_HDR_H];
        ok = rd(f, p.nodes, (size_t)9 * h[I_NUM_NODES]) && rd(f, p.vert_indices, (size_t)3 * h[I_NUM_TRI_REFS]) &&
             rd(f, p.vertices, (size_t)4 * h[I_NUM_VERTICES]) && rd(f, p.normals, (size_t)4 * h[I_NUM_VERTICES]) &&
             rd(f, p.transforms, (size_t)16 * h[I_NUM_INSTANCES]) && rd(f, p.materials, (size_t)28 * h[I_NUM_MATERIALS]) &&
             rd(f, p.lights, (size_t)15 * h[I_NUM_LIGHTS]) &&
             rd(f, p.textures, (size_t)4 * h[I_TEX_W] * h[I_TEX_H] * h[I_NUM_TEX]) && rd(f, p.hdr_cols, 3 * hdrn) &&
             rd(f, p.hdr_marginal, hdrn ? (size_t)2 * h[I_HDR_H] : 0) && rd(f, p.hdr_conditional, 2 * hdrn);
    }
    fclose(f);
    if (!ok && err) *err = "bad or truncated pack " + path;
    return ok;
}

}  // namespace lfpack
