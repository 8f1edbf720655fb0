// lf_repack.h — host-side re-packing of the reference's flat scene arrays into the GPU layout of lf_types.h.
#pragma once

#include <string>
#include <vector>

#include "lf_types.h"

namespace lf {

struct PackedScene {
    std::vector<float4> nodes;      // 4 per inner node
    std::vector<float4> tris;       // kTriStride per triangle ref
    std::vector<float4> trinrm;     // 3 per triangle ref
    std::vector<int>    tri_vx;
    std::vector<float4> inst;       // kInstStride per instance
    std::vector<float4> lights;     // kLightStride per light
    int top_ref = 0;
    int stack_depth = 0;            // stack entries a traversal can need
    int num_inner = 0;
    int max_blas_height = 0;        // inner nodes on the longest chain of any instanced BLAS (stack_depth = 2 + TLAS height + this + 1)
    int num_blas_inner = 0;         // inner nodes below top_bvh_index: the packed index of the first TLAS inner node
};

// Returns false and fills `err` when the arrays are inconsistent or exceed a structural limit.
bool repack_scene(const LfSceneView& v, PackedScene& out, std::string& err);
// TLAS-only update (Renderer::Update, LavaFrame/Renderer.cpp:190-205): `nodes` is the full node array with
// the new TLAS range already patched in.
bool repack_instances(const float* nodes, int num_nodes, int top_index, const float* transforms, int num_instances,
                      PackedScene& out, std::string& err);

}  // namespace lf
