// oracle/ref_bvh/lf_ref_bvh.cpp — TEST INFRASTRUCTURE: the reference's OWN mesh-BVH builder as a library (oracle/_ref/liblfrefbvh.so).
//
// Compiled by oracle/Makefile (`make ref`) together with the reference's unchanged thirdparty/RadeonRays/{split_bvh,bvh,bbox}.cpp, where they
// lie.  It constructs the builder exactly as LavaFrame/Mesh.h:18 does, runs Bvh::Build on the caller's triangle boxes (what Mesh::BuildBVH
// passes, Mesh.cpp:93-111) and writes the tree out in the pre-order of BvhTranslator::ProcessBLASNodes (bvh_translator.cpp:35-60), in the
// layout of lfcuda_build_blas (include/lfcuda.h).  It is the ground truth tests/test_blas_build.py and tests/test_blas_device_gpu.py hold
// the device builder against; nothing in the product links or calls it.
#include <cstdint>
#include <cstring>

#include "split_bvh.h"

namespace {
struct RefBlasProbe : RadeonRays::SplitBvh {
    RefBlasProbe() : RadeonRays::SplitBvh(2.0f, 64, 0, 0.001f, 0) {}      // Mesh.h:18
    int cur = 0;
    int Emit(const Node* nd, float* out) {
        const int k = cur++;
        float* r = out + 9 * (size_t)k;
        r[0] = nd->bounds.pmin.x; r[1] = nd->bounds.pmin.y; r[2] = nd->bounds.pmin.z;
        r[3] = nd->bounds.pmax.x; r[4] = nd->bounds.pmax.y; r[5] = nd->bounds.pmax.z;
        int32_t* ri = reinterpret_cast<int32_t*>(r + 6);
        if (nd->type == kLeaf) { ri[0] = nd->startidx; ri[1] = nd->numprims; ri[2] = 1; }
        else { ri[2] = 0; ri[0] = Emit(nd->lc, out); ri[1] = Emit(nd->rc, out); }
        return k;
    }
    int Flatten(float* out) { cur = 0; Emit(m_root, out); return (int)m_nodecnt; }
};
}  // namespace

// out_nodes: (2 n - 1) x 9 words, out_indices: n; info = {nodes, indices, height}
extern "C" __attribute__((visibility("default")))
int lfref_build_blas(const float* prim_bounds, int n, float* out_nodes, int32_t* out_indices, int32_t* info) {
    if (n < 1) return -1;
    static_assert(sizeof(RadeonRays::bbox) == 6 * sizeof(float), "bbox is {Vec3 pmin, pmax}");
    RefBlasProbe bvh;
    bvh.Build(reinterpret_cast<const RadeonRays::bbox*>(prim_bounds), n);
    info[0] = bvh.Flatten(out_nodes);
    info[1] = (int32_t)bvh.GetNumIndices();
    info[2] = bvh.GetHeight();
    memcpy(out_indices, bvh.GetIndices(), sizeof(int32_t) * bvh.GetNumIndices());
    return 0;
}
