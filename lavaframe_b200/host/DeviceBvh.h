// DeviceBvh.h — RadeonRays::LfDeviceSplitBvh: the mesh BVH builder the reference constructs in Mesh.h:18, with its build step on the GPU.
//
// The reference builds every mesh's BVH on one host core at scene load (Scene::createBLAS, Scene.cpp:148-163 -> Mesh::BuildBVH,
// Mesh.cpp:93-111 -> Bvh::Build -> virtual BuildImpl; 8 s for C2's 869 880 triangles).  This subclass overrides BuildImpl: when the device
// build is switched on (lfhost_set_device_blas / LF_DEVICE_BLAS=1) it hands the triangle boxes to lfcuda_build_blas (csrc/lf_blas.cu), which
// returns the SAME tree node for node, and rebuilds the Bvh's protected members from it (m_nodes with lc / rc pointers, m_root, m_nodecnt,
// m_packed_indices, m_height) - so everything downstream runs unchanged on it: BvhTranslator::ProcessBLAS (bvh_translator.cpp:91-114), the
// vertIndices of Scene.cpp:196-209, createTLAS' Bounds().  Switched off (the default) it IS SplitBvh::BuildImpl.
//
// How the reference gets to construct it without an edit: lf_mesh_bvh_hook.h, force-included for the one translation unit that holds the one
// `new Mesh` (Scene.cpp:31).  INTEGRATION.md shows the one-line change a maintainer would make in Mesh.h instead.
#pragma once

#include "split_bvh.h"

namespace RadeonRays {

class LfDeviceSplitBvh : public SplitBvh {
public:
    LfDeviceSplitBvh(float traversal_cost, int num_bins, int max_split_depth, float min_overlap, float extra_refs_budget)
        : SplitBvh(traversal_cost, num_bins, max_split_depth, min_overlap, extra_refs_budget), cost_(traversal_cost), bins_(num_bins), split_depth_(max_split_depth) {}

protected:
    void BuildImpl(bbox const* bounds, int numbounds) override;

private:
    float cost_;
    int bins_, split_depth_;
};

}  // namespace RadeonRays

namespace lfhost {
struct BlasStats { int device_builds, host_builds, negative_zero_meshes; double device_ms, device_total_ms, host_ms; long long device_prims, host_prims; };
void SetDeviceBlas(int enable, int min_prims, int device);     // enable < 0: leave as is (environment: LF_DEVICE_BLAS, LF_DEVICE_BLAS_MIN)
BlasStats GetBlasStats(bool reset);
}  // namespace lfhost
