// lf_kernels.h — host-callable launchers of the kernels in lf_kernels.cu.
#pragma once

#include "lf_types.h"

namespace lf {

// Ray-sort experiment (off unless LF_SORT_RAYS is set; bit 0 = extend queue of bounces >= 1, bit 1 = shadow queue): the queue is
// counting-sorted by (origin cell, direction octant) before the traversal kernel, so that the rays a warp pulls walk similar nodes.  Only the ORDER in which slots are traced
// changes; every ray's result lands in its own slot, so images are identical bit for bit.
constexpr int kSortCellBits = 4;                                  // cells per axis = 16
constexpr int kSortBins = 1 << (3 * kSortCellBits + 3);           // x 8 direction octants = 32 768 bins
struct SortCtx {
    int*      sorted = nullptr;     // the sorted copy of the queue (k_shade keeps reading the unsorted one: its state loads stay coalesced)
    unsigned* keys = nullptr;       // bin of every queue entry
    unsigned* hist = nullptr;       // kSortBins counters -> exclusive offsets
    float lo[3] = {0, 0, 0}, inv[3] = {0, 0, 0};   // scene bounds -> cell index
    int mode = 0;                   // LF_SORT_RAYS bits
};

struct LaunchCtx {
    DevScene     scene;
    DevParams    params;
    PathSoA      soa;
    Queues       queues;
    DevCounters* counters;
    cudaStream_t stream;
    int sm_count;
    int persistent_blocks;   // CTAs of the persistent traversal kernels (multiple of the SM count)
    int stack_depth;         // traversal stack entries the scene needs (<= 64)
    bool cull, count;
    const SortCtx* sort = nullptr;   // non-null = sort the extend queue of bounces >= 1
};

void launch_generate(const LaunchCtx& L);
void launch_extend(const LaunchCtx& L, int depth);
void launch_shade(const LaunchCtx& L, int depth);
void launch_sample(const LaunchCtx& L, int depth);
bool shade_is_fused(const LaunchCtx& L);   // launch_shade runs both halves in one kernel, launch_sample is a no-op
void launch_shadow(const LaunchCtx& L, int depth);
bool shadow_before_sample();   // order of the two passes after launch_shade (see lf_kernels.cu LF_SHADOW_FIRST)
void launch_accumulate(const LaunchCtx& L, float* accum);
void launch_preview_store(const LaunchCtx& L, float* preview);
void launch_megakernel(const LaunchCtx& L);
void launch_export_hits(const LaunchCtx& L, float* t, int* tri, int* mat, int* emitter);
void launch_node_probe(cudaStream_t stream, const float4* nodes, unsigned num_nodes_pow2, int steps, unsigned* sink, int blocks);
void launch_read_probe(cudaStream_t stream, const float4* buf, size_t n4, int passes, float* sink, int blocks);
// Device-side TLAS rebuild (lf_tlas.cu).  All pointers are device memory.
struct TlasBuild {
    const float*  transforms;     // n x 16, the reference's Mat4 layout
    const float*  blas_box;       // n x 6: pmin, pmax of the instance's mesh BVH (the box of its BLAS root node)
    const int*    inst_blas_root; // n: flat node index of the instance's BLAS root (TLAS leaf LRLeaf.x)
    const int*    inst_blas_ref;  // n: the same as a packed child reference (lf_types.h)
    const int*    inst_mat;       // n: materialID (TLAS leaf LRLeaf.y)
    int n, top_index, inner_base; // instances; first flat TLAS node; packed index of the first TLAS inner node
    float*  flat_tlas;            // out: (2n - 1) x 9 floats, the reference's node layout, pre-order from top_index
    float4* packed_nodes;         // out: base of the packed inner-node array (the TLAS part is rewritten)
    float4* inst_records;         // out: kInstStride float4 per instance
    int*    result;               // out: [0] top reference, [1] inner nodes on the longest root-to-leaf chain
    float *bmin, *bmax, *cent;    // scratch: n x 3 each
    int*   prim;                  // scratch: n
    void  *queue0, *queue1;       // scratch: n requests each (tlas_request_bytes())
};
size_t tlas_request_bytes();
void launch_tlas_build(cudaStream_t stream, const TlasBuild& T, int sm_count);

constexpr int kMaxGroupDevices = 16;
void launch_post_sum(cudaStream_t stream, const float* const* accums, int naccum, float* out_f, unsigned char* out_u8, float* sum_out, int W, int H,
                     float inv, int tonemap, const LfPostParams& pp);
void launch_post(cudaStream_t stream, const float* accum, float* out_f, unsigned char* out_u8, int W, int H, float inv, int tonemap, const LfPostParams& pp);

}  // namespace lf
