// lf_tlas.cu — the top level of the two-level BVH rebuilt ON THE DEVICE after an instance edit (first slice).
//
// What the reference does on the host for every edited frame (Scene::RebuildInstances, LavaFrame/Scene.cpp:165-178):
//   createTLAS (Scene.cpp:106-146)          world-space box of every instance = its mesh's BLAS bounds through the instance matrix
//   sceneBvh = new Bvh(10.0f, 64, false)    RadeonRays' plain BVH WITHOUT the SAH (usesah = false): every node splits its primitives at
//   Bvh::Build / BuildNode                  the centre of their centroid box along its longest axis with an in-place two-pointer
//   (thirdparty/RadeonRays/bvh.cpp:40-243)  partition (direction alternating with (numprims + startidx) & 1), halving the range when
//                                           the partition leaves a side empty; leaves hold exactly one instance
//   BvhTranslator::UpdateTLAS               pre-order flattening into the node array from topLevelIndex on
//   (bvh_translator.cpp:61-89,134-140)      ({bboxmin, bboxmax, LRLeaf} per node; TLAS leaf = (BLAS root, materialID, -instance - 1))
// and then re-uploads those nodes (Renderer::Update, Renderer.cpp:190-205).  Here the same tree is produced on the GPU from the
// instance matrices alone, NODE FOR NODE: the box arithmetic restates the host's fp32 operations (std::min / std::max operand order,
// left-to-right sums, (pmax + pmin) * 0.5f), and the partition is the reference's sequential one executed literally, one thread per
// node, level by level - so the order of the primitive indices, which decides the halving fallback and the leaf order, is the
// reference's.  Since every leaf holds one instance, a subtree of k instances has 2k - 1 nodes, and pre-order positions (and the
// positions of inner nodes among them) follow from the counts alone; no serial numbering pass is needed.
// Outputs: the flat TLAS nodes in the reference's own layout (what the parity tests compare with the host build), the re-packed
// inner nodes the traversal kernels walk (lf_types.h), and the instance records (inverse matrices etc., lf_repack.cpp build_instances).
// The critical path is the root's partition (n sequential steps), about a millisecond per thousand instances; BLAS builds stay on
// the host (RadeonRays' sequential spatial-split builder, Mesh.h:18).
#include <cfloat>

#include "lf_kernels.h"
#include "lf_matrix.h"

namespace lf {

namespace {

struct Box { float mn[3], mx[3]; };
__device__ __forceinline__ float smin(float a, float b) { return (b < a) ? b : a; }     // std::min(a, b)
__device__ __forceinline__ float smax(float a, float b) { return (a < b) ? b : a; }     // std::max(a, b)
__device__ __forceinline__ void box_clear(Box& b) { for (int k = 0; k < 3; k++) { b.mn[k] = FLT_MAX; b.mx[k] = -FLT_MAX; } }     // bbox()
__device__ __forceinline__ void grow_point(Box& b, const float* p) { for (int k = 0; k < 3; k++) { b.mn[k] = smin(b.mn[k], p[k]); b.mx[k] = smax(b.mx[k], p[k]); } }
__device__ __forceinline__ void grow_box(Box& b, const float* mn, const float* mx) {
    for (int k = 0; k < 3; k++) { b.mn[k] = smin(b.mn[k], mn[k]); b.mx[k] = smax(b.mx[k], mx[k]); }
}
__device__ __forceinline__ int maxdim(const Box& b) {                                    // bbox::maxdim, bbox.h:71-83
    float ex = b.mx[0] - b.mn[0], ey = b.mx[1] - b.mn[1], ez = b.mx[2] - b.mn[2];
    if (ex >= ey && ex >= ez) return 0;
    if (ey >= ex && ey >= ez) return 1;
    if (ez >= ex && ez >= ey) return 2;
    return 0;
}

struct Req {                 // Bvh::SplitRequest (bvh.h) + where the node lands
    int start, count;
    int pre;                 // pre-order position within the TLAS (flat index = top_index + pre)
    int rank;                // position among the TLAS' inner nodes in pre-order (packed index = inner_base + rank)
    int level;
    Box bounds, cbounds;
};

}  // namespace

// ---- instance world boxes + centroids (Scene::createTLAS, Scene.cpp:113-142; Bvh::BuildImpl, bvh.cpp:364-371)
__global__ void k_tlas_bounds(TlasBuild T) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T.n; i += gridDim.x * blockDim.x) {
        const float* m = T.transforms + 16 * (size_t)i;
        const float* bb = T.blas_box + 6 * (size_t)i;                                    // meshes[meshID]->bvh->Bounds() = the box of the BLAS root node
        float mn[3], mx[3];
        for (int k = 0; k < 3; k++) {
            float xa = m[0 + k] * bb[0], xb = m[0 + k] * bb[3];                          // right * minBound.x, right * maxBound.x
            float ya = m[4 + k] * bb[1], yb = m[4 + k] * bb[4];
            float za = m[8 + k] * bb[2], zb = m[8 + k] * bb[5];
            mn[k] = ((smin(xa, xb) + smin(ya, yb)) + smin(za, zb)) + m[12 + k];
            mx[k] = ((smax(xa, xb) + smax(ya, yb)) + smax(za, zb)) + m[12 + k];
        }
        for (int k = 0; k < 3; k++) {
            T.bmin[3 * i + k] = mn[k]; T.bmax[3 * i + k] = mx[k];
            T.cent[3 * i + k] = (mx[k] + mn[k]) * 0.5f;                                  // bbox::center
        }
        T.prim[i] = i;                                                                   // std::iota(m_indices)
    }
}

// ---- instance records (lf_repack.cpp build_instances; the inverses are the shared text of lf_matrix.h)
__global__ void k_instance_records(TlasBuild T) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < T.n; k += gridDim.x * blockDim.x) {
        const float* t = T.transforms + 16 * (size_t)k;
        float m[4][4], inv[4][4];
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++) m[c][r] = t[4 * c + r];
        inverse4(m, inv);
        float m3[3][3], i3[3][3];
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++) m3[c][r] = m[c][r];
        inverse3(m3, i3);
        float4* ip = T.inst_records + (size_t)kInstStride * k;
        for (int r = 0; r < 3; r++) ip[r] = make_float4(inv[0][r], inv[1][r], inv[2][r], inv[3][r]);
        ip[3] = make_float4(__int_as_float(T.inst_blas_ref[k]), __int_as_float(T.inst_mat[k]), 0.f, 0.f);
        for (int r = 0; r < 3; r++) ip[4 + r] = make_float4(m[0][r], m[1][r], m[2][r], m[3][r]);
        for (int r = 0; r < 3; r++) ip[7 + r] = make_float4(i3[r][0], i3[r][1], i3[r][2], 0.f);
    }
}

// ---- the build: one CTA, one thread per node of the current level
namespace {

__device__ void write_flat(const TlasBuild& T, int pre, const Box& b, float x, float y, float z) {
    float* o = T.flat_tlas + 9 * (size_t)pre;
    o[0] = b.mn[0]; o[1] = b.mn[1]; o[2] = b.mn[2]; o[3] = b.mx[0]; o[4] = b.mx[1]; o[5] = b.mx[2];
    o[6] = x; o[7] = y; o[8] = z;
}
__device__ void write_leaf(const TlasBuild& T, int pre, const Box& b, int inst) {        // ProcessTLASNodes, bvh_translator.cpp:73-81
    write_flat(T, pre, b, (float)T.inst_blas_root[inst], (float)T.inst_mat[inst], (float)(-inst - 1));
}
__device__ int tlas_leaf_ref(int inst) { return kRefLeafBit | kRefTlasBit | inst; }

// Bvh::BuildNode without the SAH (bvh.cpp:68-243 with m_usesah == false), for a request of >= 2 primitives.
__device__ void split_node(const TlasBuild& T, const Req& req, Req& L, Req& R) {
    const int axis = maxdim(req.cbounds);
    const float border = (req.cbounds.mx[axis] + req.cbounds.mn[axis]) * 0.5f;           // centroid_bounds.center()[axis]
    Box lb, rb, lcb, rcb;
    box_clear(lb); box_clear(rb); box_clear(lcb); box_clear(rcb);
    int* prim = T.prim;
    auto growL = [&](int p) { grow_box(lb, T.bmin + 3 * p, T.bmax + 3 * p); grow_point(lcb, T.cent + 3 * p); };
    auto growR = [&](int p) { grow_box(rb, T.bmin + 3 * p, T.bmax + 3 * p); grow_point(rcb, T.cent + 3 * p); };
    auto cen = [&](int idx) { return T.cent[3 * prim[idx] + axis]; };
    int splitidx = req.start;
    const bool near2far = ((req.count + req.start) & 0x1) != 0;
    if (req.cbounds.mx[axis] - req.cbounds.mn[axis] > 0.f) {
        int first = req.start, last = req.start + req.count;
        if (near2far) {
            while (true) {
                while ((first != last) && cen(first) < border) { growL(prim[first]); ++first; }
                if (first == last--) break;
                growR(prim[first]);
                while ((first != last) && cen(last) >= border) { growR(prim[last]); --last; }
                if (first == last) break;
                growL(prim[last]);
                int tmp = prim[first]; prim[first] = prim[last]; prim[last] = tmp;
                first++;
            }
        } else {
            while (true) {
                while ((first != last) && cen(first) >= border) { growL(prim[first]); ++first; }
                if (first == last--) break;
                growR(prim[first]);
                while ((first != last) && cen(last) < border) { growR(prim[last]); --last; }
                if (first == last) break;
                growL(prim[last]);
                int tmp = prim[first]; prim[first] = prim[last]; prim[last] = tmp;
                first++;
            }
        }
        splitidx = first;
    }
    if (splitidx == req.start || splitidx == req.start + req.count) {                    // a side stayed empty: halve the range (the boxes
        splitidx = req.start + (req.count >> 1);                                         // keep what the partition already grew, as in the reference)
        for (int i = req.start; i < splitidx; ++i) growL(prim[i]);
        for (int i = splitidx; i < req.start + req.count; ++i) growR(prim[i]);
    }
    L.start = req.start; L.count = splitidx - req.start; L.bounds = lb; L.cbounds = lcb; L.level = req.level + 1;
    R.start = splitidx; R.count = req.count - (splitidx - req.start); R.bounds = rb; R.cbounds = rcb; R.level = req.level + 1;
    L.pre = req.pre + 1; L.rank = req.rank + 1;                                          // pre-order: node, left subtree (2 kl - 1 nodes, kl - 1 inner), right
    R.pre = req.pre + 1 + (2 * L.count - 1); R.rank = req.rank + 1 + (L.count - 1);
}

}  // namespace

__global__ void __launch_bounds__(256) k_tlas_build(TlasBuild T) {
    __shared__ int curCount, nextCount, height;
    Req* cur = reinterpret_cast<Req*>(T.queue0);
    Req* nxt = reinterpret_cast<Req*>(T.queue1);
    if (threadIdx.x == 0) {
        // Bvh::Build (bvh.cpp:40-49): scene bounds; BuildImpl (:364-373): centroid bounds; the root request
        Req root;
        box_clear(root.bounds); box_clear(root.cbounds);
        for (int i = 0; i < T.n; i++) { grow_box(root.bounds, T.bmin + 3 * i, T.bmax + 3 * i); grow_point(root.cbounds, T.cent + 3 * i); }
        root.start = 0; root.count = T.n; root.pre = 0; root.rank = 0; root.level = 0;
        height = 0;
        if (T.n < 2) {                                                                   // a single instance: the root is the leaf
            write_leaf(T, 0, root.bounds, T.prim[0]);
            T.result[0] = tlas_leaf_ref(T.prim[0]);
            curCount = 0;
        } else {
            cur[0] = root;
            T.result[0] = T.inner_base;
            curCount = 1;
        }
        nextCount = 0;
    }
    __syncthreads();
    while (curCount > 0) {
        const int m = curCount;
        for (int q = threadIdx.x; q < m; q += blockDim.x) {
            const Req req = cur[q];
            Req L, R;
            split_node(T, req, L, R);
            write_flat(T, req.pre, req.bounds, (float)(T.top_index + L.pre), (float)(T.top_index + R.pre), 0.f);   // ProcessTLASNodes :82-88
            const int lref = L.count < 2 ? tlas_leaf_ref(T.prim[L.start]) : T.inner_base + L.rank;
            const int rref = R.count < 2 ? tlas_leaf_ref(T.prim[R.start]) : T.inner_base + R.rank;
            float4* d = T.packed_nodes + (size_t)4 * (T.inner_base + req.rank);                                   // lf_repack.cpp build_nodes
            d[0] = make_float4(L.bounds.mn[0], L.bounds.mn[1], L.bounds.mn[2], L.bounds.mx[0]);
            d[1] = make_float4(L.bounds.mx[1], L.bounds.mx[2], R.bounds.mn[0], R.bounds.mn[1]);
            d[2] = make_float4(R.bounds.mn[2], R.bounds.mx[0], R.bounds.mx[1], R.bounds.mx[2]);
            d[3] = make_float4(__int_as_float(lref), __int_as_float(rref), 0.f, 0.f);
            atomicMax(&height, req.level + 1);
            if (L.count < 2) write_leaf(T, L.pre, L.bounds, T.prim[L.start]); else nxt[atomicAdd(&nextCount, 1)] = L;
            if (R.count < 2) write_leaf(T, R.pre, R.bounds, T.prim[R.start]); else nxt[atomicAdd(&nextCount, 1)] = R;
        }
        __syncthreads();
        if (threadIdx.x == 0) { curCount = nextCount; nextCount = 0; }
        Req* t = cur; cur = nxt; nxt = t;
        __syncthreads();
    }
    if (threadIdx.x == 0) T.result[1] = height;                                          // inner nodes on the longest root-to-leaf chain
}

size_t tlas_request_bytes() { return sizeof(Req); }

void launch_tlas_build(cudaStream_t stream, const TlasBuild& T, int sm_count) {
    int blocks = (T.n + 255) / 256;
    if (blocks > sm_count * 4) blocks = sm_count * 4;
    k_tlas_bounds<<<blocks, 256, 0, stream>>>(T);
    k_instance_records<<<blocks, 256, 0, stream>>>(T);
    k_tlas_build<<<1, 256, 0, stream>>>(T);
}

}  // namespace lf
