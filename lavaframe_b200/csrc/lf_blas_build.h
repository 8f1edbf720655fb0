// lf_blas_build.h — the bottom-level BVH of one mesh built level by level from data-parallel steps (GPU BVH build).
//
// What the reference does on the host, once per mesh at scene load (Mesh::BuildBVH, LavaFrame/Mesh.cpp:93-111):
//   bvh = new RadeonRays::SplitBvh(2.0f, 64, 0, 0.001f, 0)          Mesh.h:18: traversal cost 2, 64 SAH bins, max_split_depth 0
//   Bvh::Build -> SplitBvh::BuildImpl -> BuildNode (recursive)      thirdparty/RadeonRays/bvh.cpp:41-50, split_bvh.cpp:11-168
//   FindObjectSahSplit                                              split_bvh.cpp:170-289
//   BvhTranslator::ProcessBLASNodes                                 bvh_translator.cpp:35-60 (pre-order flattening, left child = own index + 1)
// With max_split_depth = 0 the spatial-split branch (split_bvh.cpp:75-85) is never entered (`req.level < 0` is false), so the builder is a
// binned-SAH object-split builder: a node with fewer than 4 primitives is a leaf; otherwise 64 bins per axis over the centroid box, the
// cheapest of 3 x 63 candidate planes (strict `<`, axis 0 first), an in-place two-pointer partition whose direction alternates with
// (numprims + startidx) & 1, the range halved when a side stays empty (the child boxes then KEEP what the partition loop already grew), the
// right child built before the left one (which fixes the order of the leaf indices: a leaf over array positions [s, s + k) of N primitive
// references gets startidx = N - (s + k)).
//
// This file restates that builder as level-synchronous data-parallel steps whose results do not depend on the order in which the items of
// a step run, and which produce the reference's tree NODE FOR NODE:
//   * min / max / counts are exact and order-independent - the sign of a zero aside: std::min / std::max keep the FIRST of +0 / -0 they meet, so a
//     box plane that is exactly zero carries the sign of the first zero in the reference's growth order.  That order is restated as a key per
//     element (array index for the root, destination position for a left child, (block, position) for a right child, see acc_zero) and the
//     smallest key among a plane's zeros decides the sign; no decision of the builder looks at it, only the node boxes it writes out;
//   * the SAH sweep of a node is the reference's sequential fp32 code, run literally by one work item per node;
//   * the two-pointer partition is determined by the L / R classes alone: the elements in place stay, the k-th misplaced element from the left
//     (an R in the first nL positions) is exchanged with the misplaced L that has k L's after it; positions follow from a prefix sum.
// The SAME text runs under two executors: `lf_blas.cu` launches every step as a CUDA kernel (the product path, lfcuda_build_blas), and
// tests/hostcheck compiles it for the host, where the steps run as loops - forwards, backwards and shuffled - and are compared with the
// reference builder's output on every scene the tests hold (test infrastructure: it proves the algorithm and its independence of order; the
// GPU tests prove the CUDA execution).
#pragma once

#include <cfloat>
#include <cstdint>
#include <vector>

#include <vector_types.h>
#include <vector_functions.h>

#if defined(__CUDACC__)
#define BB_HD __host__ __device__ __forceinline__
#else
#define BB_HD inline
#endif

namespace lf {
namespace blas {

constexpr int kMaxBins = 64;
constexpr int kBinFields = 7;        // count, pmin.xyz, pmax.xyz
constexpr int kScanChunk = 32;       // elements one work item of the prefix sum handles sequentially

struct Box { float mn[3], mx[3]; };
BB_HD float smin(float a, float b) { return (b < a) ? b : a; }     // std::min(a, b)
BB_HD float smax(float a, float b) { return (a < b) ? b : a; }     // std::max(a, b)
BB_HD void box_clear(Box& b) { for (int k = 0; k < 3; k++) { b.mn[k] = FLT_MAX; b.mx[k] = -FLT_MAX; } }                 // bbox(), bbox.h:47-55
BB_HD void box_grow(Box& b, const Box& o) { for (int k = 0; k < 3; k++) { b.mn[k] = smin(b.mn[k], o.mn[k]); b.mx[k] = smax(b.mx[k], o.mx[k]); } }
BB_HD float box_area(const Box& b) {                                                                                     // bbox::surface_area, bbox.cpp:31-35
    float ex = b.mx[0] - b.mn[0], ey = b.mx[1] - b.mn[1], ez = b.mx[2] - b.mn[2];
    return 2.f * ((ex * ey + ex * ez) + ey * ez);
}
BB_HD int box_maxdim(const Box& b) {                                                                                     // bbox::maxdim, bbox.h:71-83
    float ex = b.mx[0] - b.mn[0], ey = b.mx[1] - b.mn[1], ez = b.mx[2] - b.mn[2];
    if (ex >= ey && ex >= ez) return 0;
    if (ey >= ex && ey >= ez) return 1;
    if (ez >= ex && ez >= ey) return 2;
    return 0;
}
BB_HD float center_of(float mn, float mx) { return (mx + mn) * 0.5f; }                                                   // bbox::center, bbox.cpp:28
BB_HD bool is_nan(float v) { return v != v; }
BB_HD unsigned f2i_bits(float f) { union { float f; unsigned u; } c; c.f = f; return c.u; }

// ---- order-independent accumulation.  Device: atomics (float min / max through the integer order of IEEE floats); host executor: plain.
BB_HD void acc_min(float* a, float v) {
#ifdef __CUDA_ARCH__
    if (v == 0.f) v = 0.f;                                   // -0 -> +0 (see negative_zero above)
    if (!(v < __ldcg(a))) return;                            // *a only ever decreases: a stale read can only send a redundant atomic
    if (v >= 0.f) atomicMin((int*)a, __float_as_int(v)); else atomicMax((unsigned*)a, __float_as_uint(v));
#else
    *a = smin(*a, v);
#endif
}
BB_HD void acc_max(float* a, float v) {
#ifdef __CUDA_ARCH__
    if (v == 0.f) v = 0.f;
    if (!(__ldcg(a) < v)) return;
    if (v >= 0.f) atomicMax((int*)a, __float_as_int(v)); else atomicMin((unsigned*)a, __float_as_uint(v));
#else
    *a = smax(*a, v);
#endif
}
BB_HD void acc_add(int* a, int v) {
#ifdef __CUDA_ARCH__
    atomicAdd(a, v);
#else
    *a += v;
#endif
}
BB_HD void acc_or(int* a, int v) {
#ifdef __CUDA_ARCH__
    if (v) atomicOr(a, v);
#else
    *a |= v;
#endif
}
#ifdef __CUDA_ARCH__
BB_HD int ordered_int(float f) { if (f == 0.f) f = 0.f; const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }    // monotonic float -> int
BB_HD float ordered_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
#endif
BB_HD void acc_box(Box* dst, const Box& b) {
#ifdef __CUDA_ARCH__
    // the lanes of a warp that grow the same box (near the root: all of them) reduce among themselves first: one atomic per group and plane
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, (unsigned long long)dst);
    if (__popc(peers) >= 4) {
        const bool leader = (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1);
        for (int k = 0; k < 3; k++) {
            const int lo = __reduce_min_sync(peers, ordered_int(b.mn[k])), hi = __reduce_max_sync(peers, ordered_int(b.mx[k]));
            if (leader) { acc_min(&dst->mn[k], ordered_float(lo)); acc_max(&dst->mx[k], ordered_float(hi)); }
        }
        return;
    }
#endif
    for (int k = 0; k < 3; k++) { acc_min(&dst->mn[k], b.mn[k]); acc_max(&dst->mx[k], b.mx[k]); }
}

BB_HD void acc_min64(unsigned long long* a, unsigned long long v) {
#ifdef __CUDA_ARCH__
    atomicMin(a, v);
#else
    if (v < *a) *a = v;
#endif
}
// ---- one node of the level being split (Bvh::SplitRequest + the decisions taken for it)
struct LevelNode {
    int start, count;        // range of primitive references
    int gid;                 // index in the node store (allocation order: level by level)
    int rank;                // position among the inner nodes of its level (-1: leaf)
    int axis, part;          // split axis; 1 = the partition loop runs (centroid extent along the axis > 0, split_bvh.cpp:108)
    float border;
    int near2far;            // (numprims + startidx) & 1: which class goes left (split_bvh.cpp:100-105)
    int nL, split;           // elements of the left class; first position of the right child
    int allL, allR;          // the partition ran and left a side empty: that side's grown boxes are kept (split_bvh.cpp:145-160)
    int moved;               // 1 = elements change places
    int sidx[3]; float sahv[3];  // best plane of every axis and its cost (StepSahAxis)
    unsigned long long zkey[6];  // per plane of b: (growth order << 1 | sign bit) of the first zero-valued contribution (acc_zero)
    Box b, cb;               // bounds, centroid bounds
};
struct GNode {               // node store
    Box b;
    int left, right;         // inner: gids of the children; leaf: startidx into the packed indices, numprims
    int leaf;
    int size, pre;           // nodes of the subtree; pre-order position
};

// The sign of zero planes (see the header).  `order` = the element's place in the sequence in which the reference grows this box.
BB_HD void zkey_clear(LevelNode& nd) { for (int k = 0; k < 6; k++) nd.zkey[k] = ~0ull; }
BB_HD void acc_zero(LevelNode* nd, const Box& b, unsigned long long order) {
    for (int k = 0; k < 3; k++) {
        if (b.mn[k] == 0.f) acc_min64(&nd->zkey[k], (order << 1) | (unsigned long long)(f2i_bits(b.mn[k]) >> 31));
        if (b.mx[k] == 0.f) acc_min64(&nd->zkey[3 + k], (order << 1) | (unsigned long long)(f2i_bits(b.mx[k]) >> 31));
    }
}
BB_HD void zero_signs(Box& b, const LevelNode& nd) {
    for (int k = 0; k < 3; k++) {
        if (b.mn[k] == 0.f) b.mn[k] = (nd.zkey[k] & 1ull) ? -0.f : 0.f;
        if (b.mx[k] == 0.f) b.mx[k] = (nd.zkey[3 + k] & 1ull) ? -0.f : 0.f;
    }
}

struct State {
    int n, nbins;
    float tc;                           // traversal cost (Mesh.h:18: 2.0f)
    const float* in_bounds;             // n x 6: pmin.xyz, pmax.xyz of every primitive (Mesh.cpp:96-108)
    float4* lo[2]; float4* hi[2];       // primitive references, ping-pong: (pmin.xyz, index bits), (pmax.xyz, -)
    int* node_of[2];                    // level-local node of every array position, -1 = already in a leaf
    int* flag;                          // n + 1: element belongs to the left class
    int* scan;                          // n + 1: exclusive prefix sum of flag
    int* chunk;                         // partial sums of the prefix sum
    int* pairL; int* pairR;             // position of the k-th misplaced element from the left / with k left-class elements after it
    LevelNode* lev[2];                  // current and next level
    int* nflag; int* nscan;             // per level node: is inner; exclusive prefix sum
    GNode* g;                           // 2 n - 1
    float* bins;                        // bin_cap x 3 x kBinFields x kMaxBins (counts stored as int bits)
    int bin_cap;
    int* misc;                          // [0] a primitive bound is -0.0 (reported, nothing depends on it), [1] inner nodes of the current level
    float* out_nodes;                   // num_nodes x 9 words: pmin, pmax, then three INT32: (left, right, 0) / (startidx, numprims, 1)
    int* out_indices;                   // n: packed primitive indices (Bvh::GetIndices)
};
// bins[axis][field][bin][slot]: the slot (= inner node of the level) is the fastest index, so that the SAH sweeps of neighbouring nodes, one
// work item each, read neighbouring words
BB_HD float* bin_at(const State& S, int slot, int axis, int field, int bin) { return S.bins + ((((size_t)axis * kBinFields + field) * kMaxBins + bin) * (size_t)S.bin_cap + slot); }
BB_HD Box ref_box(const float4& lo, const float4& hi) { Box b; b.mn[0] = lo.x; b.mn[1] = lo.y; b.mn[2] = lo.z; b.mx[0] = hi.x; b.mx[1] = hi.y; b.mx[2] = hi.z; return b; }
BB_HD int f2i(float f) { union { float f; int i; } u; u.f = f; return u.i; }
BB_HD float i2f(int i) { union { float f; int i; } u; u.i = i; return u.f; }

// =============================================================================================== steps (one work item = one call)
// SplitBvh::BuildImpl (split_bvh.cpp:11-34) + Bvh::Build (bvh.cpp:41-50): primitive references, world bounds, centroid bounds
struct StepInit {
    State S;
    BB_HD void operator()(int i) const {
        const float* p = S.in_bounds + 6 * (size_t)i;
        Box b; for (int k = 0; k < 3; k++) { b.mn[k] = p[k]; b.mx[k] = p[3 + k]; }
        int negzero = 0;
        for (int k = 0; k < 6; k++) negzero |= (p[k] == 0.f && f2i(p[k]) < 0) ? 1 : 0;
        acc_or(&S.misc[0], negzero);
        S.lo[0][i] = make_float4(b.mn[0], b.mn[1], b.mn[2], i2f(i));
        S.hi[0][i] = make_float4(b.mx[0], b.mx[1], b.mx[2], 0.f);
        S.node_of[0][i] = 0;
        Box c; for (int k = 0; k < 3; k++) c.mn[k] = c.mx[k] = center_of(b.mn[k], b.mx[k]);
        acc_box(&S.lev[0][0].b, b);
        acc_box(&S.lev[0][0].cb, c);
        acc_zero(&S.lev[0][0], b, (unsigned long long)i);                     // Bvh::Build grows m_bounds in index order (bvh.cpp:43-47)
    }
};
struct StepRoot {        // before StepInit
    State S;
    BB_HD void operator()(int) const {
        LevelNode& r = S.lev[0][0];
        r.start = 0; r.count = S.n; r.gid = 0; r.rank = -1;
        box_clear(r.b); box_clear(r.cb); zkey_clear(r);
        S.misc[0] = 0;
    }
};

// Leaf or inner (split_bvh.cpp:45-56); leaves hand their primitive indices out at once
struct StepClassify {
    State S; int cur, level;
    BB_HD void operator()(int j) const {
        LevelNode& nd = S.lev[cur][j];
        GNode& g = S.g[nd.gid];
        g.b = nd.b;
        zero_signs(g.b, nd);
        const bool leaf = nd.count < 4;
        S.nflag[j] = leaf ? 0 : 1;
        g.leaf = leaf ? 1 : 0;
        if (leaf) {
            g.left = S.n - (nd.start + nd.count); g.right = nd.count;       // the right child is built first: leaves are numbered from the array's end
            g.size = 1;
            for (int k = 0; k < nd.count; k++) {
                const int p = nd.start + k;
                S.out_indices[g.left + k] = f2i(S.lo[cur][p].w);
                S.node_of[cur][p] = -1;
            }
        }
    }
};
struct StepRank {        // after the prefix sum over nflag; misc[1] = inner nodes of the level (the one number the host loop reads back)
    State S; int cur, ncur;
    BB_HD void operator()(int j) const {
        S.lev[cur][j].rank = S.nflag[j] ? S.nscan[j] : -1;
        if (j == ncur - 1) S.misc[1] = S.nscan[j] + S.nflag[j];
    }
};
struct StepBinsClear {
    State S; int nslots;
    BB_HD void operator()(int i) const {
        const int slot = i % nslots, r = i / nslots;          // r = (axis * kBinFields + field) * kMaxBins + bin
        const int field = (r / kMaxBins) % kBinFields;
        S.bins[(size_t)r * S.bin_cap + slot] = field == 0 ? i2f(0) : (field <= 3 ? FLT_MAX : -FLT_MAX);
    }
};
// Histogram of the primitive references over the centroid box, per axis (split_bvh.cpp:214-232)
struct StepBin {
    State S; int cur, rank_lo, rank_hi;
    BB_HD void operator()(int p) const {
        const int j = S.node_of[cur][p];
        if (j < 0) return;
        const LevelNode& nd = S.lev[cur][j];
        if (nd.rank < rank_lo || nd.rank >= rank_hi) return;
        const float4 lo = S.lo[cur][p], hi = S.hi[cur][p];
        const Box b = ref_box(lo, hi);
        for (int axis = 0; axis < 3; axis++) {
            const float rootminc = nd.cb.mn[axis];
            const float rng = nd.cb.mx[axis] - nd.cb.mn[axis];
            if (rng == 0.f) continue;
            const float inv = 1.f / rng;
            const float c = center_of(b.mn[axis], b.mx[axis]);
            const float x = (float)S.nbins * ((c - rootminc) * inv);
            const float lim = (float)(S.nbins - 1);
            const int bin = (int)((lim < x) ? lim : x);                       // (int)std::min<float>(x, lim)
            const int slot = nd.rank - rank_lo;
            acc_add((int*)bin_at(S, slot, axis, 0, bin), 1);
            for (int k = 0; k < 3; k++) { acc_min(bin_at(S, slot, axis, 1 + k, bin), b.mn[k]); acc_max(bin_at(S, slot, axis, 4 + k, bin), b.mx[k]); }
        }
    }
};
// FindObjectSahSplit (split_bvh.cpp:170-289), one work item per (inner node, axis): the cheapest of the axis' 63 planes, first one on ties.
// The reference carries ONE running minimum through the three axes (strict `<`), i.e. it keeps the first minimum in (axis, plane) order;
// the first minimum of every axis, combined in axis order with the same strict `<` (StepSahPick), is the same plane.
struct StepSahAxis {
    State S; int cur, rank_lo, rank_hi;
    BB_HD void operator()(int i) const {
        const int j = i / 3, a = i - 3 * j;
        LevelNode& nd = S.lev[cur][j];
        if (nd.rank < rank_lo || nd.rank >= rank_hi) return;
        const int slot = nd.rank - rank_lo, nb = S.nbins;
        nd.sidx[a] = -1; nd.sahv[a] = FLT_MAX;
        const float cext[3] = {nd.cb.mx[0] - nd.cb.mn[0], nd.cb.mx[1] - nd.cb.mn[1], nd.cb.mx[2] - nd.cb.mn[2]};
        if ((cext[0] * cext[0] + cext[1] * cext[1]) + cext[2] * cext[2] == 0.f) return;      // split_bvh.cpp:186-190
        if (cext[a] == 0.f) return;                                                          // :212
        const float invarea = 1.f / box_area(nd.b);
        float rsa[kMaxBins];                                                                 // surface areas of rightbounds[i]
        Box rb; box_clear(rb);
#pragma unroll 4
        for (int k = nb - 1; k > 0; --k) {
            Box bb; for (int c = 0; c < 3; c++) { bb.mn[c] = *bin_at(S, slot, a, 1 + c, k); bb.mx[c] = *bin_at(S, slot, a, 4 + c, k); }
            box_grow(rb, bb);
            rsa[k - 1] = box_area(rb);
        }
        Box lb; box_clear(lb);
        int leftcount = 0, rightcount = nd.count, splitidx = -1;
        float sah = FLT_MAX;
#pragma unroll 4
        for (int k = 0; k < nb - 1; ++k) {
            Box bb; for (int c = 0; c < 3; c++) { bb.mn[c] = *bin_at(S, slot, a, 1 + c, k); bb.mx[c] = *bin_at(S, slot, a, 4 + c, k); }
            box_grow(lb, bb);
            const int cnt = f2i(*bin_at(S, slot, a, 0, k));
            leftcount += cnt;
            rightcount -= cnt;
            const float sahtmp = S.tc + ((float)leftcount * box_area(lb) + (float)rightcount * rsa[k]) * invarea;
            if (sahtmp < sah) { splitidx = k; sah = sahtmp; }
        }
        nd.sidx[a] = splitidx; nd.sahv[a] = sah;
    }
};
// ... and the choice of plane in BuildNode (split_bvh.cpp:59-98,108), one work item per inner node
struct StepSahPick {
    State S; int cur, rank_lo, rank_hi;
    BB_HD void operator()(int j) const {
        LevelNode& nd = S.lev[cur][j];
        if (nd.rank < rank_lo || nd.rank >= rank_hi) return;
        int axis = box_maxdim(nd.cb);
        float border = center_of(nd.cb.mn[axis], nd.cb.mx[axis]);
        int splitidx = -1, dim = 0;
        float sah = FLT_MAX;
        for (int a = 0; a < 3; a++)
            if (nd.sidx[a] != -1 && nd.sahv[a] < sah) { dim = a; splitidx = nd.sidx[a]; sah = nd.sahv[a]; }
        float split = i2f(0x7fc00000);
        if (splitidx != -1) split = nd.cb.mn[dim] + (float)(splitidx + 1) * ((nd.cb.mx[dim] - nd.cb.mn[dim]) / (float)S.nbins);
        // max_split_depth = 0: the object split, or the centre of the centroid box when there is none
        if (!is_nan(split)) { border = split; axis = dim; }
        nd.axis = axis; nd.border = border;
        nd.near2far = (nd.count + nd.start) & 1;
        nd.part = (nd.cb.mx[axis] - nd.cb.mn[axis]) > 0.f ? 1 : 0;
    }
};
// Class of every element: 1 = it ends in the left child (cmp1, split_bvh.cpp:100-105)
struct StepFlag {
    State S; int cur;
    BB_HD void operator()(int p) const {
        if (p == S.n) { S.flag[p] = 0; return; }
        const int j = S.node_of[cur][p];
        int L = 0;
        if (j >= 0) {
            const LevelNode& nd = S.lev[cur][j];
            if (nd.part) {
                const float4 lo = S.lo[cur][p], hi = S.hi[cur][p];
                const float mn = nd.axis == 0 ? lo.x : (nd.axis == 1 ? lo.y : lo.z), mx = nd.axis == 0 ? hi.x : (nd.axis == 1 ? hi.y : hi.z);
                const float c = center_of(mn, mx);
                L = nd.near2far ? (c < nd.border) : (c >= nd.border);
            }
        }
        S.flag[p] = L;
    }
};
// Split position, halving fallback, the two requests (split_bvh.cpp:140-163)
struct StepSplit {
    State S; int cur, next_gid_base, level;
    BB_HD void operator()(int j) const {
        LevelNode& nd = S.lev[cur][j];
        if (nd.rank < 0) return;
        const int nL = nd.part ? S.scan[nd.start + nd.count] - S.scan[nd.start] : 0;
        nd.nL = nL;
        nd.allL = nd.part && nL == nd.count;
        nd.allR = nd.part && nL == 0;
        nd.moved = nd.part && !nd.allL && !nd.allR;
        nd.split = nd.moved ? nd.start + nL : nd.start + (nd.count >> 1);
        GNode& g = S.g[nd.gid];
        g.left = next_gid_base + 2 * nd.rank; g.right = g.left + 1;
        LevelNode& l = S.lev[cur ^ 1][2 * nd.rank];
        LevelNode& r = S.lev[cur ^ 1][2 * nd.rank + 1];
        l.start = nd.start; l.count = nd.split - nd.start; l.gid = g.left; l.rank = -1;
        r.start = nd.split; r.count = nd.count - l.count; r.gid = g.right; r.rank = -1;
        box_clear(l.b); box_clear(l.cb); box_clear(r.b); box_clear(r.cb); zkey_clear(l); zkey_clear(r);
    }
};
// Where the misplaced elements are (see the header: the two-pointer partition exchanges them pairwise, split_bvh.cpp:113-137)
struct StepPair {
    State S; int cur;
    BB_HD void operator()(int p) const {
        const int j = S.node_of[cur][p];
        if (j < 0) return;
        const LevelNode& nd = S.lev[cur][j];
        if (!nd.moved) return;
        const int L = S.flag[p];
        if (p < nd.split && !L) S.pairL[nd.start + ((p - nd.start) - (S.scan[p] - S.scan[nd.start]))] = p;      // k = right-class elements before p
        if (p >= nd.split && L) S.pairR[nd.start + (nd.nL - (S.scan[p + 1] - S.scan[nd.start]))] = p;          // k = left-class elements after p
    }
};
// Every element to its place in the next level's array, and into the boxes of its child (split_bvh.cpp:113-160)
struct StepMove {
    State S; int cur;
    BB_HD void operator()(int p) const {
        const int j = S.node_of[cur][p];
        if (j < 0) { S.node_of[cur ^ 1][p] = -1; return; }
        const LevelNode& nd = S.lev[cur][j];
        int q = p;
        if (nd.moved) {
            const int L = S.flag[p];
            if (p < nd.split && !L) q = S.pairR[nd.start + ((p - nd.start) - (S.scan[p] - S.scan[nd.start]))];
            else if (p >= nd.split && L) q = S.pairL[nd.start + (nd.nL - (S.scan[p + 1] - S.scan[nd.start]))];
        }
        const float4 lo = S.lo[cur][p], hi = S.hi[cur][p];
        S.lo[cur ^ 1][q] = lo; S.hi[cur ^ 1][q] = hi;
        const int right = q >= nd.split ? 1 : 0;
        S.node_of[cur ^ 1][q] = 2 * nd.rank + right;
        const Box b = ref_box(lo, hi);
        Box c; for (int k = 0; k < 3; k++) c.mn[k] = c.mx[k] = center_of(b.mn[k], b.mx[k]);
        LevelNode* ch = &S.lev[cur ^ 1][2 * nd.rank];
        // the halving fallback re-grows the boxes by position on top of what the partition loop put into them: everything, on the side all went to
        if (!right || nd.allL) { acc_box(&ch[0].b, b); acc_box(&ch[0].cb, c); }
        if (right || nd.allR) { acc_box(&ch[1].b, b); acc_box(&ch[1].cb, c); }
        // Growth order of the two boxes (only the sign of a zero plane depends on it).  Left: the first pointer passes the elements in place in
        // ascending position and a misplaced left-class element joins when it is swapped in, i.e. ascending DESTINATION.  Right, when the loop
        // runs to a split (or finds no left-class element at all): the element the first pointer stops at, then the elements in place above
        // it in DESCENDING position down to the next misplaced one, and so on: blocks numbered by the left-class elements above, each block's
        // lowest destination first.  Fresh boxes of the halving fallback (split_bvh.cpp:145-160) grow in ascending position.
        const bool hasZero = b.mn[0] == 0.f || b.mn[1] == 0.f || b.mn[2] == 0.f || b.mx[0] == 0.f || b.mx[1] == 0.f || b.mx[2] == 0.f;
        if (hasZero) {
            if (!right || nd.allL) acc_zero(&ch[0], b, (unsigned long long)q);
            if (right || nd.allR) {
                unsigned long long order = (unsigned long long)q;
                if (nd.moved || nd.allR) {
                    const int loopsplit = nd.moved ? nd.split : nd.start;
                    const unsigned long long block = (unsigned long long)(nd.nL - (S.scan[q + 1] - S.scan[nd.start]));     // left-class elements above q
                    const bool first = S.flag[q] != 0 || q == loopsplit;
                    order = (block << 31) | (first ? 0ull : (unsigned long long)(0x7fffffff - q));
                }
                acc_zero(&ch[1], b, order);
            }
        }
    }
};
// ---- flattening (BvhTranslator::ProcessBLASNodes, bvh_translator.cpp:35-60): subtree sizes bottom-up, pre-order positions top-down
struct StepSize {
    State S; int base;
    BB_HD void operator()(int i) const { GNode& g = S.g[base + i]; if (!g.leaf) g.size = 1 + S.g[g.left].size + S.g[g.right].size; }
};
struct StepPre {
    State S; int base;
    BB_HD void operator()(int i) const {
        GNode& g = S.g[base + i];
        if (base + i == 0) g.pre = 0;
        if (!g.leaf) { S.g[g.left].pre = g.pre + 1; S.g[g.right].pre = g.pre + 1 + S.g[g.left].size; }
    }
};
struct StepEmit {
    State S;
    BB_HD void operator()(int i) const {
        const GNode& g = S.g[i];
        float* o = S.out_nodes + 9 * (size_t)g.pre;
        for (int k = 0; k < 3; k++) { o[k] = g.b.mn[k]; o[3 + k] = g.b.mx[k]; }
        int* oi = (int*)(o + 6);
        if (g.leaf) { oi[0] = g.left; oi[1] = g.right; oi[2] = 1; }
        else { oi[0] = S.g[g.left].pre; oi[1] = S.g[g.right].pre; oi[2] = 0; }
    }
};
// ---- exclusive prefix sum of n ints: sums of chunks of 32, the sums scanned the same way (recursively, in place), chunk-local scans.
// Integer sums: exact whatever the grouping.  `in == out` is allowed (a work item reads an element before it overwrites it).
struct StepScanSum {
    const int* in; int* sums; int n;
    BB_HD void operator()(int c) const {
        int s = 0; const int e = (c + 1) * kScanChunk < n ? (c + 1) * kScanChunk : n;
        for (int i = c * kScanChunk; i < e; i++) s += in[i];
        sums[c] = s;
    }
};
struct StepScanWrite {
    const int* in; int* out; const int* sums; int n;      // sums: exclusive prefix of the chunk sums (nullptr: a single chunk)
    BB_HD void operator()(int c) const {
        int s = sums ? sums[c] : 0; const int e = (c + 1) * kScanChunk < n ? (c + 1) * kScanChunk : n;
        for (int i = c * kScanChunk; i < e; i++) { const int v = in[i]; out[i] = s; s += v; }
    }
};
template <class Exec>
void exclusive_scan(Exec& ex, const int* in, int* out, int* scratch, int n) {
    const int nchunks = (n + kScanChunk - 1) / kScanChunk;
    if (nchunks <= 1) { ex.run(1, StepScanWrite{in, out, nullptr, n}); return; }
    ex.run(nchunks, StepScanSum{in, scratch, n});
    exclusive_scan(ex, scratch, scratch, scratch + nchunks, nchunks);
    ex.run(nchunks, StepScanWrite{in, out, scratch, n});
}

struct Result { int num_nodes, height, negative_zero, levels; };

// Storage a build of n primitives needs (bytes per array), for the executors' allocators
inline int level_capacity(int n) { return n / 2 + 2; }         // an inner node holds >= 4 references, so a level has <= 2 * (n / 4) nodes
inline int chunk_capacity(int n) { return (n + 1) / (kScanChunk - 1) + 64; }     // all levels of the recursive prefix sum

// The level loop.  Exec: run(count, step) executes step(0..count-1) in any order / in parallel, with a barrier between runs;
// read_int(ptr) reads one int the steps wrote.
template <class Exec>
Result build(Exec& ex, State& S) {
    Result R{};
    ex.run(1, StepRoot{S});
    ex.run(S.n, StepInit{S});
    int cur = 0, ncur = 1, level = 0, gid_base = 0;
    std::vector<int> level_base, level_count;
    for (;;) {
        level_base.push_back(gid_base); level_count.push_back(ncur);
        ex.run(ncur, StepClassify{S, cur, level});
        exclusive_scan(ex, S.nflag, S.nscan, S.chunk, ncur);
        ex.run(ncur, StepRank{S, cur, ncur});
        const int ninner = ex.read_int(S.misc + 1);
        if (ninner == 0) break;
        for (int lo = 0; lo < ninner; lo += S.bin_cap) {
            const int hi = lo + S.bin_cap < ninner ? lo + S.bin_cap : ninner;
            ex.run((hi - lo) * 3 * kBinFields * kMaxBins, StepBinsClear{S, hi - lo});
            ex.run(S.n, StepBin{S, cur, lo, hi});
            ex.run(3 * ncur, StepSahAxis{S, cur, lo, hi});
            ex.run(ncur, StepSahPick{S, cur, lo, hi});
        }
        ex.run(S.n + 1, StepFlag{S, cur});
        exclusive_scan(ex, S.flag, S.scan, S.chunk, S.n + 1);
        ex.run(ncur, StepSplit{S, cur, gid_base + ncur, level});
        ex.run(S.n, StepPair{S, cur});
        ex.run(S.n, StepMove{S, cur});
        gid_base += ncur; ncur = 2 * ninner; cur ^= 1; level++;
    }
    R.num_nodes = gid_base + ncur;
    R.height = level;                                                    // m_height = max(req.level)
    R.levels = level + 1;
    for (int l = level; l >= 0; l--) ex.run(level_count[l], StepSize{S, level_base[l]});
    for (int l = 0; l <= level; l++) ex.run(level_count[l], StepPre{S, level_base[l]});
    ex.run(R.num_nodes, StepEmit{S});
    R.negative_zero = ex.read_int(S.misc);
    return R;
}

}  // namespace blas
}  // namespace lf
