// lf_repack.cpp — see lf_repack.h / lf_types.h.  All arithmetic here is plain fp32 in the same operation
// order the shader uses for the values it would otherwise recompute per ray (edge vectors, inverse
// transforms, light planes), so the kernels see bit-identical operands.
#include "lf_repack.h"
#include "lf_matrix.h"

#include <cmath>
#include <cstring>
#include <functional>

namespace lf {

namespace {

inline float4 f4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
inline float as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }

struct NodeView {
    const float* nodes; int n;
    const float* mn(int i) const { return nodes + 9 * (size_t)i; }
    const float* mx(int i) const { return nodes + 9 * (size_t)i + 3; }
    int l(int i) const { return (int)nodes[9 * (size_t)i + 6]; }
    int r(int i) const { return (int)nodes[9 * (size_t)i + 7]; }
    int leaf(int i) const { return (int)nodes[9 * (size_t)i + 8]; }
};

}  // namespace

static bool build_nodes(const NodeView& nv, int top_index, int num_instances, int num_tri_refs, PackedScene& out, std::string& err) {
    const int n = nv.n;
    if (top_index < 0 || top_index >= n) { err = "top_bvh_index out of range"; return false; }
    std::vector<int> inner_id(n, -1);
    int ninner = 0;
    for (int i = 0; i < n; i++)
        if (nv.leaf(i) == 0) inner_id[i] = ninner++;
    bool bad = false;
    auto ref_of = [&](int i) -> int {
        int leaf = nv.leaf(i);
        if (leaf == 0) return inner_id[i];
        if (leaf > 0) {
            int first = nv.l(i), count = nv.r(i);
            if (first < 0 || count < 1 || count > kMaxLeafTris || first + count > num_tri_refs || first >= (1 << 24)) { bad = true; return kRefSentinel; }
            return make_blas_leaf_ref(first, count);
        }
        int inst = -leaf - 1;
        if (inst >= num_instances || inst >= (1 << 24)) { bad = true; return kRefSentinel; }
        return make_tlas_leaf_ref(inst);
    };
    out.nodes.assign((size_t)4 * ninner, f4(0, 0, 0, 0));
    for (int i = 0; i < n; i++) {
        if (nv.leaf(i) != 0) continue;
        int l = nv.l(i), r = nv.r(i);
        if (l < 0 || l >= n || r < 0 || r >= n) { err = "BVH child index out of range"; return false; }
        const float *lmn = nv.mn(l), *lmx = nv.mx(l), *rmn = nv.mn(r), *rmx = nv.mx(r);
        float4* d = &out.nodes[(size_t)4 * inner_id[i]];
        d[0] = f4(lmn[0], lmn[1], lmn[2], lmx[0]);
        d[1] = f4(lmx[1], lmx[2], rmn[0], rmn[1]);
        d[2] = f4(rmn[2], rmx[0], rmx[1], rmx[2]);
        d[3] = f4(as_float(ref_of(l)), as_float(ref_of(r)), 0.f, 0.f);
    }
    out.top_ref = ref_of(top_index);
    out.num_inner = ninner;

    // instance meta from the TLAS leaves; BLAS heights for the stack bound
    std::vector<int> height(n, -1);
    std::function<int(int)> h = [&](int i) -> int {   // inner nodes on the longest root-to-leaf chain below i
        if (nv.leaf(i) != 0) return 0;
        if (height[i] >= 0) return height[i];
        height[i] = 0;                                  // guards against cycles in corrupt input
        int v = 1 + std::max(h(nv.l(i)), h(nv.r(i)));
        height[i] = v;
        return v;
    };
    int maxBlas = 0;
    std::vector<char> seen(num_instances, 0);
    for (int i = top_index; i < n; i++) {
        int leaf = nv.leaf(i);
        if (leaf >= 0) continue;
        int inst = -leaf - 1;
        if (inst < 0 || inst >= num_instances) { err = "TLAS leaf names an instance that does not exist"; return false; }
        int root = nv.l(i);
        if (root < 0 || root >= top_index) { err = "TLAS leaf BLAS root out of range"; return false; }
        float4* ip = &out.inst[(size_t)kInstStride * inst];
        ip[3] = f4(as_float(ref_of(root)), as_float(nv.r(i)), 0.f, 0.f);
        seen[inst] = 1;
        maxBlas = std::max(maxBlas, h(root));
    }
    if (bad) { err = "BVH leaf exceeds a structural limit (triangle refs >= 2^24, > 64 triangles per leaf, or bad instance)"; return false; }
    int tlasH = h(top_index);
    out.max_blas_height = maxBlas;
    out.num_blas_inner = 0;
    for (int i = 0; i < top_index; i++)
        if (nv.leaf(i) == 0) out.num_blas_inner++;
    out.stack_depth = 2 + tlasH + maxBlas + 1;
    if (out.stack_depth > 64) { err = "BVH needs a traversal stack deeper than 64 entries (the reference's limit, closest_hit.glsl:70)"; return false; }
    return true;
}

static void build_instances(const float* transforms, int num_instances, PackedScene& out) {
    out.inst.assign((size_t)kInstStride * num_instances, f4(0, 0, 0, 0));
    for (int k = 0; k < num_instances; k++) {
        const float* t = transforms + 16 * (size_t)k;
        float m[4][4], inv[4][4];
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++) m[c][r] = t[4 * c + r];    // Mat4 data[c] is GLSL column c (closest_hit.glsl:152-157)
        inverse4(m, inv);
        float m3[3][3], i3[3][3];
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++) m3[c][r] = m[c][r];
        inverse3(m3, i3);
        float4* ip = &out.inst[(size_t)kInstStride * k];
        for (int r = 0; r < 3; r++) ip[r] = f4(inv[0][r], inv[1][r], inv[2][r], inv[3][r]);       // rows of inverse(M)
        for (int r = 0; r < 3; r++) ip[4 + r] = f4(m[0][r], m[1][r], m[2][r], m[3][r]);         // rows of M
        // normalMatrix = transpose(inverse(mat3(M))): its column j is row j of the inverse, so its ROW i is column i of the inverse
        for (int r = 0; r < 3; r++) ip[7 + r] = f4(i3[r][0], i3[r][1], i3[r][2], 0.f);
    }
}

bool repack_instances(const float* nodes, int num_nodes, int top_index, const float* transforms, int num_instances,
                      PackedScene& out, std::string& err) {
    build_instances(transforms, num_instances, out);
    NodeView nv{nodes, num_nodes};
    return build_nodes(nv, top_index, num_instances, (int)(out.tris.size() / kTriStride), out, err);
}

bool repack_scene(const LfSceneView& v, PackedScene& out, std::string& err) {
    if (!v.bvh_nodes || v.num_nodes <= 0 || !v.vert_indices || !v.vertices_uvx || !v.normals_uvy || !v.transforms || !v.materials ||
        v.num_instances <= 0 || v.num_materials <= 0) { err = "scene view has empty mandatory arrays"; return false; }
    // triangles in leaf order
    out.tris.assign((size_t)kTriStride * v.num_tri_refs, f4(0, 0, 0, 0));
    out.trinrm.resize((size_t)3 * v.num_tri_refs);
    out.tri_vx.resize(v.num_tri_refs);
    for (int i = 0; i < v.num_tri_refs; i++) {
        const int32_t* vi = v.vert_indices + 3 * (size_t)i;
        for (int k = 0; k < 3; k++)
            if (vi[k] < 0 || vi[k] >= v.num_vertices) { err = "vertex index out of range"; return false; }
        const float* a = v.vertices_uvx + 4 * (size_t)vi[0];
        const float* b = v.vertices_uvx + 4 * (size_t)vi[1];
        const float* c = v.vertices_uvx + 4 * (size_t)vi[2];
        out.tris[kTriStride * (size_t)i + 0] = f4(a[0], a[1], a[2], a[3]);
        out.tris[kTriStride * (size_t)i + 1] = f4(b[0] - a[0], b[1] - a[1], b[2] - a[2], b[3]);   // e0 = v1 - v0 (closest_hit.glsl:119)
        out.tris[kTriStride * (size_t)i + 2] = f4(c[0] - a[0], c[1] - a[1], c[2] - a[2], c[3]);   // e1 = v2 - v0 (:120)
        for (int k = 0; k < 3; k++) {
            const float* nn = v.normals_uvy + 4 * (size_t)vi[k];
            out.trinrm[3 * (size_t)i + k] = f4(nn[0], nn[1], nn[2], nn[3]);
        }
        out.tri_vx[i] = vi[0];
    }
    // lights: the 5 texels + derived values of closest_hit.glsl:29-35
    out.lights.assign((size_t)kLightStride * std::max(v.num_lights, 0), f4(0, 0, 0, 0));
    for (int i = 0; i < v.num_lights; i++) {
        const float* p = v.lights + 15 * (size_t)i;
        float u[3] = {p[6], p[7], p[8]}, w[3] = {p[9], p[10], p[11]};
        float cr[3] = {u[1] * w[2] - w[1] * u[2], u[2] * w[0] - w[2] * u[0], u[0] * w[1] - w[0] * u[1]};   // cross(u, v)
        float il = 1.0f / sqrtf((cr[0] * cr[0] + cr[1] * cr[1]) + cr[2] * cr[2]);
        float nrm[3] = {cr[0] * il, cr[1] * il, cr[2] * il};                                                // normalize
        float planeW = (nrm[0] * p[0] + nrm[1] * p[1]) + nrm[2] * p[2];                                      // dot(normal, position)
        float su = 1.0f / ((u[0] * u[0] + u[1] * u[1]) + u[2] * u[2]);                                       // u *= 1/dot(u,u)
        float sv = 1.0f / ((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2]);
        float4* d = &out.lights[(size_t)kLightStride * i];
        d[0] = f4(p[0], p[1], p[2], p[3]);
        d[1] = f4(p[4], p[5], u[0], u[1]);
        d[2] = f4(u[2], w[0], w[1], w[2]);
        d[3] = f4(p[12], p[13], p[14], planeW);
        d[4] = f4(nrm[0], nrm[1], nrm[2], u[0] * su);
        d[5] = f4(u[1] * su, u[2] * su, w[0] * sv, w[1] * sv);
        d[6] = f4(w[2] * sv, 0.f, 0.f, 0.f);
    }
    build_instances(v.transforms, v.num_instances, out);
    NodeView nv{v.bvh_nodes, v.num_nodes};
    return build_nodes(nv, v.top_bvh_index, v.num_instances, v.num_tri_refs, out, err);
}

}  // namespace lf
