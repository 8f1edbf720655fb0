// lf_hostcheck — TEST INFRASTRUCTURE, never shipped and never on a product path.
//
// Compiles the TEXT of the CUDA device functions (lavaframe_b200/csrc/lf_device.cuh, lf_math.cuh, lf_shade.cuh) for the
// host with g++ (-ffp-contract=off, the host's counterpart of nvcc -fmad=false) and runs the megakernel's per-sample loop
// (k_megakernel, lf_kernels.cu) on the CPU over the same re-packed arrays the GPU reads (lf_repack.cpp).  The CPU test
// suite compares its image with the oracle's bit for bit, so a statement that differs between the kernels' source and
// the oracle is caught here, without a GPU; what only the GPU can show (the warp-cooperative walk of k_trace, texture
// objects, the device's rounding of / and sqrt) stays with the `-m gpu` parity tests.
//
// The device intrinsics the headers use are supplied below with their documented meaning.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#define LF_HOST_CHECK 1
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#define LF_NODE_LDG128 1

static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint2float_rn(unsigned u) { return (float)u; }                 // round to nearest even (default mode)
static inline int __float2int_rn(float f) { return (int)nearbyintf(f); }             // round to nearest even (default mode)
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }

// point-sampled, unnormalised, clamped texture fetches on host arrays (what lfcuda.cpp configures the texture objects for)
struct HostTex { const void* data; int w, h, layers; };
template <class T> static inline T tex2D(cudaTextureObject_t obj, float x, float y) {
    const HostTex* t = reinterpret_cast<const HostTex*>(obj);
    int xi = min(max((int)floorf(x), 0), t->w - 1), yi = min(max((int)floorf(y), 0), t->h - 1);
    return static_cast<const T*>(t->data)[(size_t)yi * t->w + xi];
}
template <class T> static inline T tex2DLayered(cudaTextureObject_t obj, float x, float y, int layer) {
    const HostTex* t = reinterpret_cast<const HostTex*>(obj);
    int xi = min(max((int)floorf(x), 0), t->w - 1), yi = min(max((int)floorf(y), 0), t->h - 1);
    return static_cast<const T*>(t->data)[((size_t)layer * t->h + yi) * t->w + xi];
}

#include "lf_shade.cuh"
#include "lf_post.cuh"
#include "lf_repack.h"
#include "scenepack.h"

using namespace lf;

namespace {

// the host's copy of lfcuda.cpp glsl_tan (tan as llvmpipe evaluates it), through the headers' own sincos
float host_tan(float x) { return lf_tan(x); }

struct Check {
    lfpack::ScenePack pack;
    PackedScene packed;
    std::vector<float4> hdr_rgba;
    HostTex tex8{}, texf{};
    DevScene S{};
    LfParams P{};
    LfCamera C{};
};

void fill_params(const Check& c, DevParams& D, int first_frame, int nframes, int stride, int tile_x, int tile_y) {   // lfcuda.cpp fill_dev_params
    const LfParams& P = c.P;
    const LfCamera& C = c.C;
    std::memset(&D, 0, sizeof D);
    D.width = P.width; D.height = P.height; D.tile_w = P.tile_width; D.tile_h = P.tile_height;
    D.max_depth = P.max_depth; D.enable_rr = P.enable_rr; D.rr_depth = P.rr_depth;
    D.use_envmap = (P.use_envmap && c.S.hdr_w > 0) ? 1 : 0;
    D.use_constant_bg = P.use_constant_bg;
    for (int k = 0; k < 3; k++) {
        D.bg[k] = P.bg_color[k];
        D.cam_pos[k] = C.position[k]; D.cam_right[k] = C.right[k]; D.cam_up[k] = C.up[k]; D.cam_fwd[k] = C.forward[k];
    }
    D.hdr_multiplier = P.hdr_multiplier;
    D.hdr_resolution = (float)(c.S.hdr_w * c.S.hdr_h);
    D.inv_tiles_x = 1.0f / ((float)P.width / P.tile_width);
    D.inv_tiles_y = 1.0f / ((float)P.height / P.tile_height);
    D.cam_scale = host_tan(C.fov * 0.5f);
    D.focal_dist = C.focal_dist; D.aperture = C.aperture;
    D.tile_x = tile_x; D.tile_y = tile_y;
    D.first_frame = first_frame; D.frame_stride = stride; D.num_frames = nframes;
}

// k_megakernel's loop body for one pixel-sample
template <bool CULL>
f3 sample(const DevScene& S, const DevParams& P, int lx, int ly, int frame, int* stk) {
    PathRegs ps;
    if (P.preview) ps.ray = preview_ray(P, lx, P.pv_y0 + ly, ps.rng);   // k_generate
    else ps.ray = camera_ray(P, lx, ly, frame, ps.rng);
    ps.thr = mk3(1.0f); ps.rad = mk3(0.0f); ps.absn = mk3(0.0f); ps.bsdf_pdf = 0.f;
    ps.stale = xyz(ldg4(S.materials + 1));
    for (int depth = 0; depth < P.max_depth; depth++) {
        Hit h;
        trace<false, CULL, false>(S, ps.ray, 0.f, h, stk, nullptr);
        Nee nee;
        Surf sf;
        f3 absnNext;
        bool go = shade_hit<false, true, true, true>(S, P, depth, ps, h, nee, sf, absnNext, nullptr);
        if (go) {
            f3 Li = mk3(0.0f);
            Ray sr; sr.o = nee.origin;
            Hit dummy;
            if (nee.has0) { sr.d = nee.d0; if (!trace<true, CULL, false>(S, sr, nee.m0, dummy, stk, nullptr)) Li = Li + nee.c0; }
            if (nee.has1) { sr.d = nee.d1; if (!trace<true, CULL, false>(S, sr, nee.m1, dummy, stk, nullptr)) Li = Li + nee.c1; }
            ps.rad = ps.rad + Li * nee.T;
            go = shade_sample(P, depth, ps, sf, h.fhp, absnNext);
        }
        if (!go) break;
    }
    return ps.rad;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

void* lfhc_open_pack(const char* path) {
    Check* c = new Check;
    std::string err;
    if (!lfpack::read(path, c->pack, &err)) { delete c; return nullptr; }
    LfSceneView v = c->pack.view();
    c->P = c->pack.params();
    c->C = c->pack.camera();
    if (!repack_scene(v, c->packed, err)) { delete c; return nullptr; }
    DevScene& D = c->S;
    const PackedScene& K = c->packed;
    D.nodes = K.nodes.data(); D.tris = K.tris.data(); D.trinrm = K.trinrm.data(); D.tri_vx = K.tri_vx.data();
    D.inst = K.inst.data(); D.materials = reinterpret_cast<const float4*>(v.materials); D.lights = K.lights.data();
    D.top_ref = K.top_ref;
    D.num_lights = v.num_lights; D.num_materials = v.num_materials; D.num_instances = v.num_instances;
    if (v.num_textures > 0 && v.texture_maps) {
        c->tex8 = HostTex{v.texture_maps, v.tex_width, v.tex_height, v.num_textures};
        D.tex_maps = reinterpret_cast<cudaTextureObject_t>(&c->tex8);
        D.tex_w = v.tex_width; D.tex_h = v.tex_height; D.num_tex = v.num_textures;
    }
    if (v.hdr_cols && v.hdr_width > 0 && v.hdr_height > 0) {
        size_t n = (size_t)v.hdr_width * v.hdr_height;
        c->hdr_rgba.resize(n);
        for (size_t i = 0; i < n; i++) c->hdr_rgba[i] = make_float4(v.hdr_cols[3 * i], v.hdr_cols[3 * i + 1], v.hdr_cols[3 * i + 2], 1.f);
        c->texf = HostTex{c->hdr_rgba.data(), v.hdr_width, v.hdr_height, 1};
        D.hdr_tex = reinterpret_cast<cudaTextureObject_t>(&c->texf);
        D.marginal = reinterpret_cast<const float2*>(v.hdr_marginal);
        D.conditional = reinterpret_cast<const float2*>(v.hdr_conditional);
        D.hdr_w = v.hdr_width; D.hdr_h = v.hdr_height;
    }
    return c;
}
void lfhc_close(void* h) { delete static_cast<Check*>(h); }

// uniforms / camera overrides (lfcuda_set_params, lfcuda_set_camera); either may be null
void lfhc_set_params(void* h, const LfParams* p, const LfCamera* cam) {
    Check* c = static_cast<Check*>(h);
    if (p) c->P = *p;
    if (cam) c->C = *cam;
}
void lfhc_get_params(void* h, LfParams* p, LfCamera* cam) {
    Check* c = static_cast<Check*>(h);
    if (p) *p = c->P;
    if (cam) *cam = c->C;
}

void lfhc_size(void* h, int* w, int* hh) { Check* c = static_cast<Check*>(h); *w = c->P.width; *hh = c->P.height; }

// accum (W*H*3, rows bottom-up like lfcuda_read_accum) += the samples of frames first, first+stride, ... in frame order
int lfhc_render_frames(void* h, int first_frame, int nframes, int stride, int cull, float* accum) {
    Check* c = static_cast<Check*>(h);
    if (c->packed.stack_depth > 64) return 1;
    const int W = c->P.width, H = c->P.height, TW = c->P.tile_width, TH = c->P.tile_height;
    if (TW != W || TH != H) return 2;   // single-tile scenes only
    DevParams D;
    fill_params(*c, D, first_frame, nframes, stride, 0, 0);
#pragma omp parallel
    {
        std::vector<int> stack((size_t)64 * kBlockThreads);
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < W * H; i++) {
            int lx = i % W, ly = i / W;
            for (int f = 0; f < nframes; f++) {
                f3 r = cull ? sample<true>(c->S, D, lx, ly, first_frame + f * stride, stack.data())
                            : sample<false>(c->S, D, lx, ly, first_frame + f * stride, stack.data());
                float* o = accum + 3 * ((size_t)ly * W + lx);
                o[0] += r.x; o[1] += r.y; o[2] += r.z;
            }
        }
    }
    return 0;
}

// one draw of the preview engine (lfcuda_render_preview): pv_w x pv_h x 3 floats, rows bottom-up
int lfhc_render_preview(void* h, int pv_w, int pv_h, int max_depth, int use_dof, float* out) {
    Check* c = static_cast<Check*>(h);
    if (c->packed.stack_depth > 64) return 1;
    DevParams D;
    fill_params(*c, D, 1, 1, 1, 0, 0);
    D.max_depth = max_depth;
    D.preview = 1; D.pv_w = pv_w; D.pv_h = pv_h; D.pv_y0 = 0; D.use_dof = use_dof ? 1 : 0;
#pragma omp parallel
    {
        std::vector<int> stack((size_t)64 * kBlockThreads);
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < pv_w * pv_h; i++) {
            f3 r = sample<true>(c->S, D, i % pv_w, i / pv_w, 1, stack.data());
            out[3 * (size_t)i] = r.x; out[3 * (size_t)i + 1] = r.y; out[3 * (size_t)i + 2] = r.z;
        }
    }
    return 0;
}

// The kernels' BSDF functions on explicit arguments: same item layout and ops as Oracle::BsdfKat (oracle/lf_oracle.cpp), compared by
// tests/test_hostcheck.py with what the reference's own GLSL functions return on llvmpipe (tests/golden/llvmpipe_bsdf.npz).
void lfhc_bsdf_kat(int op, const float* in, int n, float* out4) {
    for (int i = 0; i < n; i++) {
        const float* a = in + 36 * (size_t)i;
        float* r = out4 + 4 * (size_t)i;
        Surf s;
        std::memset(&s, 0, sizeof s);
        f3 V = mk3(a[0], a[1], a[2]), N = mk3(a[4], a[5], a[6]), L = mk3(a[8], a[9], a[10]);
        s.mat.albedo = mk3(a[12], a[13], a[14]); s.mat.specular = a[15];
        s.mat.metallic = a[16]; s.mat.roughness = a[17]; s.mat.subsurface = a[18]; s.mat.specularTint = a[19];
        s.mat.sheen = a[20]; s.mat.sheenTint = a[21]; s.mat.clearcoat = a[22]; s.mat.clearcoatRoughness = a[23];
        s.mat.specTrans = a[24]; s.eta = a[25];
        s.normal = N; s.ffnormal = N;
        s.tangent = mk3(a[28], a[29], a[30]); s.bitangent = mk3(a[32], a[33], a[34]);
        Rng g;
        g.x = (unsigned)a[31]; g.y = (unsigned)a[35]; g.z = (unsigned)a[26]; g.w = (unsigned)a[27];
        float pdf = 0.0f;
        r[0] = r[1] = r[2] = r[3] = 0.0f;
        switch (op) {
        case 0: { f3 f = DisneyEval(s, V, N, L, pdf); r[0] = f.x; r[1] = f.y; r[2] = f.z; r[3] = pdf; break; }
        case 1: { f3 Ls = mk3(0.0f); DisneySample(s, V, N, g, Ls, pdf); r[0] = Ls.x; r[1] = Ls.y; r[2] = Ls.z; r[3] = pdf; break; }
        case 2: { f3 Ls = mk3(0.0f); f3 f = DisneySample(s, V, N, g, Ls, pdf); r[0] = f.x; r[1] = f.y; r[2] = f.z; r[3] = rnd(g); break; }
        case 3: r[0] = GTR1(a[0], a[1]); r[1] = GTR2(a[0], a[1]); r[2] = SmithG_GGX(a[0], a[1]); r[3] = DielectricFresnel(a[0], a[25]); break;
        case 4: {
            f3 h1 = ImportanceSampleGTR1(a[17], a[0]), h2 = ImportanceSampleGTR2(a[17], a[0], a[1]), c = CosineSampleHemisphere(a[0], a[1]);
            r[0] = h1.x + h1.z; r[1] = h2.x + h2.z; r[2] = c.x + c.z; r[3] = h1.y + h2.y + c.y;
            break;
        }
        default: break;
        }
    }
}

// k_post for a W x H accumulation buffer
void lfhc_post_process(const float* accum, int W, int H, float inv, int tonemap, const LfPostParams* pp, float* out) {
    LfPostParams p;
    std::memset(&p, 0, sizeof p);
    if (pp) p = *pp;
    for (int i = 0; i < W * H; i++) post_pixel(accum, W, H, i, inv, tonemap, p, out + 3 * (size_t)i);
}

}  // extern "C"
#pragma GCC visibility pop
