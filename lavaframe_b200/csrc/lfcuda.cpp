// lfcuda.cpp — implementation of the C ABI in include/lfcuda.h: context, scene upload + re-pack, uniform
// handling, the wavefront launch sequence, read-backs and the NCCL sum of the accumulation buffer.
// There is NO CPU fallback: every entry point that computes needs a CUDA device and fails loudly otherwise.
#include "lfcuda.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "lf_ctx_internal.h"
#include "lf_kernels.h"
#include "lf_repack.h"

using namespace lf;

namespace {

thread_local std::string g_create_error;
constexpr int kCountRows = 5;   // rows of Queues::counts

struct StageEvent { int stage; cudaEvent_t a, b; };

// Minimal NCCL surface, bound at run time (torch ships its own libnccl; binding lazily lets both share it).
struct NcclId { char b[128]; };   // ncclUniqueId, passed by value
struct Nccl {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

}  // namespace

struct lfcuda_ctx {
    int device = 0;
    cudaDeviceProp prop{};
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;

    // scene
    bool have_scene = false;
    PackedScene packed;
    std::vector<float> host_nodes;        // reference node array (kept for lfcuda_update_instances)
    int top_index = 0, num_tri_refs = 0;
    DevScene dev{};
    std::vector<void*> scene_allocs;
    cudaArray_t tex_array = nullptr, hdr_array = nullptr;
    float4* d_nodes = nullptr; size_t nodes_cap = 0;
    float4* d_inst = nullptr;
    float4* d_materials = nullptr; int materials_cap = 0;
    // device-side TLAS rebuild (lf_tlas.cu): per-instance constants taken from the uploaded TLAS leaves, outputs and scratch
    float* d_transforms = nullptr; float* d_blas_box = nullptr;
    int *d_inst_blas_root = nullptr, *d_inst_blas_ref = nullptr, *d_inst_mat = nullptr;
    float* d_flat_tlas = nullptr; int* d_tlas_result = nullptr;
    float *d_tlas_bmin = nullptr, *d_tlas_bmax = nullptr, *d_tlas_cent = nullptr; int* d_tlas_prim = nullptr;
    void *d_tlas_q0 = nullptr, *d_tlas_q1 = nullptr;
    bool tlas_ready = false;

    // uniforms
    bool have_params = false, have_camera = false;
    LfParams params{};
    LfCamera camera{};
    LfPostParams post{};

    // path state
    size_t capacity = 0;                  // slots
    int frames_cap = 0, slots_per_frame = 0, pix_w8 = 0, pix_h4 = 0;
    std::vector<void*> state_allocs, frame_allocs;
    int state_tile_w = 0, state_tile_h = 0, state_frames_req = 0;   // what the path state was last sized for
    int frame_w = 0, frame_h = 0;                                   // ... and the accumulation / output buffers
    int counts_cap = 0;
    PathSoA soa{};
    Queues queues{};
    float* d_accum = nullptr; size_t accum_floats = 0;
    float* d_out_f = nullptr; unsigned char* d_out_u8 = nullptr;
    float* d_preview = nullptr; float* d_preview_out = nullptr;   // preview engine target (pathTraceTextureLowRes) + its post-processed copy
    size_t preview_cap = 0; int preview_w = 0, preview_h = 0;
    DevCounters* d_counters = nullptr;

    // instrumentation
    bool profiling = false;
    std::vector<StageEvent> events;
    LfStageStats stats{};
    uint64_t launches = 0;
    bool sort_rays = false;               // LF_SORT_RAYS=1|2|3: ray-sort experiment (lf_kernels.h SortCtx)
    SortCtx sort;
    int ctas_per_sm = 9;                  // CTAs per SM of the persistent traversal kernels: 9 x 128 threads x 56 registers fill the
                                          // register file exactly (measured: 8 -> 9 = -5 % extend/shadow time; 10 needs 48 registers and spills, +60 %)

    // NCCL (function table: process-wide g_nccl)
    void* comm = nullptr;
};

namespace {

int fail(lfcuda_ctx* c, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) return fail(ctx, e_ == cudaErrorMemoryAllocation ? LFCUDA_ENOMEM : LFCUDA_ECUDA, \
                                           "%s failed: %s", #call, cudaGetErrorString(e_));              \
    } while (0)

template <class T> int upload(lfcuda_ctx* ctx, const T* src, size_t n, T** dst, std::vector<void*>& owner) {
    *dst = nullptr;
    if (n == 0) return 0;
    CK(cudaMalloc((void**)dst, n * sizeof(T)));
    owner.push_back(*dst);
    CK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

void free_scene(lfcuda_ctx* c) {
    for (void* p : c->scene_allocs) cudaFree(p);
    c->scene_allocs.clear();
    if (c->dev.tex_maps) cudaDestroyTextureObject(c->dev.tex_maps);
    if (c->dev.hdr_tex) cudaDestroyTextureObject(c->dev.hdr_tex);
    if (c->dev.nodes_tex) cudaDestroyTextureObject(c->dev.nodes_tex);
    if (c->dev.tris_tex) cudaDestroyTextureObject(c->dev.tris_tex);
    if (c->tex_array) cudaFreeArray(c->tex_array);
    if (c->hdr_array) cudaFreeArray(c->hdr_array);
    c->tex_array = c->hdr_array = nullptr;
    c->dev = DevScene{};
    c->have_scene = false;
    c->tlas_ready = false;
}

// Device memory of a context comes in three independent groups, so that changing one render parameter never disturbs what
// another group holds (the reference changes maxDepth / tile uniforms without touching accumTexture, TiledRenderer.cpp:505-521):
//   frame   accumulation buffer + post-process outputs          re-created (and zeroed) only when width / height change
//   state   wavefront path state + queues                       re-created when the tile size or frames_in_flight change
//   counts  per-bounce queue counters                           grown when max_depth grows
void free_group(std::vector<void*>& g) {
    for (void* p : g) cudaFree(p);
    g.clear();
}
void free_state(lfcuda_ctx* c) {
    free_group(c->state_allocs);
    c->capacity = 0;
    c->soa = PathSoA{};
    c->queues.active[0] = c->queues.active[1] = c->queues.shadow = c->queues.sample = nullptr;
    c->sort.sorted = nullptr; c->sort.keys = nullptr; c->sort.hist = nullptr;
}
void free_frame(lfcuda_ctx* c) {
    free_group(c->frame_allocs);
    c->d_accum = nullptr; c->d_out_f = nullptr; c->d_out_u8 = nullptr; c->accum_floats = 0;
    c->d_preview = nullptr; c->d_preview_out = nullptr; c->preview_cap = 0; c->preview_w = c->preview_h = 0;
}

int alloc_frame(lfcuda_ctx* ctx, int width, int height) {
    free_frame(ctx);
    ctx->frame_w = ctx->frame_h = 0;
    const size_t n = (size_t)width * height * 3;
    auto A = [&](void** p, size_t bytes) -> int {
        CK(cudaMalloc(p, bytes));
        ctx->frame_allocs.push_back(*p);
        return 0;
    };
    int r;
    if ((r = A((void**)&ctx->d_accum, n * sizeof(float))) || (r = A((void**)&ctx->d_out_f, n * sizeof(float))) || (r = A((void**)&ctx->d_out_u8, n))) {
        free_frame(ctx);
        return r;
    }
    ctx->accum_floats = n;
    CK(cudaMemsetAsync(ctx->d_accum, 0, n * sizeof(float), ctx->stream));
    ctx->frame_w = width; ctx->frame_h = height;
    return 0;
}

int ensure_counts(lfcuda_ctx* ctx, int max_depth) {
    const int stride = std::max(max_depth, 2) + 2;   // the preview engine runs at depth 2 whatever maxDepth is (TiledRenderer.cpp:532)
    if (ctx->queues.counts && stride <= ctx->counts_cap) { ctx->queues.stride = stride; return 0; }
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->queues.counts) cudaFree(ctx->queues.counts);
    ctx->queues.counts = nullptr; ctx->counts_cap = 0;
    CK(cudaMalloc((void**)&ctx->queues.counts, (size_t)kCountRows * stride * sizeof(int)));
    ctx->counts_cap = stride;
    ctx->queues.stride = stride;
    return 0;
}

constexpr size_t kAutoSlots = (size_t)64 << 20;   // auto batch: about 64 M pixel-samples in flight (24 GB of the 180).  Measured on C2: 4 M -> 328, 8 M -> 360,
                                                  // 16 M -> 377 M samples/s with the first kernels, 16 M -> 427, 32 M -> 435, 64 M -> 436 with
                                                  // later ones: deep bounces keep enough rays to fill the persistent grid.  A 4K frame is 8.3 M slots:
                                                  // 32 M gave C4 batches of 3 frames; 8 / 16 / 32 frames per batch measured +1.0 / +1.4 / +1.5 %
                                                  // (profiles/r2/r2u_ab_c4_stress_fif*.json), so the default went to 64 M (7 frames of 4K).

// Path state for tiles of tile_w x tile_h pixels, `frames_req` frames per batch (0 = auto).  The auto batch is clamped to half of the
// device memory that is free right now, and an allocation failure retries with half the frames (a GPU shared with torch / NCCL buffers
// degrades instead of failing); an explicit frames_in_flight is honoured or refused.
int alloc_state(lfcuda_ctx* ctx, int tile_w, int tile_h, int frames_req) {
    free_state(ctx);
    ctx->state_tile_w = ctx->state_tile_h = 0;
    ctx->pix_w8 = (tile_w + 7) / 8 * 8;
    ctx->pix_h4 = (tile_h + 3) / 4 * 4;
    const size_t spf = (size_t)ctx->pix_w8 * ctx->pix_h4;
    if (spf > (size_t)INT_MAX) return fail(ctx, LFCUDA_ELIMIT, "a tile of %d x %d pixels exceeds the 2^31 slot index range", tile_w, tile_h);
    ctx->slots_per_frame = (int)spf;
    struct Arr { void** p; size_t elem; };
    PathSoA& S = ctx->soa;
    Queues& Q = ctx->queues;
    // 32-byte records of two fields each (lf_types.h Pair) + the one plain float4 array + the queues
#if LF_SAMPLE_REMAT
    void* rec[9] = {};
#else
    void* rec[10] = {};
#endif
    std::vector<Arr> arrs;
    for (void*& r : rec) arrs.push_back({&r, 32});
    arrs.push_back({(void**)&S.sh_d1, 16});
    arrs.push_back({(void**)&Q.active[0], 4}); arrs.push_back({(void**)&Q.active[1], 4});
    arrs.push_back({(void**)&Q.shadow, 4}); arrs.push_back({(void**)&Q.sample, 4});
    auto bind_records = [&]() {
        auto lo = [&](int k) { return (float4*)rec[k]; };
        S.ray_o.p = lo(0); S.ray_d.p = lo(0) + 1;
        S.hit_f.p = lo(1); S.hit_i.p = (int4*)lo(1) + 1;
        S.thr.p = lo(2); S.rad.p = lo(2) + 1;
#if LF_SAMPLE_REMAT
        S.sf0.p = lo(3); S.absn.p = lo(3) + 1;
        S.rng.p = (uint4*)lo(4); S.hit_p.p = lo(4) + 1;
        S.sf1.p = lo(5); S.sf2.p = lo(5) + 1;
        S.stale.p = lo(6); S.sh_T.p = lo(6) + 1;
        S.sh_o.p = lo(7); S.sh_d0.p = lo(7) + 1;
        S.sh_c0.p = lo(8); S.sh_c1.p = lo(8) + 1;
#else
        S.absn.p = lo(3); S.stale.p = lo(3) + 1;
        S.rng.p = (uint4*)lo(4); S.hit_p.p = lo(4) + 1;
        S.sf0.p = lo(5); S.sf1.p = lo(5) + 1;
        S.sf2.p = lo(6); S.sf3.p = lo(6) + 1;
        S.sf4.p = lo(7); S.sh_T.p = lo(7) + 1;
        S.sh_o.p = lo(8); S.sh_d0.p = lo(8) + 1;
        S.sh_c0.p = lo(9); S.sh_c1.p = lo(9) + 1;
#endif
    };
    if (ctx->sort_rays) { arrs.push_back({(void**)&ctx->sort.sorted, 4}); arrs.push_back({(void**)&ctx->sort.keys, 4}); }
    size_t per_slot = 0;
    for (const Arr& a : arrs) per_slot += a.elem;
    int F = frames_req;
    const bool autoF = F <= 0;
    if (autoF) {
        size_t slots = kAutoSlots;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) slots = std::min(slots, free_b / 2 / per_slot);
        F = (int)std::min<size_t>(256, std::max<size_t>(1, slots / spf));
    }
    if ((size_t)F * spf > (size_t)INT_MAX)
        return fail(ctx, LFCUDA_ELIMIT, "frames_in_flight %d x %zu slots per frame exceeds the 2^31 slot index range", F, spf);
    for (;;) {
        const size_t cap = (size_t)F * spf;
        cudaError_t e = cudaSuccess;
        for (const Arr& a : arrs) {
            e = cudaMalloc(a.p, cap * a.elem);
            if (e != cudaSuccess) { *a.p = nullptr; break; }
            ctx->state_allocs.push_back(*a.p);
        }
        if (e == cudaSuccess && ctx->sort_rays) {
            e = cudaMalloc((void**)&ctx->sort.hist, (size_t)kSortBins * sizeof(unsigned));
            if (e == cudaSuccess) ctx->state_allocs.push_back(ctx->sort.hist);
        }
        if (e == cudaSuccess) { bind_records(); ctx->capacity = cap; break; }
        free_state(ctx);
        cudaGetLastError();   // clear the sticky allocation error
        if (e != cudaErrorMemoryAllocation) return fail(ctx, LFCUDA_ECUDA, "path state allocation failed: %s", cudaGetErrorString(e));
        if (!autoF || F == 1)
            return fail(ctx, LFCUDA_ENOMEM, "path state of %d frames x %zu slots (%zu bytes per slot) does not fit the device", F, spf, per_slot);
        F = std::max(1, F / 2);
    }
    ctx->frames_cap = F;
    ctx->state_tile_w = tile_w; ctx->state_tile_h = tile_h; ctx->state_frames_req = frames_req;
    return 0;
}

// tan() of the reference shader as llvmpipe evaluates it: sin(x) * (1 / cos(x)) with the octant-reduced polynomial sincos of
// lf_math.cuh, restated for the host (this file is compiled with -ffp-contract=off).  The kernels read the result as
// DevParams::cam_scale; evaluating it once here instead of once per sample changes nothing in the arithmetic.
static float glsl_tan(float x) {
    float ax = std::fabs(x);
    int j = (int)(ax * 1.27323954473516f);
    j += (j & 1);
    float y = (float)j;
    float r = ((ax - y * 0.78515625f) - y * 2.4187564849853515625e-4f) - y * 3.77489497744594108e-8f;
    float z = r * r;
    float ps = ((-1.9515295891E-4f * z + 8.3321608736E-3f) * z - 1.6666654611E-1f) * z * r + r;
    float pc = ((2.443315711809948E-005f * z - 1.388731625493765E-003f) * z + 4.166664568298827E-002f) * z * z - 0.5f * z + 1.0f;
    int q = j & 7;
    bool swap = (q == 2) || (q == 6);
    float sv = swap ? pc : ps, cv = swap ? ps : pc;
    if (q == 4 || q == 6) sv = -sv;
    if (q == 2 || q == 4) cv = -cv;
    if (x < 0.0f) sv = -sv;
    return sv * (1.0f / cv);
}

void fill_dev_params(const lfcuda_ctx* c, DevParams& D, int first_frame, int nframes, int stride, int tile_x, int tile_y) {
    const LfParams& P = c->params;
    const LfCamera& C = c->camera;
    std::memset(&D, 0, sizeof D);
    D.width = P.width; D.height = P.height; D.tile_w = P.tile_width; D.tile_h = P.tile_height;
    D.max_depth = P.max_depth; D.enable_rr = P.enable_rr; D.rr_depth = P.rr_depth;
    D.use_envmap = (P.use_envmap && c->dev.hdr_w > 0) ? 1 : 0;
    D.use_constant_bg = P.use_constant_bg;
    for (int k = 0; k < 3; k++) {
        D.bg[k] = P.bg_color[k];
        D.cam_pos[k] = C.position[k]; D.cam_right[k] = C.right[k]; D.cam_up[k] = C.up[k]; D.cam_fwd[k] = C.forward[k];
    }
    D.hdr_multiplier = P.hdr_multiplier;
    D.hdr_resolution = (float)(c->dev.hdr_w * c->dev.hdr_h);          // TiledRenderer.cpp:222
    D.inv_tiles_x = 1.0f / ((float)P.width / P.tile_width);            // TiledRenderer.cpp:226-227
    D.inv_tiles_y = 1.0f / ((float)P.height / P.tile_height);
    D.cam_scale = glsl_tan(C.fov * 0.5f);                              // renderer.glsl:51
    D.focal_dist = C.focal_dist; D.aperture = C.aperture;
    D.tile_x = tile_x; D.tile_y = tile_y;
    D.first_frame = first_frame; D.frame_stride = stride; D.num_frames = nframes;
    D.pix_w8 = c->pix_w8; D.pix_h4 = c->pix_h4; D.slots_per_frame = c->slots_per_frame;
}

struct StageTimer {
    lfcuda_ctx* c; int stage; cudaEvent_t a = nullptr, b = nullptr;
    StageTimer(lfcuda_ctx* c_, int stage_) : c(c_), stage(stage_) {
        c->launches++;
        c->stats.launches[stage]++;
        if (c->profiling) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, c->stream); }
    }
    ~StageTimer() {
        if (c->profiling) { cudaEventRecord(b, c->stream); c->events.push_back({stage, a, b}); }
    }
};

int check_ready(lfcuda_ctx* ctx) {
    if (!ctx) return LFCUDA_EINVAL;
    if (!ctx->have_scene) return fail(ctx, LFCUDA_EINVAL, "no scene uploaded (lfcuda_upload_scene)");
    if (!ctx->have_params) return fail(ctx, LFCUDA_EINVAL, "render parameters not set (lfcuda_set_params)");
    if (!ctx->have_camera) return fail(ctx, LFCUDA_EINVAL, "camera not set (lfcuda_set_camera)");
    return 0;
}

void make_launch_ctx(lfcuda_ctx* c, LaunchCtx& L, const DevParams& D) {
    L.scene = c->dev; L.params = D; L.soa = c->soa; L.queues = c->queues; L.counters = c->d_counters; L.stream = c->stream;
    L.sm_count = c->prop.multiProcessorCount;
    // the 64-entry stack variant needs 38.4 KB of shared memory per CTA: 5 CTAs per SM are resident, not 9
    L.persistent_blocks = L.sm_count * (c->packed.stack_depth > 32 ? std::min(c->ctas_per_sm, 5) : c->ctas_per_sm);
    L.stack_depth = c->packed.stack_depth;
    L.cull = !c->params.no_cull;
    L.count = c->params.count_work != 0;
    L.sort = (c->sort_rays && c->sort.sorted) ? &c->sort : nullptr;
}

// One batch of `nframes` (<= frames_cap) frames of one tile through the pipeline.
int run_batch(lfcuda_ctx* ctx, int first_frame, int nframes, int stride, int tile_x, int tile_y, bool accumulate) {
    DevParams D;
    fill_dev_params(ctx, D, first_frame, nframes, stride, tile_x, tile_y);
    if (D.max_depth <= 0) return 0;   // maxDepth 0: the bounce loop never runs (pathtrace.glsl:218), every sample adds zero radiance
    LaunchCtx L;
    make_launch_ctx(ctx, L, D);
    if (ctx->params.kernel_mode == 1) {
        { StageTimer t(ctx, LF_STAGE_MEGAKERNEL); launch_megakernel(L); }
    } else {
        CK(cudaMemsetAsync(ctx->queues.counts, 0, (size_t)kCountRows * ctx->queues.stride * sizeof(int), ctx->stream));
        { StageTimer t(ctx, LF_STAGE_GENERATE); launch_generate(L); }
        for (int d = 0; d < D.max_depth; d++) {
            { StageTimer t(ctx, LF_STAGE_EXTEND); launch_extend(L, d); }
            { StageTimer t(ctx, LF_STAGE_SHADE); launch_shade(L, d); }
            if (shadow_before_sample()) { StageTimer t(ctx, LF_STAGE_SHADOW); launch_shadow(L, d); }
            if (d + 1 < D.max_depth && !shade_is_fused(L)) { StageTimer t(ctx, LF_STAGE_SAMPLE); launch_sample(L, d); }
            if (!shadow_before_sample()) { StageTimer t(ctx, LF_STAGE_SHADOW); launch_shadow(L, d); }
        }
    }
    if (accumulate) { StageTimer t(ctx, LF_STAGE_ACCUMULATE); launch_accumulate(L, ctx->d_accum); }
    CK(cudaGetLastError());
    return 0;
}

// The NCCL function table is bound once per process (dlopen is reference counted; the handle is kept for the process' life).
Nccl g_nccl;
std::mutex g_nccl_mutex;
bool load_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.lib) return g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return false;
    *(void**)&g_nccl.GetUniqueId = dlsym(g_nccl.lib, "ncclGetUniqueId");
    *(void**)&g_nccl.CommInitRank = dlsym(g_nccl.lib, "ncclCommInitRank");
    *(void**)&g_nccl.CommInitAll = dlsym(g_nccl.lib, "ncclCommInitAll");
    *(void**)&g_nccl.AllReduce = dlsym(g_nccl.lib, "ncclAllReduce");
    *(void**)&g_nccl.GroupStart = dlsym(g_nccl.lib, "ncclGroupStart");
    *(void**)&g_nccl.GroupEnd = dlsym(g_nccl.lib, "ncclGroupEnd");
    *(void**)&g_nccl.CommDestroy = dlsym(g_nccl.lib, "ncclCommDestroy");
    *(void**)&g_nccl.GetErrorString = dlsym(g_nccl.lib, "ncclGetErrorString");
    return g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy;
}
const char* nccl_err(int rc) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"; }
// ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0; checked against the header of the NCCL this image ships)
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

}  // namespace

namespace lf {
bool ctx_view(lfcuda_ctx* c, CtxView* v) {
    if (!c || !c->have_params || !c->d_accum) return false;
    v->device = c->device; v->stream = c->stream; v->accum = c->d_accum; v->accum_floats = c->accum_floats;
    v->out_f = c->d_out_f; v->out_u8 = c->d_out_u8; v->width = c->params.width; v->height = c->params.height; v->post = c->post;
    return true;
}
void ctx_count_launch(lfcuda_ctx* c) { c->launches++; }
int ctx_fail(lfcuda_ctx* c, int code, const char* msg) { return fail(c, code, "%s", msg); }
}  // namespace lf

extern "C" {

int lfcuda_abi_version(void) { return LFCUDA_ABI_VERSION; }

int lfcuda_create(lfcuda_ctx** out, int device) {
    lfcuda_ctx* ctx = nullptr;   // CK reports into g_create_error while ctx == nullptr
    if (!out) return fail(nullptr, LFCUDA_EINVAL, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, LFCUDA_ECUDA, "no CUDA device available (%s); liblfcuda has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, LFCUDA_EINVAL, "device %d out of range (%d devices)", device, ndev);
    CK(cudaSetDevice(device));
    lfcuda_ctx* c = new lfcuda_ctx;
    c->device = device;
    ctx = c;
    cudaError_t e2 = cudaGetDeviceProperties(&c->prop, device);
    if (e2 == cudaSuccess) e2 = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&c->d_counters, sizeof(DevCounters));
    if (e2 == cudaSuccess) e2 = cudaMemset(c->d_counters, 0, sizeof(DevCounters));
    if (e2 != cudaSuccess) {
        fail(nullptr, LFCUDA_ECUDA, "context setup failed: %s", cudaGetErrorString(e2));
        delete c;
        return LFCUDA_ECUDA;
    }
    c->stream = c->own_stream;
    if (const char* e = getenv("LF_CTAS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 16) c->ctas_per_sm = v; }
    if (const char* e = getenv("LF_SORT_RAYS")) { c->sort.mode = atoi(e) & 3; c->sort_rays = c->sort.mode != 0; }
    *out = c;
    return 0;
}

void lfcuda_destroy(lfcuda_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (auto& ev : c->events) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    free_scene(c);
    free_state(c);
    free_frame(c);
    if (c->queues.counts) cudaFree(c->queues.counts);
    cudaFree(c->d_counters);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char* lfcuda_last_error(const lfcuda_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int lfcuda_set_stream(lfcuda_ctx* ctx, void* cuda_stream) {
    if (!ctx) return LFCUDA_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return 0;
}

int lfcuda_synchronize(lfcuda_ctx* ctx) {
    if (!ctx) return LFCUDA_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int upload_scene_impl(lfcuda_ctx* ctx, const LfSceneView* v);
int lfcuda_upload_scene(lfcuda_ctx* ctx, const LfSceneView* v) {
    if (!ctx || !v) return LFCUDA_EINVAL;
    int r = upload_scene_impl(ctx, v);
    if (r) {   // the header promises that host arrays are copied during the call: no copy may still be in flight when we report
        std::string err = ctx->err;
        cudaStreamSynchronize(ctx->stream);
        free_scene(ctx);
        ctx->err = err;
    }
    return r;
}
static int upload_scene_impl(lfcuda_ctx* ctx, const LfSceneView* v) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    free_scene(ctx);
    ctx->packed = PackedScene{};
    std::string err;
    if (!repack_scene(*v, ctx->packed, err)) return fail(ctx, LFCUDA_ELIMIT, "scene rejected: %s", err.c_str());
    ctx->host_nodes.assign(v->bvh_nodes, v->bvh_nodes + (size_t)9 * v->num_nodes);
    ctx->top_index = v->top_bvh_index;
    ctx->num_tri_refs = v->num_tri_refs;
    const PackedScene& P = ctx->packed;
    DevScene& D = ctx->dev;
    int r;
    float4 *nodes, *tris, *trinrm, *inst, *mats, *lights; int* trivx;
    // inner-node buffer sized for the worst case so that a TLAS rebuild (update_instances) never reallocates
    ctx->nodes_cap = std::max<size_t>(P.nodes.size(), (size_t)4 * v->num_nodes);
    CK(cudaMalloc((void**)&nodes, ctx->nodes_cap * sizeof(float4)));
    ctx->scene_allocs.push_back(nodes);
    CK(cudaMemcpyAsync(nodes, P.nodes.data(), P.nodes.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    if ((r = upload(ctx, P.tris.data(), P.tris.size(), &tris, ctx->scene_allocs))) return r;
    if ((r = upload(ctx, P.trinrm.data(), P.trinrm.size(), &trinrm, ctx->scene_allocs))) return r;
    if ((r = upload(ctx, P.tri_vx.data(), P.tri_vx.size(), &trivx, ctx->scene_allocs))) return r;
    if ((r = upload(ctx, P.inst.data(), P.inst.size(), &inst, ctx->scene_allocs))) return r;
    if ((r = upload(ctx, reinterpret_cast<const float4*>(v->materials), (size_t)7 * v->num_materials, &mats, ctx->scene_allocs))) return r;
    if ((r = upload(ctx, P.lights.data(), P.lights.size(), &lights, ctx->scene_allocs))) return r;
    ctx->d_nodes = nodes; ctx->d_inst = inst; ctx->d_materials = mats; ctx->materials_cap = v->num_materials;
    D.nodes = nodes; D.tris = tris; D.trinrm = trinrm; D.tri_vx = trivx; D.inst = inst; D.materials = mats; D.lights = lights;
    {   // the same node buffer seen through the texture unit (linear float4 texture; only the LF_NODE_TEX kernel variants read it)
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = nodes;
        rd.res.linear.desc = cudaCreateChannelDesc<float4>();
        rd.res.linear.sizeInBytes = ctx->nodes_cap * sizeof(float4);
        cudaTextureDesc td = {};
        td.readMode = cudaReadModeElementType;
        CK(cudaCreateTextureObject(&D.nodes_tex, &rd, &td, nullptr));
        rd.res.linear.devPtr = tris;
        rd.res.linear.sizeInBytes = P.tris.size() * sizeof(float4);
        if (!P.tris.empty()) CK(cudaCreateTextureObject(&D.tris_tex, &rd, &td, nullptr));
    }
    if (P.top_ref >= 0 && (size_t)4 * P.top_ref + 3 < P.nodes.size()) {   // scene bounds for the ray-sort grid: the root's two child boxes
        const float4* n = P.nodes.data() + (size_t)4 * P.top_ref;
        const float lmin[3] = {n[0].x, n[0].y, n[0].z}, lmax[3] = {n[0].w, n[1].x, n[1].y};
        const float rmin[3] = {n[1].z, n[1].w, n[2].x}, rmax[3] = {n[2].y, n[2].z, n[2].w};
        for (int k = 0; k < 3; k++) {
            float lo = std::min(lmin[k], rmin[k]), hi = std::max(lmax[k], rmax[k]);
            ctx->sort.lo[k] = lo;
            ctx->sort.inv[k] = hi > lo ? (float)(1 << kSortCellBits) / (hi - lo) : 0.f;
        }
    }
    {   // what the device-side TLAS rebuild needs per instance: BLAS root (flat index, packed reference, box) and materialID, from the TLAS leaves
        const int ni = v->num_instances;
        std::vector<int> root(ni, 0), ref(ni, 0), mat(ni, 0);
        std::vector<float> box((size_t)6 * ni, 0.f);
        for (int i = v->top_bvh_index; i < v->num_nodes; i++) {
            const float* nd = v->bvh_nodes + 9 * (size_t)i;
            const int leaf = (int)nd[8];
            if (leaf >= 0) continue;
            const int inst = -leaf - 1;
            if (inst < 0 || inst >= ni) continue;
            root[inst] = (int)nd[6]; mat[inst] = (int)nd[7];
            std::memcpy(&ref[inst], &P.inst[(size_t)kInstStride * inst + 3].x, 4);
            std::memcpy(&box[(size_t)6 * inst], v->bvh_nodes + 9 * (size_t)root[inst], 6 * sizeof(float));
        }
        float* dbox; int *droot, *dref, *dmat;
        if ((r = upload(ctx, box.data(), box.size(), &dbox, ctx->scene_allocs))) return r;
        if ((r = upload(ctx, root.data(), root.size(), &droot, ctx->scene_allocs))) return r;
        if ((r = upload(ctx, ref.data(), ref.size(), &dref, ctx->scene_allocs))) return r;
        if ((r = upload(ctx, mat.data(), mat.size(), &dmat, ctx->scene_allocs))) return r;
        CK(cudaStreamSynchronize(ctx->stream));             // the vectors above go out of scope
        ctx->d_blas_box = dbox; ctx->d_inst_blas_root = droot; ctx->d_inst_blas_ref = dref; ctx->d_inst_mat = dmat;
        auto S = [&](void** p, size_t bytes) -> int {
            CK(cudaMalloc(p, std::max<size_t>(bytes, 16)));
            ctx->scene_allocs.push_back(*p);
            return 0;
        };
        if ((r = S((void**)&ctx->d_transforms, (size_t)16 * ni * sizeof(float)))) return r;
        if ((r = S((void**)&ctx->d_flat_tlas, (size_t)9 * (2 * ni) * sizeof(float)))) return r;
        if ((r = S((void**)&ctx->d_tlas_result, 4 * sizeof(int)))) return r;
        if ((r = S((void**)&ctx->d_tlas_bmin, (size_t)3 * ni * sizeof(float)))) return r;
        if ((r = S((void**)&ctx->d_tlas_bmax, (size_t)3 * ni * sizeof(float)))) return r;
        if ((r = S((void**)&ctx->d_tlas_cent, (size_t)3 * ni * sizeof(float)))) return r;
        if ((r = S((void**)&ctx->d_tlas_prim, (size_t)ni * sizeof(int)))) return r;
        if ((r = S(&ctx->d_tlas_q0, (size_t)ni * tlas_request_bytes()))) return r;
        if ((r = S(&ctx->d_tlas_q1, (size_t)ni * tlas_request_bytes()))) return r;
        CK(cudaMemcpyAsync(ctx->d_flat_tlas, v->bvh_nodes + 9 * (size_t)v->top_bvh_index,
                           (size_t)9 * std::min(2 * ni, v->num_nodes - v->top_bvh_index) * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        ctx->tlas_ready = true;
    }
    D.top_ref = P.top_ref;
    D.num_lights = v->num_lights; D.num_materials = v->num_materials; D.num_instances = v->num_instances;

    if (v->num_textures > 0 && v->texture_maps) {   // Renderer.cpp:151-160
        cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
        cudaExtent ext = make_cudaExtent(v->tex_width, v->tex_height, v->num_textures);
        CK(cudaMalloc3DArray(&ctx->tex_array, &cd, ext, cudaArrayLayered));
        cudaMemcpy3DParms cp = {};
        cp.srcPtr = make_cudaPitchedPtr((void*)v->texture_maps, (size_t)v->tex_width * 4, v->tex_width, v->tex_height);
        cp.dstArray = ctx->tex_array;
        cp.extent = ext;
        cp.kind = cudaMemcpyHostToDevice;
        CK(cudaMemcpy3D(&cp));
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = ctx->tex_array;
        cudaTextureDesc td = {};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        CK(cudaCreateTextureObject(&D.tex_maps, &rd, &td, nullptr));
        D.tex_w = v->tex_width; D.tex_h = v->tex_height; D.num_tex = v->num_textures;
    }
    if (v->hdr_cols && v->hdr_width > 0 && v->hdr_height > 0) {   // Renderer.cpp:163-185
        size_t n = (size_t)v->hdr_width * v->hdr_height;
        std::vector<float4> rgba(n);
        for (size_t i = 0; i < n; i++) { rgba[i].x = v->hdr_cols[3 * i]; rgba[i].y = v->hdr_cols[3 * i + 1]; rgba[i].z = v->hdr_cols[3 * i + 2]; rgba[i].w = 1.f; }
        cudaChannelFormatDesc cd = cudaCreateChannelDesc<float4>();
        CK(cudaMallocArray(&ctx->hdr_array, &cd, v->hdr_width, v->hdr_height));
        CK(cudaMemcpy2DToArray(ctx->hdr_array, 0, 0, rgba.data(), (size_t)v->hdr_width * sizeof(float4), (size_t)v->hdr_width * sizeof(float4),
                               v->hdr_height, cudaMemcpyHostToDevice));
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = ctx->hdr_array;
        cudaTextureDesc td = {};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        CK(cudaCreateTextureObject(&D.hdr_tex, &rd, &td, nullptr));
        float2 *marg, *cond;
        if ((r = upload(ctx, reinterpret_cast<const float2*>(v->hdr_marginal), (size_t)v->hdr_height, &marg, ctx->scene_allocs))) return r;
        if ((r = upload(ctx, reinterpret_cast<const float2*>(v->hdr_conditional), n, &cond, ctx->scene_allocs))) return r;
        D.marginal = marg; D.conditional = cond;
        D.hdr_w = v->hdr_width; D.hdr_h = v->hdr_height;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_scene = true;
    return 0;
}

int lfcuda_update_instances(lfcuda_ctx* ctx, const float* transforms, int32_t num_instances, const float* materials, int32_t num_materials,
                            const float* tlas_nodes, int32_t first_node, int32_t num_tlas_nodes) {
    if (!ctx || !ctx->have_scene) return fail(ctx, LFCUDA_EINVAL, "no scene uploaded");
    if (!transforms || num_instances != ctx->dev.num_instances) return fail(ctx, LFCUDA_EINVAL, "instance count changed (%d != %d)", num_instances, ctx->dev.num_instances);
    int total = (int)(ctx->host_nodes.size() / 9);
    if (!tlas_nodes || first_node < 0 || first_node + num_tlas_nodes > total) return fail(ctx, LFCUDA_EINVAL, "TLAS node range out of bounds");
    if (materials && num_materials > ctx->materials_cap) return fail(ctx, LFCUDA_EINVAL, "material count grew (%d > %d)", num_materials, ctx->materials_cap);
    CK(cudaSetDevice(ctx->device));
    std::memcpy(ctx->host_nodes.data() + (size_t)9 * first_node, tlas_nodes, (size_t)9 * num_tlas_nodes * sizeof(float));
    std::string err;
    if (!repack_instances(ctx->host_nodes.data(), total, ctx->top_index, transforms, num_instances, ctx->packed, err))
        return fail(ctx, LFCUDA_ELIMIT, "instance update rejected: %s", err.c_str());
    if (ctx->packed.nodes.size() > ctx->nodes_cap) return fail(ctx, LFCUDA_ELIMIT, "inner node count grew past the uploaded capacity");
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_nodes, ctx->packed.nodes.data(), ctx->packed.nodes.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_inst, ctx->packed.inst.data(), ctx->packed.inst.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    if (materials) {
        CK(cudaMemcpyAsync(ctx->d_materials, materials, (size_t)28 * num_materials * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        ctx->dev.num_materials = num_materials;
    }
    ctx->dev.top_ref = ctx->packed.top_ref;
    if (ctx->tlas_ready)   // the device's copy of the flat TLAS follows (lfcuda_read_tlas_nodes)
        CK(cudaMemcpyAsync(ctx->d_flat_tlas, tlas_nodes, (size_t)9 * std::min(num_tlas_nodes, 2 * num_instances) * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// The same update with the TLAS rebuilt on the device (lf_tlas.cu): only the matrices (and the materials) cross the bus.
int lfcuda_update_instances_device(lfcuda_ctx* ctx, const float* transforms, int32_t num_instances, const float* materials, int32_t num_materials,
                                   const int32_t* instance_material_ids) {
    if (!ctx || !ctx->have_scene || !ctx->tlas_ready) return fail(ctx, LFCUDA_EINVAL, "no scene uploaded");
    if (!transforms || num_instances != ctx->dev.num_instances) return fail(ctx, LFCUDA_EINVAL, "instance count changed (%d != %d)", num_instances, ctx->dev.num_instances);
    if (materials && num_materials > ctx->materials_cap) return fail(ctx, LFCUDA_EINVAL, "material count grew (%d > %d)", num_materials, ctx->materials_cap);
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const int n = num_instances;
    CK(cudaMemcpyAsync(ctx->d_transforms, transforms, (size_t)16 * n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (instance_material_ids) CK(cudaMemcpyAsync(ctx->d_inst_mat, instance_material_ids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (materials) {
        CK(cudaMemcpyAsync(ctx->d_materials, materials, (size_t)28 * num_materials * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        ctx->dev.num_materials = num_materials;
    }
    TlasBuild T{};
    T.transforms = ctx->d_transforms; T.blas_box = ctx->d_blas_box; T.inst_blas_root = ctx->d_inst_blas_root; T.inst_blas_ref = ctx->d_inst_blas_ref;
    T.inst_mat = ctx->d_inst_mat; T.n = n; T.top_index = ctx->top_index; T.inner_base = ctx->packed.num_blas_inner;
    T.flat_tlas = ctx->d_flat_tlas; T.packed_nodes = ctx->d_nodes; T.inst_records = ctx->d_inst; T.result = ctx->d_tlas_result;
    T.bmin = ctx->d_tlas_bmin; T.bmax = ctx->d_tlas_bmax; T.cent = ctx->d_tlas_cent; T.prim = ctx->d_tlas_prim;
    T.queue0 = ctx->d_tlas_q0; T.queue1 = ctx->d_tlas_q1;
    ctx->launches += 3;
    launch_tlas_build(ctx->stream, T, ctx->prop.multiProcessorCount);
    CK(cudaGetLastError());
    int res[2] = {0, 0};
    CK(cudaMemcpyAsync(res, ctx->d_tlas_result, sizeof res, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int depth = 2 + res[1] + ctx->packed.max_blas_height + 1;
    if (depth > 64) {
        ctx->have_scene = false;   // the device TLAS now describes a scene the 64-entry stack cannot walk: refuse to render it
        return fail(ctx, LFCUDA_ELIMIT, "instance update rejected: the rebuilt TLAS needs a traversal stack deeper than 64 entries; upload the scene again");
    }
    ctx->packed.stack_depth = depth;
    ctx->packed.top_ref = res[0];
    ctx->dev.top_ref = res[0];
    return 0;
}

// Parity probe: the flat TLAS nodes the device holds (uploaded, or rebuilt by lfcuda_update_instances_device), in the reference's layout.
int lfcuda_read_tlas_nodes(lfcuda_ctx* ctx, float* nodes_out, int32_t max_nodes, int32_t* num_nodes_out) {
    if (!ctx || !ctx->have_scene || !ctx->tlas_ready || !nodes_out) return fail(ctx, LFCUDA_EINVAL, "no scene uploaded or NULL output");
    const int n = 2 * ctx->dev.num_instances - 1;
    if (max_nodes < n) return fail(ctx, LFCUDA_EINVAL, "output holds %d nodes, the TLAS has %d", max_nodes, n);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(nodes_out, ctx->d_flat_tlas, (size_t)9 * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (num_nodes_out) *num_nodes_out = n;
    return 0;
}

int lfcuda_set_params(lfcuda_ctx* ctx, const LfParams* p) {
    if (!ctx || !p) return LFCUDA_EINVAL;
    if (p->width <= 0 || p->height <= 0 || p->tile_width <= 0 || p->tile_height <= 0 || p->max_depth < 0 || p->frames_in_flight < 0)
        return fail(ctx, LFCUDA_EINVAL, "bad resolution / tile size / depth / frames_in_flight");
    CK(cudaSetDevice(ctx->device));
    // Only a change of the RESOLUTION re-creates (and thereby clears) the accumulation buffer - the reference re-creates its FBOs then
    // too (Main.cpp reloads the renderer).  Tile size, depth, batch size and every other uniform leave the accumulated image alone.
    const bool new_frame = !ctx->d_accum || p->width != ctx->frame_w || p->height != ctx->frame_h;
    const bool new_state = !ctx->capacity || p->tile_width != ctx->state_tile_w || p->tile_height != ctx->state_tile_h ||
                           p->frames_in_flight != ctx->state_frames_req;
    if (new_frame || new_state) CK(cudaStreamSynchronize(ctx->stream));
    int r = 0;
    if (new_frame) r = alloc_frame(ctx, p->width, p->height);
    if (!r && new_state) r = alloc_state(ctx, p->tile_width, p->tile_height, p->frames_in_flight);
    if (!r) r = ensure_counts(ctx, p->max_depth);
    if (r) { ctx->have_params = false; return r; }
    ctx->params = *p;
    ctx->have_params = true;
    return 0;
}

int lfcuda_set_camera(lfcuda_ctx* ctx, const LfCamera* c) {
    if (!ctx || !c) return LFCUDA_EINVAL;
    ctx->camera = *c;
    ctx->have_camera = true;
    return 0;
}

int lfcuda_set_post(lfcuda_ctx* ctx, const LfPostParams* p) {
    if (!ctx) return LFCUDA_EINVAL;
    if (p) ctx->post = *p; else ctx->post = LfPostParams{};
    return 0;
}

int lfcuda_clear(lfcuda_ctx* ctx) {
    if (!ctx || !ctx->have_params) return fail(ctx, LFCUDA_EINVAL, "render parameters not set");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_accum, 0, ctx->accum_floats * sizeof(float), ctx->stream));
    return 0;
}

int lfcuda_render_frames(lfcuda_ctx* ctx, int32_t first_frame, int32_t nframes, int32_t frame_stride, int32_t tile_x, int32_t tile_y) {
    int r = check_ready(ctx);
    if (r) return r;
    if (nframes < 0) return fail(ctx, LFCUDA_EINVAL, "nframes < 0");
    CK(cudaSetDevice(ctx->device));
    for (int done = 0; done < nframes;) {
        int n = std::min(ctx->frames_cap, nframes - done);
        r = run_batch(ctx, first_frame + done * frame_stride, n, frame_stride, tile_x, tile_y, true);
        if (r) return r;
        done += n;
    }
    return 0;
}

// Preview engine: TiledRenderer::Render while camera->isMoving || instancesModified (TiledRenderer.cpp:327-333).
int lfcuda_render_preview(lfcuda_ctx* ctx, int32_t pv_width, int32_t pv_height, int32_t max_depth, int32_t use_dof) {
    int r = check_ready(ctx);
    if (r) return r;
    if (pv_width < 1 || pv_height < 1 || max_depth < 1) return fail(ctx, LFCUDA_EINVAL, "preview size and depth must be positive");
    if (max_depth + 2 > ctx->queues.stride) return fail(ctx, LFCUDA_EINVAL, "preview depth %d exceeds the depth the path state was sized for (%d)", max_depth, ctx->queues.stride - 2);
    CK(cudaSetDevice(ctx->device));
    const size_t need = (size_t)pv_width * pv_height * 3;
    if (need > ctx->preview_cap) {
        // grown on demand (previewScale changes reload the renderer in the reference, Main.cpp:525); freed with the frame buffers
        CK(cudaStreamSynchronize(ctx->stream));
        for (float* old : {ctx->d_preview, ctx->d_preview_out}) {   // a smaller pair from an earlier previewScale
            if (!old) continue;
            cudaFree(old);
            ctx->frame_allocs.erase(std::remove(ctx->frame_allocs.begin(), ctx->frame_allocs.end(), (void*)old), ctx->frame_allocs.end());
        }
        ctx->d_preview = ctx->d_preview_out = nullptr; ctx->preview_cap = 0;
        float *a = nullptr, *b = nullptr;
        CK(cudaMalloc((void**)&a, need * sizeof(float)));
        ctx->frame_allocs.push_back(a);
        CK(cudaMalloc((void**)&b, need * sizeof(float)));
        ctx->frame_allocs.push_back(b);
        ctx->d_preview = a; ctx->d_preview_out = b; ctx->preview_cap = need;
    }
    ctx->preview_w = pv_width; ctx->preview_h = pv_height;
    const int w8 = (pv_width + 7) / 8 * 8;
    int band = (int)std::min<size_t>((size_t)pv_height, ctx->capacity / (size_t)w8) / 4 * 4;   // rows per batch, whole 8x4 blocks
    if (band < 4) return fail(ctx, LFCUDA_ELIMIT, "preview row of %d pixels does not fit the path state (%zu slots)", pv_width, ctx->capacity);
    for (int y0 = 0; y0 < pv_height; y0 += band) {
        const int rows = std::min(band, pv_height - y0);
        DevParams D;
        fill_dev_params(ctx, D, 1, 1, 1, 0, 0);
        D.preview = 1; D.pv_w = pv_width; D.pv_h = pv_height; D.pv_y0 = y0; D.use_dof = use_dof ? 1 : 0;
        D.max_depth = max_depth;
        D.tile_w = pv_width; D.tile_h = rows;
        D.pix_w8 = w8; D.pix_h4 = (rows + 3) / 4 * 4; D.slots_per_frame = D.pix_w8 * D.pix_h4;
        LaunchCtx L;
        make_launch_ctx(ctx, L, D);
        CK(cudaMemsetAsync(ctx->queues.counts, 0, (size_t)kCountRows * ctx->queues.stride * sizeof(int), ctx->stream));
        { StageTimer t(ctx, LF_STAGE_GENERATE); launch_generate(L); }
        for (int d = 0; d < D.max_depth; d++) {
            { StageTimer t(ctx, LF_STAGE_EXTEND); launch_extend(L, d); }
            { StageTimer t(ctx, LF_STAGE_SHADE); launch_shade(L, d); }
            if (shadow_before_sample()) { StageTimer t(ctx, LF_STAGE_SHADOW); launch_shadow(L, d); }
            if (d + 1 < D.max_depth && !shade_is_fused(L)) { StageTimer t(ctx, LF_STAGE_SAMPLE); launch_sample(L, d); }
            if (!shadow_before_sample()) { StageTimer t(ctx, LF_STAGE_SHADOW); launch_shadow(L, d); }
        }
        { StageTimer t(ctx, LF_STAGE_ACCUMULATE); launch_preview_store(L, ctx->d_preview); }
        CK(cudaGetLastError());
    }
    return 0;
}

// The preview target through the post-process pass, as Present()/SetViewport() show it (TiledRenderer.cpp:361-364,558-562;
// invSampleCounter = 1 because Update() has reset sampleCounter to 1, :475).  pv_width * pv_height * 3 floats, rows bottom-up.
int lfcuda_read_preview(lfcuda_ctx* ctx, int32_t tonemap_index, int32_t is_in_preview, float* rgb_out) {
    if (!ctx || !rgb_out) return fail(ctx, LFCUDA_EINVAL, "NULL context or output");
    if (!ctx->d_preview || ctx->preview_w < 1) return fail(ctx, LFCUDA_EINVAL, "no preview rendered (lfcuda_render_preview)");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->preview_w * ctx->preview_h * 3;
    ctx->launches++;
    LfPostParams pp = ctx->post;
    if (is_in_preview) pp.use_ca = 0;                       // postprocess.glsl:128-132: `if (!isInPreview) { ... useCA ... } else plain fetch`
    launch_post(ctx->stream, ctx->d_preview, ctx->d_preview_out, nullptr, ctx->preview_w, ctx->preview_h, 1.0f, tonemap_index, pp);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(rgb_out, ctx->d_preview_out, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int lfcuda_read_accum(lfcuda_ctx* ctx, float* rgb_out) {
    if (!ctx || !ctx->have_params || !rgb_out) return fail(ctx, LFCUDA_EINVAL, "render parameters not set or NULL output");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(rgb_out, ctx->d_accum, ctx->accum_floats * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int lfcuda_read_output(lfcuda_ctx* ctx, float inv_sample_counter, int32_t tonemap_index, float* rgb_out) {
    if (!ctx || !ctx->have_params || !rgb_out) return fail(ctx, LFCUDA_EINVAL, "render parameters not set or NULL output");
    CK(cudaSetDevice(ctx->device));
    ctx->launches++;
    launch_post(ctx->stream, ctx->d_accum, ctx->d_out_f, nullptr, ctx->params.width, ctx->params.height, inv_sample_counter, tonemap_index, ctx->post);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(rgb_out, ctx->d_out_f, ctx->accum_floats * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int lfcuda_read_output_u8(lfcuda_ctx* ctx, float inv_sample_counter, int32_t tonemap_index, uint8_t* rgb_out) {
    if (!ctx || !ctx->have_params || !rgb_out) return fail(ctx, LFCUDA_EINVAL, "render parameters not set or NULL output");
    CK(cudaSetDevice(ctx->device));
    ctx->launches++;
    launch_post(ctx->stream, ctx->d_accum, nullptr, ctx->d_out_u8, ctx->params.width, ctx->params.height, inv_sample_counter, tonemap_index, ctx->post);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(rgb_out, ctx->d_out_u8, ctx->accum_floats, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int lfcuda_accum_device_ptr(lfcuda_ctx* ctx, void** dev_ptr, size_t* num_floats) {
    if (!ctx || !ctx->have_params) return fail(ctx, LFCUDA_EINVAL, "render parameters not set");
    if (dev_ptr) *dev_ptr = ctx->d_accum;
    if (num_floats) *num_floats = ctx->accum_floats;
    return 0;
}

int lfcuda_read_primary_hits(lfcuda_ctx* ctx, int32_t frame, float* t_out, int32_t* tri_out, int32_t* mat_out, int32_t* emitter_out) {
    int r = check_ready(ctx);
    if (r) return r;
    if (!t_out || !tri_out || !mat_out || !emitter_out) return fail(ctx, LFCUDA_EINVAL, "NULL output");
    CK(cudaSetDevice(ctx->device));
    // The probe covers the full frame as one tile, like the llvmpipe probe shader run on a single-tile scene.  It never touches the
    // accumulation buffer: it runs in the existing path state when one full frame fits it, and otherwise re-sizes the path state
    // (only that group) for the duration of the call.  Parameters and state are restored on every exit path.
    const LfParams saved = ctx->params;
    const int saved_w8 = ctx->pix_w8, saved_h4 = ctx->pix_h4, saved_spf = ctx->slots_per_frame;
    const int saved_tw = ctx->state_tile_w, saved_th = ctx->state_tile_h, saved_req = ctx->state_frames_req;
    const size_t need = (size_t)((saved.width + 7) / 8 * 8) * ((saved.height + 3) / 4 * 4);
    bool resized = false;
    const size_t n = (size_t)saved.width * saved.height;
    float* d_t = nullptr; int *d_tri = nullptr, *d_mat = nullptr, *d_em = nullptr;
    auto body = [&]() -> int {
        CK(cudaStreamSynchronize(ctx->stream));
        if (need > ctx->capacity) {
            resized = true;
            int rr = alloc_state(ctx, saved.width, saved.height, 1);
            if (rr) return rr;
        } else {
            ctx->pix_w8 = (saved.width + 7) / 8 * 8; ctx->pix_h4 = (saved.height + 3) / 4 * 4; ctx->slots_per_frame = (int)need;
        }
        ctx->params.tile_width = saved.width; ctx->params.tile_height = saved.height; ctx->params.kernel_mode = 0;
        CK(cudaMalloc((void**)&d_t, n * 4)); CK(cudaMalloc((void**)&d_tri, n * 4)); CK(cudaMalloc((void**)&d_mat, n * 4)); CK(cudaMalloc((void**)&d_em, n * 4));
        DevParams D;
        fill_dev_params(ctx, D, frame, 1, 1, 0, 0);
        LaunchCtx L;
        make_launch_ctx(ctx, L, D);
        CK(cudaMemsetAsync(ctx->queues.counts, 0, (size_t)kCountRows * ctx->queues.stride * sizeof(int), ctx->stream));
        { StageTimer t(ctx, LF_STAGE_GENERATE); launch_generate(L); }
        { StageTimer t(ctx, LF_STAGE_EXTEND); launch_extend(L, 0); }
        ctx->launches++;
        launch_export_hits(L, d_t, d_tri, d_mat, d_em);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(t_out, d_t, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(tri_out, d_tri, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(mat_out, d_mat, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(emitter_out, d_em, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return 0;
    };
    r = body();
    std::string err = ctx->err;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_t); cudaFree(d_tri); cudaFree(d_mat); cudaFree(d_em);
    ctx->params = saved;
    ctx->pix_w8 = saved_w8; ctx->pix_h4 = saved_h4; ctx->slots_per_frame = saved_spf;
    if (resized) {
        int r2 = alloc_state(ctx, saved_tw, saved_th, saved_req);
        if (r2) { ctx->have_params = false; if (!r) return r2; }
    }
    if (r) ctx->err = err;
    return r;
}

int lfcuda_nccl_unique_id(void* id128_out) {
    if (!id128_out) return LFCUDA_EINVAL;
    if (!load_nccl()) return fail(nullptr, LFCUDA_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
    int rc = g_nccl.GetUniqueId(id128_out);
    return rc == 0 ? 0 : fail(nullptr, LFCUDA_ENCCL, "ncclGetUniqueId failed (%d)", rc);
}

int lfcuda_nccl_init(lfcuda_ctx* ctx, const void* id128, int32_t rank, int32_t nranks) {
    if (!ctx || !id128) return LFCUDA_EINVAL;
    if (!load_nccl()) return fail(ctx, LFCUDA_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
    CK(cudaSetDevice(ctx->device));
    NcclId id;
    std::memcpy(&id, id128, 128);
    int rc = g_nccl.CommInitRank(&ctx->comm, nranks, id, rank);
    if (rc != 0) return fail(ctx, LFCUDA_ENCCL, "ncclCommInitRank failed: %s", nccl_err(rc));
    return 0;
}

int lfcuda_reduce(lfcuda_ctx* ctx) {
    if (!ctx || !ctx->have_params) return fail(ctx, LFCUDA_EINVAL, "render parameters not set");
    if (!ctx->comm) return fail(ctx, LFCUDA_ENCCL, "NCCL communicator not initialised (lfcuda_nccl_init)");
    CK(cudaSetDevice(ctx->device));
    int rc = g_nccl.AllReduce(ctx->d_accum, ctx->d_accum, ctx->accum_floats, kNcclFloat32, kNcclSum, ctx->comm, ctx->stream);
    if (rc != 0) return fail(ctx, LFCUDA_ENCCL, "ncclAllReduce failed: %s", nccl_err(rc));
    return 0;
}

int lfcuda_reset_counters(lfcuda_ctx* ctx) {
    if (!ctx) return LFCUDA_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(DevCounters), ctx->stream));
    return 0;
}

int lfcuda_get_counters(lfcuda_ctx* ctx, LfCounters* out) {
    if (!ctx || !out) return LFCUDA_EINVAL;
    static_assert(sizeof(LfCounters) == sizeof(DevCounters), "counter layouts must match");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out, ctx->d_counters, sizeof(DevCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int lfcuda_set_profiling(lfcuda_ctx* ctx, int32_t on) {
    if (!ctx) return LFCUDA_EINVAL;
    ctx->profiling = on != 0;
    return 0;
}

int lfcuda_get_stage_stats(lfcuda_ctx* ctx, LfStageStats* out) {
    if (!ctx || !out) return LFCUDA_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto& ev : ctx->events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess) ctx->stats.ms[ev.stage] += ms;
        cudaEventDestroy(ev.a); cudaEventDestroy(ev.b);
    }
    ctx->events.clear();
    *out = ctx->stats;
    return 0;
}

int lfcuda_measure_read_bandwidth(lfcuda_ctx* ctx, size_t bytes, int32_t iters, double* gbps_out) {
    if (!ctx || !gbps_out || bytes < 4096 || iters < 1) return fail(ctx, LFCUDA_EINVAL, "bad bandwidth probe arguments");
    CK(cudaSetDevice(ctx->device));
    size_t n4 = bytes / sizeof(float4);
    float4* buf = nullptr; float* sink = nullptr;
    cudaEvent_t a = nullptr, b = nullptr;
    double best = 0.0;
    auto body = [&]() -> int {
        CK(cudaMalloc((void**)&buf, n4 * sizeof(float4)));
        CK(cudaMalloc((void**)&sink, sizeof(float)));
        CK(cudaMemsetAsync(buf, 0, n4 * sizeof(float4), ctx->stream));
        int blocks = ctx->prop.multiProcessorCount * 8;
        // enough passes per launch that the launch itself is negligible (>= ~256 MB read per launch)
        int passes = (int)std::max<size_t>(1, ((size_t)256 << 20) / (n4 * sizeof(float4)));
        CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        launch_read_probe(ctx->stream, buf, n4, passes, sink, blocks);   // warm-up (fills L2 when the buffer fits)
        for (int it = 0; it < iters; it++) {
            CK(cudaEventRecord(a, ctx->stream));
            launch_read_probe(ctx->stream, buf, n4, passes, sink, blocks);
            CK(cudaEventRecord(b, ctx->stream));
            CK(cudaEventSynchronize(b));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, a, b));
            double gbps = (double)n4 * sizeof(float4) * passes / (ms * 1e6);
            if (gbps > best) best = gbps;
            ctx->launches++;
        }
        return 0;
    };
    int r = body();
    cudaStreamSynchronize(ctx->stream);
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    cudaFree(buf); cudaFree(sink);
    if (!r) *gbps_out = best;
    return r;
}

int lfcuda_measure_node_fetch(lfcuda_ctx* ctx, size_t table_bytes, int32_t iters, double* gbps_out) {
    if (!ctx || !gbps_out || table_bytes < 4096 || iters < 1) return fail(ctx, LFCUDA_EINVAL, "bad node-fetch probe arguments");
    CK(cudaSetDevice(ctx->device));
    unsigned nnodes = 1;
    while ((size_t)nnodes * 2 * 64 <= table_bytes) nnodes *= 2;          // power of two records of 64 bytes
    float4* buf = nullptr; unsigned* sink = nullptr;
    cudaEvent_t a = nullptr, b = nullptr;
    double best = 0.0;
    auto body = [&]() -> int {
        CK(cudaMalloc((void**)&buf, (size_t)nnodes * 64));
        CK(cudaMalloc((void**)&sink, sizeof(unsigned)));
        CK(cudaMemsetAsync(buf, 0, (size_t)nnodes * 64, ctx->stream));
        const int blocks = ctx->prop.multiProcessorCount * 8, steps = 2000;
        CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        launch_node_probe(ctx->stream, buf, nnodes, 200, sink, blocks);   // warm-up
        for (int it = 0; it < iters; it++) {
            CK(cudaEventRecord(a, ctx->stream));
            launch_node_probe(ctx->stream, buf, nnodes, steps, sink, blocks);
            CK(cudaEventRecord(b, ctx->stream));
            CK(cudaEventSynchronize(b));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, a, b));
            double gbps = (double)blocks * 128 * steps * 64 / (ms * 1e6);
            if (gbps > best) best = gbps;
            ctx->launches++;
        }
        return 0;
    };
    int r = body();
    cudaStreamSynchronize(ctx->stream);
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    cudaFree(buf); cudaFree(sink);
    if (!r) *gbps_out = best;
    return r;
}

int lfcuda_get_launch_count(lfcuda_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return LFCUDA_EINVAL;
    *out = ctx->launches;
    return 0;
}

}  // extern "C"
