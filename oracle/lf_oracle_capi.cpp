// lf_oracle_capi.cpp — TEST INFRASTRUCTURE: ctypes-friendly entry points of the CPU oracle (lf_oracle.h).
// Loaded only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
#include <cstring>
#include <string>

#include <omp.h>
#include <xmmintrin.h>

#include "lf_oracle.h"
#include "scenepack.h"

namespace {
struct Handle {
    lfpack::ScenePack pack;
    lforacle::Oracle* oracle = nullptr;
    ~Handle() { delete oracle; }
};
}

extern "C" {

void* lforacle_open_pack(const char* path) {
    Handle* h = new Handle;
    std::string err;
    if (!lfpack::read(path, h->pack, &err)) { delete h; return nullptr; }
    h->oracle = new lforacle::Oracle(h->pack.view(), h->pack.params(), h->pack.camera());
    return h;
}
void lforacle_close(void* hv) { delete static_cast<Handle*>(hv); }

void lforacle_get_params(void* hv, LfParams* p, LfCamera* c) {
    Handle* h = static_cast<Handle*>(hv);
    if (p) *p = h->oracle->params;
    if (c) *c = h->oracle->camera;
}
void lforacle_set_params(void* hv, const LfParams* p, const LfCamera* c) {
    Handle* h = static_cast<Handle*>(hv);
    if (p) h->oracle->params = *p;
    if (c) h->oracle->camera = *c;
}
// cull: 0 = reference behaviour (visit every pierced box), 1 = conservative distance cull
void lforacle_set_options(void* hv, int cull, int count) {
    Handle* h = static_cast<Handle*>(hv);
    h->oracle->cull = cull != 0;
    h->oracle->count = count != 0;
}
void lforacle_render_frames(void* hv, int first_frame, int nframes, int frame_stride, int tile_x, int tile_y, float* accum) {
    static_cast<Handle*>(hv)->oracle->RenderFrames(first_frame, nframes, frame_stride, tile_x, tile_y, accum);
}
// OpenMP team size of the next parallel regions.  torchrun exports OMP_NUM_THREADS=1 to every rank; bench.py's reference arm
// calls this with the host's core count so that the figure it reports is the one all host cores deliver.  Returns the team size in use.
int lforacle_set_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
    int used = 1;
#pragma omp parallel
    {
#pragma omp single
        used = omp_get_num_threads();
    }
    return used;
}
void lforacle_render_preview(void* hv, int pv_w, int pv_h, int max_depth, int use_dof, float* out) {
    static_cast<Handle*>(hv)->oracle->RenderPreview(pv_w, pv_h, max_depth, use_dof != 0, out);
}
void lforacle_primary_hits(void* hv, int frame, float* t, int32_t* tri, int32_t* mat, int32_t* emitter) {
    static_cast<Handle*>(hv)->oracle->PrimaryHits(frame, t, tri, mat, emitter);
}
void lforacle_sample(void* hv, int lx, int ly, int tile_x, int tile_y, int frame, float* rgb) {
    int px, py;
    static_cast<Handle*>(hv)->oracle->Sample(lx, ly, tile_x, tile_y, frame, rgb, &px, &py, nullptr);
}
void lforacle_rand_kat(int px, int py, int frame, int n, uint32_t* seedx, float* values) {
    lforacle::Oracle::RandKat(px, py, frame, n, seedx, values);
}
void lforacle_post_process(const float* accum, int w, int h, float inv, int tonemap_index, const LfPostParams* pp, float* out) {
    LfPostParams none;
    std::memset(&none, 0, sizeof none);
    lforacle::Oracle::PostProcess(accum, w, h, inv, tonemap_index, pp ? *pp : none, out);
}
void lforacle_bsdf_kat(int op, const float* in, int n, float* out4) { lforacle::Oracle::BsdfKat(op, in, n, out4); }
void lforacle_builtin_kat(int op, const float* in4, int n, float* out4, const uint8_t* tex, int tex_w, int tex_h, int tex_l) {
    lforacle::Oracle::BuiltinKat(op, in4, n, out4, tex, tex_w, tex_h, tex_l);
}
// MXCSR exception flags (bit 1 = a denormal OPERAND was consumed, bit 4 = a result underflowed into the denormal range) OR-ed over the
// OpenMP worker threads since the last reset.  llvmpipe runs with denormals flushed to zero, this oracle and the kernels do not: if neither
// flag is ever raised by a render, that difference cannot have influenced it (tests/test_oracle_golden.py::test_no_denormals_on_the_path).
int lforacle_fp_flags(int reset) {
    int flags = 0;
#pragma omp parallel reduction(| : flags)
    {
        unsigned csr = _mm_getcsr();
        flags |= (int)(csr & 0x3fu);
        if (reset) _mm_setcsr(csr & ~0x3fu);
    }
    return flags;
}
void lforacle_get_counters(void* hv, LfCounters* out) { *out = static_cast<Handle*>(hv)->oracle->counters; }
void lforacle_reset_counters(void* hv) { std::memset(&static_cast<Handle*>(hv)->oracle->counters, 0, sizeof(LfCounters)); }

}  // extern "C"
