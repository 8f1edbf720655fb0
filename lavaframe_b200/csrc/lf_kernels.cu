// lf_kernels.cu — the CUDA kernels of the path-tracing core (sm_100a) and their launchers.
//
// Wavefront pipeline, one launch per stage and bounce, all counts kept on the device (no host sync):
//   generate    renderer.glsl:25-62      pixel -> RNG seed, jittered thin-lens camera ray, path state reset
//   extend      closest_hit.glsl         persistent warps pull 32 rays at a time from the live queue
//   shade       pathtrace.glsl:223-291   miss/emitter/surface shading, NEE candidates, BSDF sample, RR;
//                                        survivors and shadow requests are compacted with ballot + popc
//   shadow      anyhit.glsl              persistent warps; radiance += (visible NEE sum) * throughput
//   accumulate  renderer.glsl:64-68      per pixel, samples of the batch added in frame order
// plus a one-thread-per-sample megakernel built from the same device functions (cross-check / comparison)
// and the post-process kernel (postprocess.glsl tonemappers) used by the read-back calls.
#include "lf_device.cuh"
#include "lf_shade.cuh"
#include "lf_post.cuh"
#include "lf_kernels.h"

namespace lf {

// slot -> tile-local pixel: 8x4 pixel blocks per warp so that primary rays of a warp stay coherent
LFD bool slot_pixel(const DevParams& P, int q, int& lx, int& ly) {
    int blk = q >> 5, i = q & 31;
    int bw = P.pix_w8 >> 3;
    int bx = blk % bw, by = blk / bw;
    lx = bx * 8 + (i & 7);
    ly = by * 4 + (i >> 3);
    if (lx >= P.tile_w || ly >= P.tile_h) return false;
    int px = P.tile_w * P.tile_x + lx, py = P.tile_h * P.tile_y + ly;   // viewport offset of the tile copy (TiledRenderer.cpp:342)
    return px >= 0 && py >= 0 && px < P.width && py < P.height;
}
// the same slot order for a band of the preview viewport (no tile offset, no screen clipping)
LFD bool slot_preview(const DevParams& P, int q, int& lx, int& ly) {
    int blk = q >> 5, i = q & 31;
    int bw = P.pix_w8 >> 3;
    int bx = blk % bw, by = blk / bw;
    lx = bx * 8 + (i & 7);
    ly = by * 4 + (i >> 3);
    return lx < P.pv_w && ly < P.tile_h && P.pv_y0 + ly < P.pv_h;
}
LFD int pixel_slot(const DevParams& P, int lx, int ly) {
    int bw = P.pix_w8 >> 3;
    return (((ly >> 2) * bw + (lx >> 3)) << 5) + ((ly & 3) << 3) + (lx & 7);
}

// warp-aggregated queue append: one atomicAdd per warp, positions by ballot prefix
LFD void queue_push(int* queue, int* count, bool alive, int value) {
    unsigned m = __ballot_sync(0xffffffffu, alive);
    if (m == 0) return;
    int lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (alive) queue[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// The state of a path that has not bounced yet (pathtrace.glsl:210-216; renderer.glsl:43-62 made its ray and RNG).  `State` is zero-filled,
// so the emission a directly seen analytic light adds "from the last surface" (:246-253) is matID 0's.
LFD void initial_state(const DevScene& S, PathRegs& ps) {
    ps.thr = mk3(1.0f); ps.rad = mk3(0.0f); ps.absn = mk3(0.0f); ps.bsdf_pdf = 0.f;
    ps.stale = xyz(ldg4(S.materials + 1));
}
// Path state at the head of a shade kernel.  Bounce 0 starts from initial_state; `stale` is only ever read on an analytic-light hit, so
// kernels of scenes without lights neither load nor store it.
template <bool LIGHTS>
LFD void load_path(const DevScene& S, const PathSoA& A, int s, int depth, PathRegs& ps) {
    const float4 o = A.ray_o[s], d = A.ray_d[s];
    const uint4 g = A.rng[s];
    ps.ray.o = xyz(o); ps.ray.d = xyz(d);
    ps.rng.x = g.x; ps.rng.y = g.y; ps.rng.z = g.z; ps.rng.w = g.w;
    if (depth == 0) { initial_state(S, ps); return; }
    const float4 th = A.thr[s], ra = A.rad[s], ab = A.absn[s];
    ps.thr = xyz(th); ps.bsdf_pdf = th.w; ps.rad = xyz(ra); ps.absn = xyz(ab);
    ps.stale = LIGHTS ? xyz(A.stale[s]) : mk3(0.f);
}
// LF_FULL_RECORDS = 1 (experiment, off): a kernel that writes one 16-byte half of a 32-byte path-state record it has not read also writes the
// other, unused half (zeros), on the theory that a sector written in part has to be read from DRAM first.  Measured (profiles/r2/r3d_ab_*): it does
// not - k_generate got 25 % SLOWER with the 16 extra bytes per path (C3 19.1 -> 24.2 ms, C4 76 -> 97 ms per 8 steps), k_shade 1.5-3 % faster with
// whole [sh_c0 | sh_c1] and [stale | sh_T] records; in total C2 +0.2 %, C3 -1.5 %, C4 -0.3 %.  Byte-masked sector writes are cheap on this GPU.
#ifndef LF_FULL_RECORDS
#define LF_FULL_RECORDS 0
#endif
// A material whose albedo or metallic / roughness come (partly) from a texture (pathtrace.glsl:82-91; the conditions of load_surface)
LFD bool material_is_textured(const DevScene& S, float texA, float texMR) { return S.num_tex > 0 && ((int)texA >= 0 || (int)texMR >= 0); }
// The shadow request of a surface hit.  A single candidate always goes into slot 0 of the request, whichever kind it is (with one ray the
// order of the sum Li = 0 + c is the same), so that the common request touches two 32-byte records; with both, the environment ray is 0.
LFD void store_nee(const PathSoA& A, int s, const Nee& nee) {
    const bool both = nee.has0 && nee.has1, first1 = !nee.has0;       // first1: the only candidate is the analytic light's
    const f3 d0 = first1 ? nee.d1 : nee.d0, c0 = first1 ? nee.c1 : nee.c0;
    const float m0 = first1 ? nee.m1 : nee.m0;
    A.sh_o[s] = make_float4(nee.origin.x, nee.origin.y, nee.origin.z, __int_as_float(both ? 3 : 1));
    A.sh_d0[s] = make_float4(d0.x, d0.y, d0.z, m0);
    A.sh_c0[s] = make_float4(c0.x, c0.y, c0.z, 0.f);
    if (both) { A.sh_d1[s] = make_float4(nee.d1.x, nee.d1.y, nee.d1.z, nee.m1); A.sh_c1[s] = make_float4(nee.c1.x, nee.c1.y, nee.c1.z, 0.f); }
    else if (LF_FULL_RECORDS) A.sh_c1[s] = make_float4(0.f, 0.f, 0.f, 0.f);                   // the other half of [sh_c0 | sh_c1]
}
// The hit record ClosestHit left in the path state; state.fhp (closest_hit.glsl:139,143) is formed here, by full warps, instead of by the
// few lanes of a traversal warp whose rays happen to end together - and is neither written nor re-read for paths that stop at this hit.
LFD void load_hit(const DevScene& S, const PathSoA& A, int s, const Ray& ray, Hit& h) {
    const float4 hf = A.hit_f[s]; const int4 hi = A.hit_i[s];
    h.t = hf.x; h.u = hf.y; h.v = hf.z; h.lpdf = hf.w; h.tri = hi.x; h.inst = hi.y; h.light = hi.z; h.mat = hi.w;
    hit_point(S, ray, h);
}

// ---------------------------------------------------------------------------------------------- generate
template <bool COUNT>
__global__ void __launch_bounds__(256) k_generate(DevScene S, DevParams P, PathSoA A, Queues Q, DevCounters* cnt) {
    int total = P.num_frames * P.slots_per_frame;                  // multiple of 32
    int* count0 = Q.counts + 0 * Q.stride + 0;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < total; s += gridDim.x * blockDim.x) {
        int fi = s / P.slots_per_frame, q = s - fi * P.slots_per_frame;
        int lx, ly;
        bool valid = P.preview ? slot_preview(P, q, lx, ly) : slot_pixel(P, q, lx, ly);
        if (valid) {
            PathRegs ps;
            if (P.preview) ps.ray = preview_ray(P, lx, P.pv_y0 + ly, ps.rng);
            else ps.ray = camera_ray(P, lx, ly, P.first_frame + fi * P.frame_stride, ps.rng);
            // throughput 1, radiance 0, absorption 0, bsdfSampleRec.pdf 0 and the stale emission are the same for every new path:
            // they are not stored; the shade kernels of bounce 0 start from them (initial_state) instead of loading 64 bytes per path
            A.ray_o[s] = make_float4(ps.ray.o.x, ps.ray.o.y, ps.ray.o.z, 0.f);
            A.ray_d[s] = make_float4(ps.ray.d.x, ps.ray.d.y, ps.ray.d.z, 0.f);
            A.rng[s] = make_uint4(ps.rng.x, ps.rng.y, ps.rng.z, ps.rng.w);
            if (LF_FULL_RECORDS) A.hit_p[s] = make_float4(0.f, 0.f, 0.f, 0.f);      // the other half of [rng | hit_p]; k_shade fills it in when the path goes on
            bump<COUNT>(cnt, C_SAMPLES);
        }
        queue_push(Q.active[0], count0, valid, s);
    }
}

// ---------------------------------------------------------------------------------------------- extend / shadow
// Warp-cooperative persistent traversal, shared by the closest-hit (extend) and any-hit (shadow) stages.
//   * Every lane keeps one ray's walk in registers; the loop body is warp-uniform and alternates two phases:
//     (1) each lane steps through inner nodes / instance entries until it holds a triangle leaf, (2) the leaves are
//     tested together.  The single-loop version ran the leaf code with 2 of 32 lanes active; a speculative variant
//     (postponed leaf, Aila & Laine) was measured too and lost: with distance culling the stale t costs 78 % more node
//     visits (profiles/README.md).
//   * Lanes whose ray is finished are refilled from the queue as soon as kRefillMin of them are idle: one atomicAdd
//     per refill, positions by ballot prefix, so a warp never idles behind its longest ray.
//   * Shadow work item = path slot with up to two NEE rays (env, analytic light), traced one after the other by the
//     same lane; radiance += (visible sum) * throughput is applied when the second is done (pathtrace.glsl:266).
#ifndef LF_TRACE_MINBLOCKS
#define LF_TRACE_MINBLOCKS 9   // 56 registers: 9 CTAs x 128 threads fill the register file (10 -> 48 registers spills, +60 % time)
#endif
#ifndef LF_REFILL_MIN
#define LF_REFILL_MIN 12   // idle lanes that trigger a refill; 8 in round 1, re-swept with the half-parked leaf gather: 12 is +0.8 / +1.4 / +2.0 % on C2 / C4 / C3 (profiles/r2/r2h_ab_*)
#endif
// The inner-node phase of a warp ends when parked lanes * LF_GATHER_DEN >= live lanes * LF_GATHER_NUM (parked = at a triangle leaf, an
// instance entry / exit, or finished).  Round 1 used 1/3; re-swept in round 2 with the final kernels (profiles/r2/r2g_ab_*, r2h_ab_*).
#ifndef LF_STEPS_PER_CHECK
#define LF_STEPS_PER_CHECK 3   // inner steps per phase-end test: 1 -> 2 = +1.3 / +1.3 / +1.2 % on C2 / C4 / C3 (profiles/r2/r2p_ab_*); 3: another +0.6 % on C4, +2 % on C1, C2 unchanged; 4: no better (r2q_ab_*)
#endif
#ifndef LF_GATHER_NUM
#define LF_GATHER_NUM 1
#endif
#ifndef LF_GATHER_DEN
#define LF_GATHER_DEN 2
#endif
#ifndef LF_GATHER_NUM_ANY      // the any-hit (shadow) kernel's threshold, separately tunable
#define LF_GATHER_NUM_ANY LF_GATHER_NUM
#endif
#ifndef LF_GATHER_DEN_ANY
#define LF_GATHER_DEN_ANY LF_GATHER_DEN
#endif
constexpr int kRefillMin = LF_REFILL_MIN;     // idle lanes that trigger a refill from the queue

// Shared memory of a traversal CTA: the stacks [STACK][128] (16 KB at 32 entries) + the world-space rays [9][128] (4.5 KB).  Measured and
// rejected in round 2 (profiles/r2/README.md): keeping only 12 stack entries in shared memory with a global overflow array, and
// re-reading the world ray from the path state instead of parking it here.  Both give the SM's unified array back to the L1 (carve-out
// 196 KB -> 100 KB, L1 hit rate of the incoherent bounces 26 % -> 35 %), but the walk's bound is the L1 data pipe's wavefront rate, not
// its capacity: -1 % at best, and the extra live values push the any-hit kernel into spills at its 56-register limit.
template <bool ANY, bool CULL, bool COUNT, int STACK>
__global__ void __launch_bounds__(kBlockThreads, LF_TRACE_MINBLOCKS) k_trace(DevScene S, PathSoA A, const int* __restrict__ queue, const int* __restrict__ countp,
                                                       int* cursor, Pair<float4> neeT, DevCounters* cnt) {
    __shared__ int stack[STACK * kBlockThreads];
    __shared__ float wray[9 * kBlockThreads];           // world-space ray of each lane + 1/direction (restored when a BLAS is left)
    PlainStk stk;
    stk.col = stack + threadIdx.x;
    float* wr = wray + threadIdx.x;
    const int count = *countp;
    const unsigned lane = threadIdx.x & 31u, ltmask = (1u << lane) - 1u;
    constexpr unsigned FULL = 0xffffffffu;

    bool alive = false, exhausted = false;
    int slot = -1;
    Walk w;
    Hit hit;
    float maxDist = 0.f;
    int shMask = 0, shPhase = 0;                        // shadow: requested rays (bit 0 env, bit 1 light), ray being traced
    f3 Li = mk3(0.0f);
    w.ref = kRefSentinel; w.sp = 0; w.inBlas = false; w.axis = false; w.curInst = -1; w.curMat = 0;
    w.o = w.d = w.idir = mk3(0.f);
    hit_clear(hit);

    // shadow: load NEE ray `phase` of the slot
    auto shadow_ray = [&](int phase) {
        float4 so = A.sh_o[slot];
        float4 dd = phase == 0 ? A.sh_d0[slot] : A.sh_d1[slot];
        Ray r; r.o = xyz(so); r.d = xyz(dd);
        maxDist = dd.w;
        return r;
    };
    auto world_ray = [&]() { Ray r; r.o = mk3(wr[0], wr[kBlockThreads], wr[2 * kBlockThreads]);
                             r.d = mk3(wr[3 * kBlockThreads], wr[4 * kBlockThreads], wr[5 * kBlockThreads]); return r; };
    // start the walk of ray r; returns false when the ray is already decided (ANY: an analytic light blocks it)
    auto begin_ray = [&](const Ray& r) -> bool {
        wr[0] = r.o.x; wr[kBlockThreads] = r.o.y; wr[2 * kBlockThreads] = r.o.z;
        wr[3 * kBlockThreads] = r.d.x; wr[4 * kBlockThreads] = r.d.y; wr[5 * kBlockThreads] = r.d.z;
        if (!ANY) hit_clear(hit);
        bump<COUNT>(cnt, ANY ? C_RAYS_SHADOW : C_RAYS_CLOSEST);
        if (test_lights<ANY, COUNT>(S, r, maxDist, hit, cnt)) return false;
        walk_begin(S, r, w, stk);
        wr[6 * kBlockThreads] = w.idir.x; wr[7 * kBlockThreads] = w.idir.y; wr[8 * kBlockThreads] = w.idir.z;
        return true;
    };

    for (;;) {
        // ---- refill idle lanes
        unsigned idle = __ballot_sync(FULL, !alive);
        if (!exhausted && (idle == FULL || __popc(idle) >= kRefillMin)) {
            int n = __popc(idle), base = 0;
            if (lane == 0) base = atomicAdd(cursor, n);
            base = __shfl_sync(FULL, base, 0);
            if (base + n >= count) exhausted = true;
            int idx = base + __popc(idle & ltmask);
            if (!alive && idx < count) {
                slot = queue[idx];
                alive = true;
                Ray r;
                if (ANY) {
                    shMask = __float_as_int(A.sh_o[slot].w);
                    shPhase = (shMask & 1) ? 0 : 1;
                    Li = mk3(0.0f);
                    r = shadow_ray(shPhase);
                } else {
                    r.o = xyz(A.ray_o[slot]); r.d = xyz(A.ray_d[slot]);
                }
                if (!begin_ray(r)) { w.ref = kRefSentinel; w.inBlas = false; hit.light = 0; }   // decided: occluded by a light
                else if (ANY) hit.light = -1;
            }
        }
        if (!__any_sync(FULL, alive)) break;

        // ---- phase 1: every lane steps through inner nodes; a lane that reaches anything else (triangle leaf, instance
        // entry, end of a BLAS, end of the walk) parks there.  The phase ends when half of the live lanes are parked
        // (or nobody can step): the expensive, rarer steps are then done by many lanes at once, while the inner-node
        // walk never waits for the slowest lane.
        bool rayDone = false;
        const int liveLanes = __popc(__ballot_sync(FULL, alive));
        for (;;) {
            const bool canStep = alive && w.ref >= 0 && !w.axis;
            const unsigned stepMask = __ballot_sync(FULL, canStep);
            if (stepMask == 0u || (liveLanes - __popc(stepMask)) * (ANY ? LF_GATHER_DEN_ANY : LF_GATHER_DEN) >= liveLanes * (ANY ? LF_GATHER_NUM_ANY : LF_GATHER_NUM)) break;
            if (canStep) {
                Ray r;                                       // not read by an inner-node step
                walk_step<ANY, CULL, COUNT, 1>(S, r, w, ANY ? maxDist : hit.t, stk, cnt);
            }
            // the phase-end test above (ballot, popc, compare, branch) is a tenth of an inner step: it runs every LF_STEPS_PER_CHECK steps only
#pragma unroll
            for (int k = 1; k < LF_STEPS_PER_CHECK; k++) {
                if (alive && w.ref >= 0 && !w.axis) {
                    Ray r;
                    walk_step<ANY, CULL, COUNT, 1>(S, r, w, ANY ? maxDist : hit.t, stk, cnt);
                }
            }
        }
        // ---- phase 1x: lanes whose ray is parallel to an axis (1 / d infinite: the slab test takes its NaN-exact form, lf_device.cuh
        // AABBIntersect) park at inner nodes too and take ONE step here per round.  Such rays are vanishingly rare outside hand-built
        // tests, and keeping them out of the loop above keeps that loop free of the test.
        if (alive && w.ref >= 0 && w.axis) {
            Ray r;
            walk_step<ANY, CULL, COUNT, 2>(S, r, w, ANY ? maxDist : hit.t, stk, cnt);
        }
        __syncwarp();
        // ---- phase 2a: instance entries / exits of the parked lanes, together
        if (alive && w.ref < 0 && (w.ref & kRefTlasBit)) {
            if (w.ref == kRefSentinel) {
                if (!w.inBlas) rayDone = true;
                else {                                       // leave the BLAS: world ray and its reciprocal from shared memory
                    w.inBlas = false;
                    Ray r = world_ray();
                    w.o = r.o; w.d = r.d;
                    w.idir = mk3(wr[6 * kBlockThreads], wr[7 * kBlockThreads], wr[8 * kBlockThreads]);
                    w.axis = has_inf(w.idir);
                    w.ref = stk.pop(w.sp);
                }
            } else {
                Ray r = world_ray();
                walk_step<ANY, CULL, COUNT, 1>(S, r, w, ANY ? maxDist : hit.t, stk, cnt);   // an instance entry (the inner-node branch is not taken)
            }
        }
        __syncwarp();
        // ---- phase 2b: the parked leaves' triangles, together
        if (alive && !rayDone && w.ref < 0 && !(w.ref & kRefTlasBit)) {
            if (walk_leaf<ANY, COUNT>(S, w, maxDist, hit, cnt)) { rayDone = true; hit.light = 0; }      // ANY: occluded
            else w.ref = stk.pop(w.sp);
        }
        __syncwarp();
        // ---- finished rays: write the result; shadow lanes go on with their second ray
        if (rayDone) {
            if (!ANY) {
                A.hit_f[slot] = make_float4(hit.t, hit.u, hit.v, hit.lpdf);      // state.fhp is formed by the shade kernel (load_hit)
                A.hit_i[slot] = make_int4(hit.tri, hit.inst, hit.light, hit.mat);
                alive = false;
            } else {
                bool occluded = hit.light == 0;
                if (!occluded) Li = Li + xyz(shPhase == 0 ? A.sh_c0[slot] : A.sh_c1[slot]);
                if (shPhase == 0 && (shMask & 2)) {
                    shPhase = 1;
                    Ray r = shadow_ray(1);
                    if (!begin_ray(r)) { w.ref = kRefSentinel; w.inBlas = false; hit.light = 0; }
                    else hit.light = -1;
                } else {
                    float4 ra = A.rad[slot];
                    f3 rad = xyz(ra) + Li * xyz(neeT[slot]);             // radiance += DirectLight(r, state) * throughput
                    A.rad[slot] = make_float4(rad.x, rad.y, rad.z, 0.f);
                    alive = false;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- shade
// Order of the two passes after k_shade.  1: shadow, then sample: the throughput the NEE sum is multiplied with (pathtrace.glsl:266) is then
// still what k_shade left in PathSoA::thr, so it is neither stored a second time (sh_T) nor fetched from a second array.  0: round 1's order.
// (The fused kernel samples inside the shade kernel and keeps sh_T.)
#ifndef LF_SHADOW_FIRST
#define LF_SHADOW_FIRST 1
#endif
bool shadow_before_sample() { return LF_SHADOW_FIRST != 0; }
#ifndef LF_SHADE_MINBLOCKS
#define LF_SHADE_MINBLOCKS 5   // 96 registers.  ms of k_shade per 6 steps, C2 / C4: 5 CTAs per SM: 77.9 / 440, 6: 82.8 / 479, 7: 83.8 / 529 (profiles/r2/r2a_ab_*;
                               // round 1: 8: 93.9, 9: 97.3, 10: 107.7 on C2, profiles/r1_experiments/ab_shade_sample_ctas.txt): the kernel is latency-bound
                               // and wants registers (no spill traffic on top of its scattered loads), not warps
#endif
// shade, part A: hit processing + next-event estimation.  Surface hits that go on are appended to the sample queue with
// the part of `State` DisneySample needs (5 float4 per path).
template <bool COUNT, bool ENV, bool LIGHTS, bool TEX>
__global__ void __launch_bounds__(128, LF_SHADE_MINBLOCKS) k_shade(DevScene S, DevParams P, PathSoA A, Queues Q, int depth, DevCounters* cnt) {
    const int* queue = (depth & 1) ? Q.active[1] : Q.active[0];   // (a select, not an index: indexing would copy the parameter struct to local memory)
    const int count = Q.counts[0 * Q.stride + depth];
    int* sampleCount = Q.counts + 4 * Q.stride + depth;
    int* shadowCount = Q.counts + 1 * Q.stride + depth;
    const bool lastBounce = depth + 1 >= P.max_depth;          // the BSDF sample of the last bounce cannot reach the image
    const int rounded = (count + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += gridDim.x * blockDim.x) {
        bool wantSample = false, wantShadow = false;
        int s = -1;
        if (i < count) {
            s = queue[i];
            PathRegs ps;
            load_path<LIGHTS>(S, A, s, depth, ps);
            Hit h;
            load_hit(S, A, s, ps.ray, h);
            Nee nee;
            Surf sf;
            f3 absnNext;
            const bool surface = shade_hit<COUNT, ENV, LIGHTS, TEX>(S, P, depth, ps, h, nee, sf, absnNext, cnt);
            wantSample = surface && !lastBounce;
            wantShadow = nee.has0 || nee.has1;
            if (wantShadow) {
                store_nee(A, s, nee);
                if (!LF_SHADOW_FIRST) A.sh_T[s] = make_float4(nee.T.x, nee.T.y, nee.T.z, 0.f);   // else: A.thr, written below, is the same value
            } else if (surface) {
                ps.rad = ps.rad + mk3(0.0f) * nee.T;              // radiance += DirectLight() * throughput with Li == 0 (pathtrace.glsl:266)
            }
            // only what part A changed
            A.thr[s] = make_float4(ps.thr.x, ps.thr.y, ps.thr.z, ps.bsdf_pdf);
            A.rad[s] = make_float4(ps.rad.x, ps.rad.y, ps.rad.z, 0.f);
            if (wantSample) {
                A.absn[s] = make_float4(ps.absn.x, ps.absn.y, ps.absn.z, 0.f);
                if (LIGHTS) {
                    A.stale[s] = make_float4(ps.stale.x, ps.stale.y, ps.stale.z, 0.f);
                    if (LF_FULL_RECORDS && LF_SAMPLE_REMAT && depth == 0) A.sh_T[s] = make_float4(0.f, 0.f, 0.f, 0.f);   // [stale | sh_T]: not read at bounce 0
                }
                A.rng[s] = make_uint4(ps.rng.x, ps.rng.y, ps.rng.z, ps.rng.w);
                const Mat& m = sf.mat;
                A.sf0[s] = make_float4(sf.normal.x, sf.normal.y, sf.normal.z, sf.eta);
#if LF_SAMPLE_REMAT
                // k_sample starts the next ray from the hit point and re-reads the material `h.mat` from the table; what a texture changed
                // of it travels with the path
                A.hit_p[s] = make_float4(h.fhp.x, h.fhp.y, h.fhp.z, __int_as_float(h.mat));
                if (TEX && material_is_textured(S, m.texA, m.texMR)) {
                    A.sf1[s] = make_float4(m.albedo.x, m.albedo.y, m.albedo.z, m.specular);
                    A.sf2[s] = make_float4(m.metallic, m.roughness, m.specularTint, m.sheenTint);
                }
#else
                A.hit_p[s] = make_float4(h.fhp.x, h.fhp.y, h.fhp.z, 0.f);                   // k_sample starts the next ray from it
                A.sf1[s] = make_float4(m.albedo.x, m.albedo.y, m.albedo.z, m.specular);
                A.sf2[s] = make_float4(m.metallic, m.roughness, m.specularTint, m.sheenTint);
                A.sf3[s] = make_float4(m.sheen, m.clearcoat, m.clearcoatRoughness, m.specTrans);
                A.sf4[s] = make_float4(absnNext.x, absnNext.y, absnNext.z, m.subsurface);
#endif
            }
        }
        queue_push(Q.sample, sampleCount, wantSample, s);
        queue_push(Q.shadow, shadowCount, wantShadow, s);
    }
}

// shade, part B: BSDF sample, throughput, Russian roulette, next ray; survivors are appended to the next bounce's queue.
#ifndef LF_SAMPLE_MINBLOCKS
#define LF_SAMPLE_MINBLOCKS 6   // ms of k_sample per 6 C2 steps: 6 CTAs per SM 26.3, 8: 30.4 (profiles/r2/r2d_ab_c2_full_shade4s6.json)
#endif
__global__ void __launch_bounds__(128, LF_SAMPLE_MINBLOCKS) k_sample(DevScene S, DevParams P, PathSoA A, Queues Q, int depth) {
    const int count = Q.counts[4 * Q.stride + depth];
    int* next = (depth & 1) ? Q.active[0] : Q.active[1];
    int* nextCount = Q.counts + 0 * Q.stride + depth + 1;
    const int rounded = (count + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += gridDim.x * blockDim.x) {
        bool alive = false;
        int s = -1;
        if (i < count) {
            s = Q.sample[i];
            PathRegs ps;
            float4 d = A.ray_d[s], th = A.thr[s], ab = A.absn[s], hp = A.hit_p[s];
            uint4 g = A.rng[s];
#if LF_SAMPLE_REMAT
            // the material of the hit, from the table again (GetMaterialsAndTextures, pathtrace.glsl:43-77: the same loads and the same
            // expressions as load_surface / shade_hit, so the same bits); albedo, metallic and roughness from the path state when a texture
            // changed them (pathtrace.glsl:82-98)
            float4 f0 = A.sf0[s], f1, f2, f3v, f4v;
            {
                const float4* mp = S.materials + (size_t)7 * __float_as_int(hp.w);
                const float4 p1 = ldg4(mp), p3 = ldg4(mp + 2), p4 = ldg4(mp + 3), p5 = ldg4(mp + 4), p6 = ldg4(mp + 5), p7 = ldg4(mp + 6);
                f1 = p1;                                                          // albedo.xyz, specular
                f2 = make_float4(p3.x, gmax(p3.y, 0.001f), p3.w, p4.y);           // metallic, roughness, specularTint, sheenTint
                f3v = make_float4(p4.x, p4.z, p4.w, p5.x);                        // sheen, clearcoat, clearcoatRoughness, specTrans
                const f3 an = -mk3(lf_log(p6.x), lf_log(p6.y), lf_log(p6.z)) / p5.z;   // the absorption below the surface (:271-272), as in shade_hit
                f4v = make_float4(an.x, an.y, an.z, p3.z);                        // ..., subsurface
                if (material_is_textured(S, p7.x, p7.y)) { f1 = A.sf1[s]; f2 = A.sf2[s]; }
            }
#else
            float4 f0 = A.sf0[s], f1 = A.sf1[s], f2 = A.sf2[s], f3v = A.sf3[s], f4v = A.sf4[s];
#endif
            ps.ray.d = xyz(d); ps.ray.o = mk3(0.f); ps.thr = xyz(th); ps.bsdf_pdf = th.w; ps.absn = xyz(ab);
            ps.rad = mk3(0.f); ps.stale = mk3(0.f);
            ps.rng.x = g.x; ps.rng.y = g.y; ps.rng.z = g.z; ps.rng.w = g.w;
            Surf sf;
            sf.normal = xyz(f0); sf.eta = f0.w;
            sf.ffnormal = dot(sf.normal, ps.ray.d) <= 0.0f ? sf.normal : sf.normal * -1.0f;   // pathtrace.glsl:34
            Onb(sf.normal, sf.tangent, sf.bitangent);                                            // :36
            Mat& m = sf.mat;
            m.albedo = xyz(f1); m.specular = f1.w;
            m.metallic = f2.x; m.roughness = f2.y; m.specularTint = f2.z; m.sheenTint = f2.w;
            m.sheen = f3v.x; m.clearcoat = f3v.y; m.clearcoatRoughness = f3v.z; m.specTrans = f3v.w;
            m.subsurface = f4v.w;
            alive = shade_sample(P, depth, ps, sf, xyz(hp), xyz(f4v));
            A.thr[s] = make_float4(ps.thr.x, ps.thr.y, ps.thr.z, ps.bsdf_pdf);
            if (alive) {
                A.ray_o[s] = make_float4(ps.ray.o.x, ps.ray.o.y, ps.ray.o.z, 0.f);
                A.ray_d[s] = make_float4(ps.ray.d.x, ps.ray.d.y, ps.ray.d.z, 0.f);
                A.absn[s] = make_float4(ps.absn.x, ps.absn.y, ps.absn.z, 0.f);
                A.rng[s] = make_uint4(ps.rng.x, ps.rng.y, ps.rng.z, ps.rng.w);
            }
        }
        queue_push(next, nextCount, alive, s);
    }
}

// shade, both parts in one kernel: the surface record stays in registers between the NEE half and the BSDF-sample half, so the 5
// float4 of `State` (sf0..sf4) are never written or re-read, and ray_d / thr / absn / hit_p / rng are read once per bounce instead of
// twice: about 300 of the 800 bytes of path state a surviving path moves per bounce.  What it gives up is the compaction between
// the halves: lanes whose path ended in part A (miss, emitter) idle through DisneySample.  Measured (profiles/r2/r2c_ab_*, shade +
// sample ms per 6 steps, split -> fused at 4 CTAs per SM): closed scenes win, C1 7.26 -> 4.28, C4 806 -> 742; scenes whose rays escape
// lose, C2 (env map, half of the bounce-0 rays see the sky) 114 -> 146, C3 (open, textured) 140 -> 215.  With k_shade at 5 CTAs per SM
// the split pair then overtook the fused kernel on C4 as well (697 against 747, profiles/r2/r2d_ab_c4_stress_nofuse.json), while C1's
// 2 M-path batches, where a launch per bounce and its tail are what is saved, keep the gain (831 -> 981 M samples/s).  launch_shade
// therefore fuses only small batches (<= 4 M path slots) of scenes without an environment map and without textures
// (LF_FUSED_SHADE: 0 never, 1 that rule, 2 always).
#ifndef LF_FUSED_MINBLOCKS
#define LF_FUSED_MINBLOCKS 4
#endif
template <bool COUNT, bool ENV, bool LIGHTS, bool TEX>
__global__ void __launch_bounds__(128, LF_FUSED_MINBLOCKS) k_shade_fused(DevScene S, DevParams P, PathSoA A, Queues Q, int depth, DevCounters* cnt) {
    const int* queue = (depth & 1) ? Q.active[1] : Q.active[0];   // (a select, not an index: indexing would copy the parameter struct to local memory)
    const int count = Q.counts[0 * Q.stride + depth];
    int* shadowCount = Q.counts + 1 * Q.stride + depth;
    int* next = (depth & 1) ? Q.active[0] : Q.active[1];
    int* nextCount = Q.counts + 0 * Q.stride + depth + 1;
    const bool lastBounce = depth + 1 >= P.max_depth;          // the BSDF sample of the last bounce cannot reach the image
    const int rounded = (count + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += gridDim.x * blockDim.x) {
        bool alive = false, wantShadow = false;
        int s = -1;
        if (i < count) {
            s = queue[i];
            PathRegs ps;
            load_path<LIGHTS>(S, A, s, depth, ps);
            Hit h;
            load_hit(S, A, s, ps.ray, h);
            Nee nee;
            Surf sf;
            f3 absnNext;
            const bool surface = shade_hit<COUNT, ENV, LIGHTS, TEX>(S, P, depth, ps, h, nee, sf, absnNext, cnt);
            wantShadow = nee.has0 || nee.has1;
            if (wantShadow) {
                store_nee(A, s, nee);
                A.sh_T[s] = make_float4(nee.T.x, nee.T.y, nee.T.z, 0.f);
            } else if (surface) {
                ps.rad = ps.rad + mk3(0.0f) * nee.T;              // radiance += DirectLight() * throughput with Li == 0 (pathtrace.glsl:266)
            }
            A.rad[s] = make_float4(ps.rad.x, ps.rad.y, ps.rad.z, 0.f);
            if (surface && !lastBounce) {
                alive = shade_sample(P, depth, ps, sf, h.fhp, absnNext);
                if (alive) {
                    A.thr[s] = make_float4(ps.thr.x, ps.thr.y, ps.thr.z, ps.bsdf_pdf);
                    A.ray_o[s] = make_float4(ps.ray.o.x, ps.ray.o.y, ps.ray.o.z, 0.f);
                    A.ray_d[s] = make_float4(ps.ray.d.x, ps.ray.d.y, ps.ray.d.z, 0.f);
                    A.absn[s] = make_float4(ps.absn.x, ps.absn.y, ps.absn.z, 0.f);
                    if (LIGHTS) A.stale[s] = make_float4(ps.stale.x, ps.stale.y, ps.stale.z, 0.f);
                    A.rng[s] = make_uint4(ps.rng.x, ps.rng.y, ps.rng.z, ps.rng.w);
                }
            }
        }
        queue_push(next, nextCount, alive, s);
        queue_push(Q.shadow, shadowCount, wantShadow, s);
    }
}

// ---------------------------------------------------------------------------------------------- accumulate
// color = pixelColor + accumColor (renderer.glsl:68), one add per frame of the batch, in frame order.
__global__ void __launch_bounds__(256) k_accumulate(DevParams P, PathSoA A, float* __restrict__ accum) {
    int n = P.tile_w * P.tile_h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int lx = i % P.tile_w, ly = i / P.tile_w;
        int px = P.tile_w * P.tile_x + lx, py = P.tile_h * P.tile_y + ly;
        if (px >= P.width || py >= P.height) continue;
        int q = pixel_slot(P, lx, ly);
        float* a = accum + 3 * ((size_t)py * P.width + px);
        float r = a[0], g = a[1], b = a[2];
        for (int fi = 0; fi < P.num_frames; fi++) {
            float4 c = A.rad[(size_t)fi * P.slots_per_frame + q];
            r = c.x + r; g = c.y + g; b = c.z + b;
        }
        a[0] = r; a[1] = g; a[2] = b;
    }
}

// The preview target is written, not accumulated (preview_flareon.glsl:60; TiledRenderer.cpp:329-331).
__global__ void __launch_bounds__(256) k_preview_store(DevParams P, PathSoA A, float* __restrict__ preview) {
    int n = P.pv_w * P.tile_h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int lx = i % P.pv_w, ly = i / P.pv_w;
        if (P.pv_y0 + ly >= P.pv_h) continue;
        float4 c = A.rad[pixel_slot(P, lx, ly)];
        float* o = preview + 3 * ((size_t)(P.pv_y0 + ly) * P.pv_w + lx);
        o[0] = c.x; o[1] = c.y; o[2] = c.z;
    }
}

// ---------------------------------------------------------------------------------------------- megakernel
// The same device functions, one thread per pixel-sample for the whole path (kernel_mode = 1).
template <bool CULL, bool COUNT, int STACK>
__global__ void __launch_bounds__(kBlockThreads) k_megakernel(DevScene S, DevParams P, PathSoA A, DevCounters* cnt) {
    __shared__ int stack[STACK * kBlockThreads];
    int* stk = stack + threadIdx.x;
    int total = P.num_frames * P.slots_per_frame;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < total; s += gridDim.x * blockDim.x) {
        int fi = s / P.slots_per_frame, q = s - fi * P.slots_per_frame;
        int lx, ly;
        if (!slot_pixel(P, q, lx, ly)) continue;
        PathRegs ps;
        ps.ray = camera_ray(P, lx, ly, P.first_frame + fi * P.frame_stride, ps.rng);
        initial_state(S, ps);
        bump<COUNT>(cnt, C_SAMPLES);
        for (int depth = 0; depth < P.max_depth; depth++) {
            Hit h;
            trace<false, CULL, COUNT>(S, ps.ray, 0.f, h, stk, cnt);
            Nee nee;
            Surf sf;
            f3 absnNext;
            bool go = shade_hit<COUNT, true, true, true>(S, P, depth, ps, h, nee, sf, absnNext, cnt);
            if (go) {
                f3 Li = mk3(0.0f);
                Ray sr; sr.o = nee.origin;
                Hit dummy;
                if (nee.has0) { sr.d = nee.d0; if (!trace<true, CULL, COUNT>(S, sr, nee.m0, dummy, stk, cnt)) Li = Li + nee.c0; }
                if (nee.has1) { sr.d = nee.d1; if (!trace<true, CULL, COUNT>(S, sr, nee.m1, dummy, stk, cnt)) Li = Li + nee.c1; }
                ps.rad = ps.rad + Li * nee.T;
                go = shade_sample(P, depth, ps, sf, h.fhp, absnNext);
            }
            if (!go) break;
        }
        A.rad[s] = make_float4(ps.rad.x, ps.rad.y, ps.rad.z, 0.f);
    }
}

// ---------------------------------------------------------------------------------------------- probes / post
__global__ void k_export_hits(DevScene S, DevParams P, PathSoA A, float* t, int* tri, int* mat, int* emitter) {
    int n = P.tile_w * P.tile_h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int lx = i % P.tile_w, ly = i / P.tile_w;
        int q = pixel_slot(P, lx, ly);
        float4 hf = A.hit_f[q]; int4 hi = A.hit_i[q];
        size_t o = (size_t)ly * P.width + lx;
        t[o] = hf.x;
        tri[o] = (hi.z < 0 && hi.x >= 0) ? __ldg(S.tri_vx + hi.x) : -1;
        mat[o] = (hi.z < 0 && hi.x >= 0) ? hi.w : -1;
        emitter[o] = hi.z >= 0 ? 1 : 0;
    }
}

// postprocess.glsl:26-172 (lf_post.cuh): color = accum * invSampleCounter (or the chromatic-aberration fetches), tonemap, vignette.
__global__ void k_post(const float* __restrict__ accum, float* out_f, unsigned char* out_u8, int W, int H, float inv, int tonemap, LfPostParams pp) {
    const int n = W * H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float c[3];
        post_pixel(accum, W, H, i, inv, tonemap, pp, c);
        if (out_f) { out_f[3 * i] = c[0]; out_f[3 * i + 1] = c[1]; out_f[3 * i + 2] = c[2]; }
        if (out_u8)
            for (int k = 0; k < 3; k++) out_u8[3 * i + k] = (unsigned char)__float2int_rn(clampf(c[k], 0.0f, 1.0f) * 255.0f);   // GL float -> unorm8
    }
}

// The same pass over the accumulation buffers of a multi-GPU group (lf_post.cuh AccumSum): pixel i of the output is the post-processed
// SUM of the devices' buffers.  Launched on the group's first device; the other devices' buffers are read through peer access (NVLink).
// `sum_out` (optional) receives the plain sum, i.e. what lfcuda_read_accum returns for the group.
__global__ void k_post_sum(AccumSum accum, float* out_f, unsigned char* out_u8, float* sum_out, int W, int H, float inv, int tonemap, LfPostParams pp) {
    const int n = W * H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (sum_out) { sum_out[3 * i] = accum[3 * (size_t)i]; sum_out[3 * i + 1] = accum[3 * (size_t)i + 1]; sum_out[3 * i + 2] = accum[3 * (size_t)i + 2]; }
        if (!out_f && !out_u8) continue;
        float c[3];
        post_pixel(accum, W, H, i, inv, tonemap, pp, c);
        if (out_f) { out_f[3 * i] = c[0]; out_f[3 * i + 1] = c[1]; out_f[3 * i + 2] = c[2]; }
        if (out_u8)
            for (int k = 0; k < 3; k++) out_u8[3 * i + k] = (unsigned char)__float2int_rn(clampf(c[k], 0.0f, 1.0f) * 255.0f);
    }
}
void launch_post_sum(cudaStream_t stream, const float* const* accums, int naccum, float* out_f, unsigned char* out_u8, float* sum_out, int W, int H,
                     float inv, int tonemap, const LfPostParams& pp) {
    AccumSum a;
    a.n = naccum;
    for (int k = 0; k < kMaxGroup; k++) a.p[k] = k < naccum ? accums[k] : nullptr;
    k_post_sum<<<(W * H + 255) / 256, 256, 0, stream>>>(a, out_f, out_u8, sum_out, W, H, inv, tonemap, pp);
}

// ---------------------------------------------------------------------------------------------- bandwidth probe
// 16-byte read-only loads (LDG.E.128.CONSTANT, the load the traversal uses) over a buffer, 4 independent loads in
// flight per thread; the xor-sum keeps the loads alive.
__global__ void __launch_bounds__(256) k_read_probe(const float4* __restrict__ buf, size_t n4, int passes, float* sink) {
    float acc = 0.f;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; p++) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n4; i += 4 * stride) {
            float4 a = __ldg(buf + i), b = __ldg(buf + i + stride), c = __ldg(buf + i + 2 * stride), d = __ldg(buf + i + 3 * stride);
            acc += (a.x + b.y) + (c.z + d.w);
        }
        for (; i < n4; i += stride) acc += __ldg(buf + i).x;
    }
    if (acc == 123.456f) *sink = acc;
}
void launch_read_probe(cudaStream_t stream, const float4* buf, size_t n4, int passes, float* sink, int blocks) {
    k_read_probe<<<blocks, 256, 0, stream>>>(buf, n4, passes, sink);
}

// Node-fetch probe (the bound ncu names for the traversal: L1 data-pipe wavefronts).  Every lane walks its own pseudo-random chain
// of 64-byte records fetched exactly like walk_step fetches an inner node (2 x LDG.E.ENL2.256 per record, every lane a different
// record, the next index depending on the loaded data), 8 CTAs x 128 threads per SM, nothing else in the loop.  The rate it reaches
// over a table that fits L2 is the ceiling of one-node-per-lane traversal on this GPU (tools/l1_probe.cu has the other variants).
__global__ void __launch_bounds__(128, 8) k_node_probe(const float4* __restrict__ nodes, unsigned mask, int steps, unsigned* sink) {
    unsigned idx = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    float acc = 0.f;
    for (int s = 0; s < steps; s++) {
        const float4* p = nodes + (size_t)4 * ((idx >> 7) & mask);
        f8 x = ldg8(p), y = ldg8(p + 2);
        acc += ((x.lo.x + x.lo.y) + (x.lo.z + x.lo.w)) + ((x.hi.x + x.hi.y) + (x.hi.z + x.hi.w)) + ((y.lo.x + y.lo.y) + (y.lo.z + y.lo.w)) +
               ((y.hi.x + y.hi.y) + (y.hi.z + y.hi.w));
        idx = idx * 1664525u + 1013904223u + __float_as_uint(y.hi.w);
    }
    if (acc == 123.456f) *sink = 1;
}
void launch_node_probe(cudaStream_t stream, const float4* nodes, unsigned num_nodes_pow2, int steps, unsigned* sink, int blocks) {
    k_node_probe<<<blocks, 128, 0, stream>>>(nodes, num_nodes_pow2 - 1u, steps, sink);
}

// ---------------------------------------------------------------------------------------------- launchers
template <bool CULL, bool COUNT>
static void launch_trace_kernels_s(const LaunchCtx& L, int which, const int* queue, const int* countp, int* cursor) {
    // the throughput of the NEE sum: PathSoA::thr while k_sample has not run yet, the copy in sh_T otherwise
    const Pair<float4> neeT = (LF_SHADOW_FIRST && !shade_is_fused(L)) ? L.soa.thr : L.soa.sh_T;
    if (L.stack_depth <= 32) {
        if (which == 0) k_trace<false, CULL, COUNT, 32><<<L.persistent_blocks, kBlockThreads, 0, L.stream>>>(L.scene, L.soa, queue, countp, cursor, neeT, L.counters);
        else k_trace<true, CULL, COUNT, 32><<<L.persistent_blocks, kBlockThreads, 0, L.stream>>>(L.scene, L.soa, queue, countp, cursor, neeT, L.counters);
    } else {
        if (which == 0) k_trace<false, CULL, COUNT, 64><<<L.persistent_blocks, kBlockThreads, 0, L.stream>>>(L.scene, L.soa, queue, countp, cursor, neeT, L.counters);
        else k_trace<true, CULL, COUNT, 64><<<L.persistent_blocks, kBlockThreads, 0, L.stream>>>(L.scene, L.soa, queue, countp, cursor, neeT, L.counters);
    }
}
static void launch_trace(const LaunchCtx& L, int which, const int* queue, const int* countp, int* cursor) {
    if (L.cull) { if (L.count) launch_trace_kernels_s<true, true>(L, which, queue, countp, cursor); else launch_trace_kernels_s<true, false>(L, which, queue, countp, cursor); }
    else { if (L.count) launch_trace_kernels_s<false, true>(L, which, queue, countp, cursor); else launch_trace_kernels_s<false, false>(L, which, queue, countp, cursor); }
}

void launch_generate(const LaunchCtx& L) {
    int total = L.params.num_frames * L.params.slots_per_frame;
    int blocks = (total + 255) / 256;
    if (blocks > L.sm_count * 8) blocks = L.sm_count * 8;
    if (L.count) k_generate<true><<<blocks, 256, 0, L.stream>>>(L.scene, L.params, L.soa, L.queues, L.counters);
    else k_generate<false><<<blocks, 256, 0, L.stream>>>(L.scene, L.params, L.soa, L.queues, L.counters);
}
// ---- ray-sort experiment (SortCtx in lf_kernels.h): bin = (origin cell, direction octant); histogram, scan, scatter
struct SortGrid { float lo[3], inv[3]; };
template <bool SHADOW>   // SHADOW: the slot's first NEE ray (env if requested, else the analytic light's)
__global__ void __launch_bounds__(256) k_sort_keys(PathSoA A, const int* __restrict__ queue, const int* __restrict__ countp, SortGrid G,
                                                   unsigned* __restrict__ keys, unsigned* hist) {
    const int count = *countp;
    constexpr int C = 1 << kSortCellBits;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const int s = queue[i];
        float4 o, d;
        if (SHADOW) { o = A.sh_o[s]; d = (__float_as_int(o.w) & 1) ? A.sh_d0[s] : A.sh_d1[s]; }
        else { o = A.ray_o[s]; d = A.ray_d[s]; }
        int cx = min(max((int)((o.x - G.lo[0]) * G.inv[0]), 0), C - 1);
        int cy = min(max((int)((o.y - G.lo[1]) * G.inv[1]), 0), C - 1);
        int cz = min(max((int)((o.z - G.lo[2]) * G.inv[2]), 0), C - 1);
        unsigned oct = (d.x < 0.f ? 1u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 4u : 0u);
        unsigned key = ((((unsigned)cz << kSortCellBits | (unsigned)cy) << kSortCellBits | (unsigned)cx) << 3) | oct;
        keys[i] = key;
        atomicAdd(hist + key, 1u);
    }
}
__global__ void __launch_bounds__(1024) k_sort_scan(unsigned* hist) {   // exclusive scan of kSortBins counters, one CTA
    __shared__ unsigned part[1024];
    constexpr int PER = kSortBins / 1024;
    unsigned local[PER], sum = 0;
    for (int k = 0; k < PER; k++) { local[k] = hist[threadIdx.x * PER + k]; sum += local[k]; }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        unsigned v = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = part[threadIdx.x] - sum;
    for (int k = 0; k < PER; k++) { hist[threadIdx.x * PER + k] = run; run += local[k]; }
}
__global__ void __launch_bounds__(256) k_sort_scatter(const int* __restrict__ queue, const int* __restrict__ countp, const unsigned* __restrict__ keys,
                                                      unsigned* offsets, int* __restrict__ out) {
    const int count = *countp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) out[atomicAdd(offsets + keys[i], 1u)] = queue[i];
}

template <bool SHADOW>
static const int* sort_queue(const LaunchCtx& L, const int* queue, const int* countp) {
    const SortCtx& S = *L.sort;
    SortGrid G;
    for (int k = 0; k < 3; k++) { G.lo[k] = S.lo[k]; G.inv[k] = S.inv[k]; }
    cudaMemsetAsync(S.hist, 0, kSortBins * sizeof(unsigned), L.stream);
    k_sort_keys<SHADOW><<<L.sm_count * 8, 256, 0, L.stream>>>(L.soa, queue, countp, G, S.keys, S.hist);
    k_sort_scan<<<1, 1024, 0, L.stream>>>(S.hist);
    k_sort_scatter<<<L.sm_count * 8, 256, 0, L.stream>>>(queue, countp, S.keys, S.hist, S.sorted);
    return S.sorted;
}

void launch_extend(const LaunchCtx& L, int depth) {
    const Queues& Q = L.queues;
    const int* queue = Q.active[depth & 1];
    const int* countp = Q.counts + 0 * Q.stride + depth;
    if (L.sort && (L.sort->mode & 1) && depth >= 1) queue = sort_queue<false>(L, queue, countp);   // primary rays are coherent already (8x4 pixel blocks per warp)
    launch_trace(L, 0, queue, countp, Q.counts + 2 * Q.stride + depth);
}
#ifndef LF_FUSED_SHADE
#define LF_FUSED_SHADE 1
#endif
// does launch_shade run both halves in one kernel for this scene?  (then launch_sample is a no-op)
bool shade_is_fused(const LaunchCtx& L) {
    if (LF_FUSED_SHADE == 2) return true;
    if (LF_FUSED_SHADE == 0 || L.count) return false;
    return L.params.use_envmap == 0 && L.scene.num_tex == 0 && (long long)L.params.num_frames * L.params.slots_per_frame <= (4ll << 20);
}
template <bool ENV, bool LIGHTS, bool TEX>
static void launch_shade_v(const LaunchCtx& L, int depth, bool fused) {
    if (fused) k_shade_fused<false, ENV, LIGHTS, TEX><<<L.sm_count * LF_FUSED_MINBLOCKS, 128, 0, L.stream>>>(L.scene, L.params, L.soa, L.queues, depth, L.counters);
    else k_shade<false, ENV, LIGHTS, TEX><<<L.sm_count * LF_SHADE_MINBLOCKS, 128, 0, L.stream>>>(L.scene, L.params, L.soa, L.queues, depth, L.counters);
}
void launch_shade(const LaunchCtx& L, int depth) {
    // grids are exactly one resident wave (more CTAs than fit measured 25 % slower)
    if (L.count) { k_shade<true, true, true, true><<<L.sm_count * LF_SHADE_MINBLOCKS, 128, 0, L.stream>>>(L.scene, L.params, L.soa, L.queues, depth, L.counters); return; }
    const bool env = L.params.use_envmap != 0, lights = L.scene.num_lights > 0, tex = L.scene.num_tex > 0;
    const bool fused = shade_is_fused(L);
    switch ((env ? 4 : 0) | (lights ? 2 : 0) | (tex ? 1 : 0)) {
        case 0: launch_shade_v<false, false, false>(L, depth, fused); break;
        case 1: launch_shade_v<false, false, true>(L, depth, fused); break;
        case 2: launch_shade_v<false, true, false>(L, depth, fused); break;
        case 3: launch_shade_v<false, true, true>(L, depth, fused); break;
        case 4: launch_shade_v<true, false, false>(L, depth, fused); break;
        case 5: launch_shade_v<true, false, true>(L, depth, fused); break;
        case 6: launch_shade_v<true, true, false>(L, depth, fused); break;
        default: launch_shade_v<true, true, true>(L, depth, fused); break;
    }
}
void launch_sample(const LaunchCtx& L, int depth) {
    if (shade_is_fused(L)) return;                     // the BSDF sample ran inside the shade kernel
    k_sample<<<L.sm_count * LF_SAMPLE_MINBLOCKS, 128, 0, L.stream>>>(L.scene, L.params, L.soa, L.queues, depth);
}
void launch_shadow(const LaunchCtx& L, int depth) {
    const Queues& Q = L.queues;
    const int* queue = Q.shadow;
    const int* countp = Q.counts + 1 * Q.stride + depth;
    if (L.sort && (L.sort->mode & 2)) queue = sort_queue<true>(L, queue, countp);
    launch_trace(L, 1, queue, countp, Q.counts + 3 * Q.stride + depth);
}
void launch_accumulate(const LaunchCtx& L, float* accum) {
    int n = L.params.tile_w * L.params.tile_h;
    int blocks = (n + 255) / 256;
    if (blocks > L.sm_count * 8) blocks = L.sm_count * 8;
    k_accumulate<<<blocks, 256, 0, L.stream>>>(L.params, L.soa, accum);
}
void launch_preview_store(const LaunchCtx& L, float* preview) {
    int n = L.params.pv_w * L.params.tile_h;
    int blocks = (n + 255) / 256;
    if (blocks > L.sm_count * 8) blocks = L.sm_count * 8;
    k_preview_store<<<blocks, 256, 0, L.stream>>>(L.params, L.soa, preview);
}
template <bool CULL, bool COUNT>
static void launch_mega_s(const LaunchCtx& L, int blocks) {
    if (L.stack_depth <= 32) k_megakernel<CULL, COUNT, 32><<<blocks, kBlockThreads, 0, L.stream>>>(L.scene, L.params, L.soa, L.counters);
    else k_megakernel<CULL, COUNT, 64><<<blocks, kBlockThreads, 0, L.stream>>>(L.scene, L.params, L.soa, L.counters);
}
void launch_megakernel(const LaunchCtx& L) {
    int total = L.params.num_frames * L.params.slots_per_frame;
    int blocks = (total + kBlockThreads - 1) / kBlockThreads;
    if (L.cull) { if (L.count) launch_mega_s<true, true>(L, blocks); else launch_mega_s<true, false>(L, blocks); }
    else { if (L.count) launch_mega_s<false, true>(L, blocks); else launch_mega_s<false, false>(L, blocks); }
}
void launch_export_hits(const LaunchCtx& L, float* t, int* tri, int* mat, int* emitter) {
    int n = L.params.tile_w * L.params.tile_h;
    k_export_hits<<<(n + 255) / 256, 256, 0, L.stream>>>(L.scene, L.params, L.soa, t, tri, mat, emitter);
}
void launch_post(cudaStream_t stream, const float* accum, float* out_f, unsigned char* out_u8, int W, int H, float inv, int tonemap, const LfPostParams& pp) {
    k_post<<<(W * H + 255) / 256, 256, 0, stream>>>(accum, out_f, out_u8, W, H, inv, tonemap, pp);
}

}  // namespace lf
