// lf_shade.cuh — one iteration of the bounce loop after ClosestHit (pathtrace.glsl:223-291) as device functions shared by
// the wavefront kernels (k_shade, k_sample), the megakernel, and tests/hostcheck (the same text compiled for the host).
#pragma once

#include "lf_device.cuh"

namespace lf {

// ---------------------------------------------------------------------------------------------- path state
struct PathRegs {
    Ray ray;
    f3 thr, rad, absn, stale;
    float bsdf_pdf;
    Rng rng;
};
struct Nee {
    f3 origin, d0, c0, d1, c1, T;
    float m0, m1;
    bool has0, has1;
};

// One iteration of the bounce loop after ClosestHit (pathtrace.glsl:223-291), in two parts that the wavefront runs as
// two kernels (one kernel holding both was instruction-cache and register bound: the two halves cost 0.86 + 0.95 ms
// apart and 3.07 ms together on C2's first bounce):
//   shade_hit     :223-266  miss / emitter / surface fetch, emission, absorption, NEE candidates (already weighted,
//                           visibility pending) -> `nee`; returns true when a surface was hit and `s` is valid
//   shade_sample  :268-291  DisneySample, throughput, Russian roulette, next ray; returns true when the path continues
// ENV / LIGHTS / TEX mirror the reference's shader variants (#define ENVMAP, LIGHTS; a bound texture array): a scene
// without an env map, analytic lights or textures runs a kernel that does not contain that code at all.
template <bool COUNT, bool ENV = true, bool LIGHTS = true, bool TEX = true>
LFD bool shade_hit(const DevScene& S, const DevParams& P, int depth, PathRegs& ps, const Hit& hit, Nee& nee, Surf& s, f3& absnNext, DevCounters* cnt) {
    nee.has0 = nee.has1 = false;
    const float t = hit.t;
    const f3 rd = ps.ray.d;

    if (t == kINF) {   // pathtrace.glsl:223-244
        if (P.use_constant_bg) {
            ps.rad = ps.rad + mk3(P.bg[0], P.bg[1], P.bg[2]) * ps.thr;
        } else if (ENV && P.use_envmap) {
            float misWeight = 1.0f;
            float ux = (kPI + lf_atan2(rd.z, rd.x)) * (1.0f / kTWO_PI), uy = lf_acos(rd.y) * (1.0f / kPI);
            if (depth > 0) {
                float lightPdf = EnvPdf(S, P, rd);
                misWeight = powerHeuristic(ps.bsdf_pdf, lightPdf);
            }
            bump<COUNT>(cnt, C_ENV_MISS);
            ps.rad = ps.rad + misWeight * hdrLinear(S, ux, uy) * ps.thr * P.hdr_multiplier;
        }
        return false;
    }

    if (LIGHTS && hit.light >= 0) {   // analytic light is the nearest hit (pathtrace.glsl:246-261 with the stale State)
        ps.rad = ps.rad + ps.stale * ps.thr;
        LightRec L = load_light(S, hit.light);
        f3 Le = L.emission;
        if (depth != 0) Le = powerHeuristic(ps.bsdf_pdf, hit.lpdf) * L.emission;   // EmitterSample, sampling.glsl:271-282
        ps.rad = ps.rad + Le * ps.thr;
        return false;
    }

    load_surface<COUNT, TEX>(S, hit, rd, s, cnt);
    ps.stale = s.mat.emission;

    if (dot(s.normal, s.ffnormal) > 0.0f) ps.absn = mk3(0.0f);   // :250-251
    ps.rad = ps.rad + s.mat.emission * ps.thr;                     // :253
    {
        f3 a = -ps.absn * t;                                      // :264
        ps.thr = ps.thr * mk3(lf_exp(a.x), lf_exp(a.y), lf_exp(a.z));
    }

    // ---- DirectLight (pathtrace.glsl:126-204): weighted candidates now, visibility later
    const f3 V = -rd;
    f3 surfacePos = hit.fhp + s.normal * kEPS;
    nee.origin = surfacePos;
    nee.T = ps.thr;
    if (ENV && P.use_envmap && !P.use_constant_bg) {
        f3 color;
        bump<COUNT>(cnt, C_ENV_NEE);
        float4 dirPdf = EnvSample(S, P, ps.rng, color);
        f3 lightDir = mk3(dirPdf.x, dirPdf.y, dirPdf.z);
        float lightPdf = dirPdf.w;
        float pdf;
        f3 f = DisneyEval(s, V, s.ffnormal, lightDir, pdf);
        if (pdf > 0.0f) {
            float misWeight = powerHeuristic(lightPdf, pdf);
            if (misWeight > 0.0f) {
                nee.c0 = misWeight * f * fabsf(dot(lightDir, s.ffnormal)) * color / lightPdf;
                nee.d0 = lightDir;
                nee.m0 = kINF - kEPS;
                nee.has0 = true;
            }
        }
    }
    if (LIGHTS && S.num_lights > 0) {
        int index = (int)(rnd(ps.rng) * (float)S.num_lights);
        LightRec light = load_light(S, index);
        LightSample ls;
        sampleOneLight(light, S.num_lights, surfacePos, ps.rng, ls);
        if (dot(ls.direction, ls.normal) < 0.0f) {
            float pdf;
            f3 f = DisneyEval(s, V, s.ffnormal, ls.direction, pdf);
            float weight = 1.0f;
            if (light.area > 0.0f) weight = powerHeuristic(ls.pdf, pdf);
            if (pdf > 0.0f) {
                nee.c1 = weight * f * fabsf(dot(s.ffnormal, ls.direction)) * ls.emission / ls.pdf;
                nee.d1 = ls.direction;
                nee.m1 = ls.dist - kEPS;
                nee.has1 = true;
            }
        }
    }

    {   // the absorption the path takes on if the sampled direction goes below the surface (:271-272)
        f3 e = s.mat.extinction;
        absnNext = -mk3(lf_log(e.x), lf_log(e.y), lf_log(e.z)) / s.mat.atDistance;
    }
    return true;
}

// BSDF sample (pathtrace.glsl:268-291).  `s` needs normal, ffnormal, tangent frame, eta and the material.
LFD bool shade_sample(const DevParams& P, int depth, PathRegs& ps, const Surf& s, f3 fhp, f3 absnNext) {
    const f3 V = -ps.ray.d;
    f3 L;
    float pdf;
    f3 f = DisneySample(s, V, s.ffnormal, ps.rng, L, pdf);
    ps.bsdf_pdf = pdf;
    if (dot(s.ffnormal, L) < 0.0f) ps.absn = absnNext;
    if (pdf > 0.0f) ps.thr = ps.thr * (f * fabsf(dot(s.ffnormal, L)) / pdf);
    else return false;

    if (P.enable_rr && depth >= P.rr_depth) {
        float q = gmin(gmax(ps.thr.x, gmax(ps.thr.y, ps.thr.z)) + 0.001f, 0.95f);
        if (rnd(ps.rng) > q) return false;
        ps.thr = ps.thr / q;
    }
    ps.ray.d = L;
    ps.ray.o = fhp + L * kEPS;
    return true;
}

}  // namespace lf
