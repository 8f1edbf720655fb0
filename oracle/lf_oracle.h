// lf_oracle.h — TEST INFRASTRUCTURE.  CPU restatement (tier-2 oracle) of LavaFrame's path-tracing
// fragment shader: shaders/renderer.glsl + shaders/common/*.glsl of the reference.
//
// This is a checker, not a product path: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, load or run it.  Parity status: PINNED against the
// reference itself — the unmodified GLSL executed on Mesa llvmpipe (oracle/_ref/lf_ref_llvmpipe) — on
// primary-hit IDs/t, 1-spp radiance and 64-spp means of the Cornell box and the generated scenes; the
// fixtures and the script that made them are under tests/golden/.
//
// Arithmetic: IEEE fp32, no FMA contraction (-ffp-contract=off), uint32 RNG; every function cites the
// GLSL lines it restates.
#pragma once

#include <cstdint>
#include <vector>

#include "lfcuda.h"   // LfSceneView / LfParams / LfCamera / LfCounters: the shared input structs

namespace lforacle {

struct Oracle {
    LfSceneView scene;
    LfParams    params;
    LfCamera    camera;
    bool        cull = false;     // false = visit every pierced box like the reference (closest_hit.glsl:169-198)
    bool        count = false;    // maintain counters (serialises nothing: per-thread then summed)
    LfCounters  counters;

    Oracle(const LfSceneView& s, const LfParams& p, const LfCamera& c);

    // One pixel-sample = one GL fragment of renderer.glsl:25-69 for tile-local pixel (lx, ly) of tile
    // (tileX, tileY) with uniform `frame`.  Returns PathTrace(ray) (not yet added to the accumulator) and the
    // full-frame pixel it lands on (-1,-1 if clipped).
    void Sample(int lx, int ly, int tileX, int tileY, int frame, float rgb[3], int* px, int* py, LfCounters* cnt) const;

    // nframes draws of one tile, each added to accum (W*H*3, rows bottom-up) in frame order; OpenMP over pixels.
    void RenderFrames(int firstFrame, int nframes, int frameStride, int tileX, int tileY, float* accum);

    // The preview engine (shaders/preview_flareon.glsl:22-61): one sample per pixel of a pvW x pvH viewport, frame 1,
    // no accumulation.  out = pvW * pvH * 3 floats, rows bottom-up.
    void RenderPreview(int pvW, int pvH, int maxDepth, bool useDof, float* out);

    // Probe 1: ClosestHit of the first camera ray of `frame` for every pixel of the full frame (single tile
    // covering the frame), like the llvmpipe "--probe hits" shader variant.
    void PrimaryHits(int frame, float* t, int32_t* triX, int32_t* matID, int32_t* emitter);

    // rand() known-answer probe: the first n draws after InitRNG((px+.5, py+.5), frame) (globals.glsl:116-133).
    static void RandKat(int px, int py, int frame, int n, uint32_t* seedx, float* values);

    // The post-process pass (shaders/postprocess.glsl:126-172, drawn by TiledRenderer::Render into tileOutputTexture,
    // TiledRenderer.cpp:346-350): out = tonemap(accum * inv) with optional chromatic aberration and vignette.
    // accum / out: W*H*3 floats, rows bottom-up.
    static void PostProcess(const float* accum, int W, int H, float inv, int tonemapIndex, const LfPostParams& pp, float* out);

    // GLSL built-in known-answer probe (tests/golden/llvmpipe_builtins.npz, made by executing the same expressions on
    // llvmpipe): n vec4 arguments -> n vec4 results of expression group `op` (see the switch in lf_oracle.cpp).
    // `tex` (W x H x L RGBA8, for the texture-filter group) may be null otherwise.
    static void BuiltinKat(int op, const float* in4, int n, float* out4, const uint8_t* tex, int texW, int texH, int texL);

    // BSDF known-answer probe (tests/golden/llvmpipe_bsdf.npz, made by executing the reference's OWN disney.glsl / sampling.glsl
    // functions on llvmpipe): n items of 9 vec4 (V, N, L, material, frame, RNG seed; layout in lf_oracle.cpp) -> n vec4.
    // op 0: DisneyEval -> (f, pdf); 1: DisneySample -> (L, pdf); 2: DisneySample -> (f, rand() after the call);
    // 3: (GTR1, GTR2, SmithG_GGX, DielectricFresnel) of (a.x, a.y); 4: ImportanceSampleGTR1 / GTR2 / CosineSampleHemisphere.
    static void BsdfKat(int op, const float* in, int n, float* out4);
};

}  // namespace lforacle
