// lf_oracle — TEST INFRASTRUCTURE: command-line front end of the CPU oracle.
//   lf_oracle <scene.lfpack> --spp N [--first-frame F] [--out img.f32] [--hits hits.f32] [--cull] [--count]
// img.f32 = accumulated sum / spp (W*H*3 float32, rows bottom-up), i.e. GetOutputBufferHDR with tonemapIndex 0.
// hits.f32 = (t, triID.x [+0.5 when an analytic light is nearest], matID) like lf_ref_llvmpipe --probe hits.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <omp.h>
#include <string>
#include <vector>

#include "lf_oracle.h"
#include "scenepack.h"

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: lf_oracle <scene.lfpack> --spp N [--out img.f32] [--hits hits.f32] [--cull] [--count]\n"); return 2; }
    std::string out, hits;
    int spp = 1, first = 2;
    bool cull = false, count = false;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--spp") spp = atoi(argv[++i]);
        else if (a == "--first-frame") first = atoi(argv[++i]);
        else if (a == "--out") out = argv[++i];
        else if (a == "--hits") hits = argv[++i];
        else if (a == "--cull") cull = true;
        else if (a == "--count") count = true;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    lfpack::ScenePack pack;
    std::string err;
    if (!lfpack::read(argv[1], pack, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    lforacle::Oracle o(pack.view(), pack.params(), pack.camera());
    o.cull = cull; o.count = count;
    const int W = o.params.width, H = o.params.height;
    if (!hits.empty()) {
        std::vector<float> t(W * H); std::vector<int32_t> tri(W * H), mat(W * H), em(W * H);
        o.PrimaryHits(first, t.data(), tri.data(), mat.data(), em.data());
        std::vector<float> img((size_t)W * H * 3);
        for (int i = 0; i < W * H; i++) { img[3 * i] = t[i]; img[3 * i + 1] = (float)tri[i] + (em[i] ? 0.5f : 0.f); img[3 * i + 2] = (float)mat[i]; }
        FILE* f = fopen(hits.c_str(), "wb"); fwrite(img.data(), 4, img.size(), f); fclose(f);
    }
    std::vector<float> accum((size_t)W * H * 3, 0.f);
    int tilesX = (W + o.params.tile_width - 1) / o.params.tile_width, tilesY = (H + o.params.tile_height - 1) / o.params.tile_height;
    auto t0 = std::chrono::steady_clock::now();
    int frame = first;
    for (int s = 0; s < spp; s++)
        for (int ty = tilesY - 1; ty >= 0; ty--)          // tile walk of TiledRenderer::Update (TiledRenderer.cpp:485-501)
            for (int tx = 0; tx < tilesX; tx++) o.RenderFrames(frame++, 1, 1, tx, ty, accum.data());
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double sum[3] = {0, 0, 0};
    float inv = 1.0f / (float)spp;
    for (size_t i = 0; i < accum.size(); i++) { accum[i] *= inv; sum[i % 3] += accum[i]; }
    if (!out.empty()) { FILE* f = fopen(out.c_str(), "wb"); fwrite(accum.data(), 4, accum.size(), f); fclose(f); }
    printf("{\"impl\": \"oracle\", \"width\": %d, \"height\": %d, \"spp\": %d, \"seconds\": %.3f, \"samples_per_s\": %.1f, \"threads\": %d, "
           "\"mean_rgb\": [%.8g, %.8g, %.8g]", W, H, spp, sec, (double)W * H * spp / sec, omp_get_max_threads(),
           sum[0] / (W * H), sum[1] / (W * H), sum[2] / (W * H));
    if (count) {
        const LfCounters& c = o.counters;
        printf(", \"counters\": {\"samples\": %llu, \"rays_closest\": %llu, \"rays_shadow\": %llu, \"inner\": %llu, \"leaf\": %llu, \"tri\": %llu, "
               "\"tlas\": %llu, \"light_tests\": %llu, \"shaded\": %llu, \"env_nee\": %llu, \"env_miss\": %llu, \"tex\": %llu}",
               (unsigned long long)c.samples, (unsigned long long)c.rays_closest, (unsigned long long)c.rays_shadow, (unsigned long long)c.inner_visits,
               (unsigned long long)c.leaf_visits, (unsigned long long)c.tri_tests, (unsigned long long)c.tlas_visits, (unsigned long long)c.light_tests,
               (unsigned long long)c.shaded_hits, (unsigned long long)c.env_nee, (unsigned long long)c.env_miss, (unsigned long long)c.tex_samples);
    }
    printf("}\n");
    return 0;
}
