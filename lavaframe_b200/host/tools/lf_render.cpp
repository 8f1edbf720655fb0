// lf_render — headless driver: what LavaFrame/Main.cpp does with `-s scene -ms N -u -w`, minus SDL/ImGui/GL:
// load the scene with the reference's loader, construct the renderer, loop Update -> Render until
// `maxSamples + 1 == GetSampleCount()` (Main.cpp:197), then fetch GetOutputBufferHDR.
//
//   lf_render <scene> --spp N [--out img.f32] [--exr f.exr] [--png f.png] [--bmp f.bmp] [--tga f.tga] [--jpg f.jpg] [--device D | --gpus N | --devices a,b,..] [--tonemap]
// --gpus N / --devices: the ONE renderer Main.cpp constructs drives several GPUs of this process (CudaRenderer(scene, dir, devices)).
// img.f32: W*H*3 float32, rows bottom-up (the exporters flip, Export.h:19).  --png / --bmp / --tga / --jpg do what SaveFrame,
// SaveFrameBMP, SaveFrameTGA and SaveFrameJPG do (LavaFrame/Export.h:14-57): GetOutputBuffer -> stbi_flip_vertically_on_write ->
// stbi_write_*, with the reference's own stb_image_write.h compiled from where it lies.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "stb_image_write.h"

#include "Scene.h"
#include "Loader.h"
#include "GlobalState.h"
#include "../CudaRenderer.h"

using namespace LavaFrame;
extern LavaFrameState GlobalState;

// SaveFrameEXR (LavaFrame/Export.h:57-152): GetOutputBufferHDR split into three FLOAT channels named B, G, R, rows in the order the
// renderer returns them (bottom row first: the reference does not flip for EXR).  The reference hands the planes to tinyexr, which is
// not in its tree (<tinyexr.h> is an un-vendored dependency); this writes the same image as a plain OpenEXR 2 scan-line file without
// compression (the reference's exrCompressionIndex picks ZIP / RLE / ZIPS / PIZ / ZFP through tinyexr; the pixel payload is the same).
static bool write_exr_bgr(const char* path, const float* rgb, int w, int h) {
    FILE* f = fopen(path, "wb");
    if (!f) return false;
    std::vector<unsigned char> hd;
    auto bytes = [&](const void* p, size_t n) { hd.insert(hd.end(), (const unsigned char*)p, (const unsigned char*)p + n); };
    auto str = [&](const char* s) { bytes(s, strlen(s) + 1); };
    auto i32 = [&](int v) { bytes(&v, 4); };
    auto f32 = [&](float v) { bytes(&v, 4); };
    const unsigned char magic[8] = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
    bytes(magic, 8);
    str("channels"); str("chlist"); i32(3 * (2 + 4 + 4 + 4 + 4) + 1);
    for (const char* c : {"B", "G", "R"}) { str(c); i32(2 /* FLOAT */); i32(0); i32(1); i32(1); }
    hd.push_back(0);
    str("compression"); str("compression"); i32(1); hd.push_back(0);
    str("dataWindow"); str("box2i"); i32(16); i32(0); i32(0); i32(w - 1); i32(h - 1);
    str("displayWindow"); str("box2i"); i32(16); i32(0); i32(0); i32(w - 1); i32(h - 1);
    str("lineOrder"); str("lineOrder"); i32(1); hd.push_back(0);
    str("pixelAspectRatio"); str("float"); i32(4); f32(1.0f);
    str("screenWindowCenter"); str("v2f"); i32(8); f32(0.f); f32(0.f);
    str("screenWindowWidth"); str("float"); i32(4); f32(1.0f);
    hd.push_back(0);
    fwrite(hd.data(), 1, hd.size(), f);
    const size_t row = 8 + (size_t)3 * w * 4;
    unsigned long long off = hd.size() + (size_t)8 * h;
    for (int y = 0; y < h; y++, off += row) fwrite(&off, 8, 1, f);
    std::vector<float> plane((size_t)3 * w);
    for (int y = 0; y < h; y++) {
        const int size = 3 * w * 4;
        fwrite(&y, 4, 1, f); fwrite(&size, 4, 1, f);
        for (int x = 0; x < w; x++) {
            const float* p = rgb + 3 * ((size_t)y * w + x);
            plane[x] = p[2]; plane[w + x] = p[1]; plane[2 * (size_t)w + x] = p[0];   // B, G, R planes
        }
        fwrite(plane.data(), 4, plane.size(), f);
    }
    fclose(f);
    return true;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: lf_render <scene> --spp N [--out img.f32] [--device D] [--tonemap]\n"); return 2; }
    std::string out, png, bmp, tga, jpg, exr;
    int spp = 1;
    std::vector<int> devices(1, 0);
    bool keepTonemap = false;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--spp") spp = atoi(argv[++i]);
        else if (a == "--out") out = argv[++i];
        else if (a == "--png") png = argv[++i];
        else if (a == "--bmp") bmp = argv[++i];
        else if (a == "--tga") tga = argv[++i];
        else if (a == "--jpg") jpg = argv[++i];
        else if (a == "--exr") exr = argv[++i];
        else if (a == "--device") devices.assign(1, atoi(argv[++i]));
        else if (a == "--gpus") { int n = atoi(argv[++i]); devices.clear(); for (int d = 0; d < n; d++) devices.push_back(d); }
        else if (a == "--devices") { devices.clear(); for (char* t = strtok(argv[++i], ","); t; t = strtok(nullptr, ",")) devices.push_back(atoi(t)); }
        else if (a == "--tonemap") keepTonemap = true;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    RenderOptions ro;
    ro.useVignette = false; ro.vignetteIntensity = 0.f; ro.vignettePower = 1.f;
    if (!keepTonemap) ro.tonemapIndex = 0;
    GlobalState.scene = new Scene();
    bool loaded = false;
    try { loaded = LoadSceneFromFile(argv[1], GlobalState.scene, ro); }
    catch (const std::exception& e) { fprintf(stderr, "lf_render: %s\n", e.what()); }         // LF_DEVICE_BLAS=1 without a usable GPU
    if (!loaded) return 1;
    if (!keepTonemap) ro.tonemapIndex = 0;
    GlobalState.scene->renderOptions = ro;
    GlobalState.scene->camera->isMoving = false;

    if (devices.empty()) { fprintf(stderr, "lf_render: no devices\n"); return 2; }
    CudaRenderer* r = new CudaRenderer(GlobalState.scene, GlobalState.shadersDir, devices);   // Main.cpp:91
    GlobalState.renderer = r;
    r->Init();
    if (!r->Ok()) { fprintf(stderr, "lf_render: %s\n", r->LastError()); return 3; }
    auto t0 = std::chrono::steady_clock::now();
    int steps = 0;
    while (true) {
        if (r->GetSampleCount() == spp + 1) break;
        r->Update(0.f);
        if (r->GetSampleCount() == spp + 1) break;
        r->Render();
        steps++;
    }
    float* img = nullptr;
    int w = 0, h = 0;
    r->GetOutputBufferHDR(&img, w, h);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double sum[3] = {0, 0, 0};
    for (long i = 0; i < (long)w * h; i++) for (int k = 0; k < 3; k++) sum[k] += img[3 * i + k];
    printf("{\"impl\": \"cuda\", \"gpus\": %d, \"width\": %d, \"height\": %d, \"spp\": %d, \"tile_steps\": %d, \"seconds\": %.4f, \"samples_per_s\": %.1f, "
           "\"mean_rgb\": [%.8g, %.8g, %.8g]}\n", (int)devices.size(), w, h, spp, steps, sec, (double)w * h * spp / sec, sum[0] / (w * h), sum[1] / (w * h), sum[2] / (w * h));
    if (!out.empty()) { FILE* f = fopen(out.c_str(), "wb"); fwrite(img, 4, (size_t)w * h * 3, f); fclose(f); }
    if (!exr.empty() && !write_exr_bgr(exr.c_str(), img, w, h)) fprintf(stderr, "lf_render: cannot write %s\n", exr.c_str());
    delete[] img;
    if (!png.empty() || !bmp.empty() || !tga.empty() || !jpg.empty()) {   // Export.h:14-57
        unsigned char* data = nullptr;
        int ew = 0, eh = 0;
        r->GetOutputBuffer(&data, ew, eh);
        stbi_flip_vertically_on_write(true);
        if (!png.empty()) stbi_write_png(png.c_str(), ew, eh, 3, data, ew * 3);
        if (!bmp.empty()) stbi_write_bmp(bmp.c_str(), ew, eh, 3, data);
        if (!tga.empty()) stbi_write_tga(tga.c_str(), ew, eh, 3, data);
        if (!jpg.empty()) stbi_write_jpg(jpg.c_str(), ew, eh, 3, data, GlobalState.currentJpgQuality);
        delete[] data;
    }
    delete r;
    return 0;
}
