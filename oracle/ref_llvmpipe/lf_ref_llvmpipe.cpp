// lf_ref_llvmpipe — TEST INFRASTRUCTURE (tier-1 oracle): runs the reference's UNMODIFIED renderer
// (LavaFrame/TiledRenderer.cpp, Renderer.cpp, Shader.cpp, Program.cpp, Quad.cpp, gl3w.c and the GLSL under
// shaders/) on the host CPU with Mesa llvmpipe, headless.  Nothing here is product code and nothing in the
// product links it.  The reference sources are compiled where they lie under /root/reference by
// oracle/Makefile; the GLSL text and the reference's assets are packed into this binary at build time
// (tar blob) so that it also runs on a machine where /root/reference does not exist.
//
// The driver loop mirrors LavaFrame/Main.cpp:313-755 (MainLoop: Update -> Render) and the auto-stop at
// `maxSamples + 1 == GetSampleCount()` (Main.cpp:197).  GL context: GLX pbuffer on a display-less stub
// Xlib (fakex11.c) — see SURVEY.md Appendix H.
//
//   lf_ref_llvmpipe --scene S --spp N --out img.f32 [--probe hits] [--timing-json]
//                   [--bg R G B]   (renderOptions.useConstantBg + bgColor, UI-only in the reference: Main.cpp:504-507)
//                   [--tonemap I] [--vignette INTENSITY POWER] [--ca DISTORTION(0|1) DISTANCE P1 P2 P3]   (RenderOptions the UI sets, Main.cpp:470-500)
//   lf_ref_llvmpipe --extract-assets DIR        (writes the reference's build_include/assets tree)
//
// Output image: W*H*3 float32, rows bottom-up, = GetOutputBufferHDR with tonemapIndex 0
// (accumulated sum / spp).  With --probe hits the shader's last two statements are replaced so the image
// holds (t, triID.x [+0.5 if an analytic light is nearest], matID) of the first camera ray (frame 2).
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#include "Scene.h"
#include "Loader.h"
#include "TiledRenderer.h"
#include "GlobalState.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

using namespace LavaFrame;
LavaFrameState GlobalState;

struct _XDisplay;
extern "C" _XDisplay* FakeOpenDisplay(void);
typedef void* GLXFBConfig;
typedef void* GLXContext;
typedef unsigned long GLXPbuffer;

extern "C" {
extern const unsigned char _binary_shaders_tar_start[], _binary_shaders_tar_end[];
extern const unsigned char _binary_assets_tar_start[], _binary_assets_tar_end[];
}

// gl3w.c binds glXGetProcAddress at link time (gl3w.c:67); forward it to the libGL found at run time so the
// binary does not have to be linked against a Mesa at a fixed path.
typedef void (*GlProc)(void);
extern "C" GlProc glXGetProcAddress(const unsigned char* name) {
    static GlProc (*real)(const unsigned char*) = nullptr;
    if (!real) {
        void* gl = dlopen("libGL.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!gl) { fprintf(stderr, "dlopen libGL.so.1: %s\n", dlerror()); exit(3); }
        real = (GlProc(*)(const unsigned char*))dlsym(gl, "glXGetProcAddress");
    }
    return real(name);
}

static void mkdirs(const std::string& path) {
    for (size_t i = 1; i <= path.size(); i++)
        if (i == path.size() || path[i] == '/') mkdir(path.substr(0, i).c_str(), 0755);
}

// Minimal ustar reader: regular files and directories only.
static bool untar(const unsigned char* p, const unsigned char* end, const std::string& dst) {
    while (p + 512 <= end && p[0] != 0) {
        std::string name(reinterpret_cast<const char*>(p), strnlen(reinterpret_cast<const char*>(p), 100));
        if (p[345]) name = std::string(reinterpret_cast<const char*>(p + 345), strnlen(reinterpret_cast<const char*>(p + 345), 155)) + "/" + name;
        size_t size = strtoul(std::string(reinterpret_cast<const char*>(p + 124), 12).c_str(), nullptr, 8);
        char type = p[156];
        std::string full = dst + "/" + name;
        if (type == '5') {
            mkdirs(full);
        } else if (type == '0' || type == 0) {
            mkdirs(full.substr(0, full.find_last_of('/')));
            FILE* f = fopen(full.c_str(), "wb");
            if (!f) return false;
            fwrite(p + 512, 1, size, f);
            fclose(f);
        }
        p += 512 + ((size + 511) / 512) * 512;
    }
    return true;
}

static std::string slurp(const std::string& path) {
    std::string s;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return s;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
    std::string sceneFile, out = "out.f32", out8, probe, extractDir, shadersOverride;
    int spp = 1;
    bool timingJson = false, previewDof = false;
    float previewScale = 0.f;                        // > 0: render the preview engine's image instead (--preview SCALE)
    int tonemap = 0;
    bool constantBg = false;
    float bg[3] = {0.5f, 0.5f, 0.5f};
    bool useVignette = false, useCA = false, caDistortion = false;
    float vigI = 0.f, vigP = 1.f, caDist = 0.05f, caP1 = 5.f, caP2 = -0.5f, caP3 = 0.5f;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--scene") sceneFile = next();
        else if (a == "--spp") spp = atoi(next().c_str());
        else if (a == "--out") out = next();
        else if (a == "--out8") out8 = next();          // also GetOutputBuffer (8-bit RGB, what SaveFrame* of Export.h write)
        else if (a == "--probe") probe = next();
        else if (a == "--extract-assets") extractDir = next();
        else if (a == "--shaders") shadersOverride = next();
        else if (a == "--timing-json") timingJson = true;
        else if (a == "--preview") previewScale = (float)atof(next().c_str());
        else if (a == "--preview-dof") previewDof = true;
        else if (a == "--tonemap") tonemap = atoi(next().c_str());
        else if (a == "--bg") { constantBg = true; for (int k = 0; k < 3; k++) bg[k] = (float)atof(next().c_str()); }
        else if (a == "--vignette") { useVignette = true; vigI = (float)atof(next().c_str()); vigP = (float)atof(next().c_str()); }
        else if (a == "--ca") {
            useCA = true; caDistortion = atoi(next().c_str()) != 0; caDist = (float)atof(next().c_str());
            caP1 = (float)atof(next().c_str()); caP2 = (float)atof(next().c_str()); caP3 = (float)atof(next().c_str());
        }
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (!extractDir.empty()) {
        mkdirs(extractDir);
        return untar(_binary_assets_tar_start, _binary_assets_tar_end, extractDir) ? 0 : 1;
    }
    if (sceneFile.empty()) { fprintf(stderr, "usage: lf_ref_llvmpipe --scene S --spp N --out img.f32 [--probe hits]\n"); return 2; }

    // ---- shaders: unpack the reference GLSL (unchanged) into a scratch directory -------------------------
    char tmpl[] = "/tmp/lfref_shaders_XXXXXX";
    std::string shadersDir = shadersOverride;
    if (shadersDir.empty()) {
        if (!mkdtemp(tmpl)) { perror("mkdtemp"); return 1; }
        if (!untar(_binary_shaders_tar_start, _binary_shaders_tar_end, tmpl)) return 1;
        shadersDir = std::string(tmpl) + "/shaders/";
    }
    if (probe == "hits") {
        // Probe variant: everything up to and including the accumTexture fetch is the reference text; only the
        // last two statements of main() (renderer.glsl:66,68) are replaced.  common/*.glsl stay unmodified.
        std::string src = slurp(shadersDir + "renderer.glsl");
        size_t a = src.find("vec3 pixelColor = PathTrace(ray);");
        size_t b = src.find("color = pixelColor + accumColor;");
        if (a == std::string::npos || b == std::string::npos) { fprintf(stderr, "probe: renderer.glsl does not look as expected\n"); return 1; }
        size_t e = src.find('\n', b);
        src.replace(a, e - a,
                    "State state; LightSampleRec lrec;\n"
                    "    state.triID = ivec3(-1); state.matID = -1; state.isEmitter = false;\n"
                    "    float t = ClosestHit(ray, state, lrec);\n"
                    "    color = vec3(t, float(state.triID.x) + (state.isEmitter ? 0.5 : 0.0), float(state.matID)) + accumColor * 0.0;");
        FILE* f = fopen((shadersDir + "renderer.glsl").c_str(), "wb");
        fwrite(src.data(), 1, src.size(), f);
        fclose(f);
    } else if (!probe.empty()) {
        fprintf(stderr, "unknown probe %s\n", probe.c_str());
        return 2;
    }

    // ---- GL context on llvmpipe through the stub Xlib ----------------------------------------------------
    void* gl = dlopen("libGL.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!gl) { fprintf(stderr, "dlopen libGL.so.1: %s\n", dlerror()); return 3; }
    auto gpa = (void* (*)(const char*))dlsym(gl, "glXGetProcAddress");   // the library's own, not the forwarder above
    auto choose = (GLXFBConfig * (*)(_XDisplay*, int, const int*, int*)) dlsym(gl, "glXChooseFBConfig");
    auto pbuf = (GLXPbuffer(*)(_XDisplay*, GLXFBConfig, const int*))dlsym(gl, "glXCreatePbuffer");
    auto mkcur = (int (*)(_XDisplay*, GLXPbuffer, GLXPbuffer, GLXContext))dlsym(gl, "glXMakeContextCurrent");
    auto ctxattr = (GLXContext(*)(_XDisplay*, GLXFBConfig, GLXContext, int, const int*))gpa("glXCreateContextAttribsARB");
    _XDisplay* d = FakeOpenDisplay();
    int at[] = {0x8010, 0x5, 0x8011, 0x1, 8, 8, 9, 8, 10, 8, 0};   // DRAWABLE_TYPE pbuffer|window, RGBA, 8/8/8
    int n = 0;
    GLXFBConfig* c = choose(d, 0, at, &n);
    if (!c || n < 1) { fprintf(stderr, "no GLX fbconfig\n"); return 3; }
    int ca[] = {0x2091, 3, 0x2092, 3, 0x9126, 1, 0};               // 3.3 core
    GLXContext ctx = ctxattr(d, c[0], 0, 1, ca);
    int pa[] = {0x8041, 64, 0x8040, 64, 0};
    GLXPbuffer pb = pbuf(d, c[0], pa);
    if (!ctx || !mkcur(d, pb, pb, ctx)) { fprintf(stderr, "glXMakeContextCurrent failed\n"); return 3; }
    if (gl3wInit() != 0) { fprintf(stderr, "gl3wInit failed\n"); return 3; }
    std::string glRenderer = (const char*)glGetString(GL_RENDERER), glVersion = (const char*)glGetString(GL_VERSION);

    // ---- scene + renderer, exactly as Main.cpp does ---------------------------------------------------------
    GlobalState.shadersDir = shadersDir;
    GlobalState.useDebug = getenv("LF_DEBUG") != nullptr;
    RenderOptions ro;
    ro.tonemapIndex = 0;
    ro.useVignette = false; ro.vignetteIntensity = 0; ro.vignettePower = 1;   // left uninitialised by the ctor (Renderer.h:41-43)
    GlobalState.scene = new Scene();
    double tLoad0 = now();
    if (!LoadSceneFromFile(sceneFile, GlobalState.scene, ro)) return 4;
    double tLoad1 = now();
    ro.tonemapIndex = tonemap;
    if (constantBg) { ro.useConstantBg = true; ro.bgColor = Vec3(bg[0], bg[1], bg[2]); }
    ro.useVignette = useVignette; ro.vignetteIntensity = vigI; ro.vignettePower = vigP;
    ro.useCA = useCA; ro.useCADistortion = caDistortion; ro.caDistance = caDist; ro.caP1 = caP1; ro.caP2 = caP2; ro.caP3 = caP3;
    GlobalState.scene->renderOptions = ro;          // Main.cpp:969
    GlobalState.scene->camera->isMoving = false;    // never initialised by Camera's ctor (Camera.cpp:101-117)

    if (previewScale > 0.f) { GlobalState.previewScale = previewScale; GlobalState.useDofInPreview = previewDof; }   // Main.cpp:525-529

    double t0 = now();
    TiledRenderer* r = new TiledRenderer(GlobalState.scene, GlobalState.shadersDir);
    GlobalState.renderer = r;
    try {
        r->Init();
    } catch (std::exception& e) {
        fprintf(stderr, "Init exception: %s\n", e.what());
        return 5;
    }
    glFinish();
    double t1 = now();

    if (previewScale > 0.f) {
        // What the main loop does while the camera moves (Main.cpp:233,101): Update resets the counters and sets
        // maxDepth 2 (TiledRenderer.cpp:471-484,532), Render draws previewEngineShader into previewFBO (:327-333), whose
        // colour attachment (pathTraceTextureLowRes, RGB32F) is read back here while the FBO is still bound.
        GlobalState.scene->camera->isMoving = true;
        r->Update(0.f);
        r->Render();
        glFinish();
        const iVec2 ss = r->GetScreenSize();
        const int pw = (int)(ss.x * previewScale), ph = (int)(ss.y * previewScale);   // glViewport's float -> GLsizei (:330)
        std::vector<float> img((size_t)pw * ph * 3);
        glReadPixels(0, 0, pw, ph, GL_RGB, GL_FLOAT, img.data());
        FILE* f = fopen(out.c_str(), "wb");
        if (!f) { perror(out.c_str()); return 6; }
        fwrite(img.data(), sizeof(float), img.size(), f);
        fclose(f);
        printf("preview %dx%d (scale %g, dof %d) written; glGetError 0x%x\n", pw, ph, previewScale, (int)previewDof, glGetError());
        return 0;
    }

    std::vector<double> stepTimes;
    while (true) {                                   // Main.cpp MainLoop: Update (auto-stop check first) then Render
        GlobalState.scene->camera->isMoving = false;
        if (r->GetSampleCount() == spp + 1) break;
        r->Update(0.f);
        if (r->GetSampleCount() == spp + 1) break;
        double s0 = now();
        r->Render();
        glFinish();
        stepTimes.push_back(now() - s0);
    }
    double t2 = now();

    float* img = nullptr;
    int w = 0, h = 0;
    r->GetOutputBufferHDR(&img, w, h);
    FILE* f = fopen(out.c_str(), "wb");
    if (!f) { perror(out.c_str()); return 6; }
    fwrite(img, sizeof(float), (size_t)w * h * 3, f);
    fclose(f);

    if (!out8.empty()) {
        unsigned char* img8 = nullptr;
        int w8 = 0, h8 = 0;
        r->GetOutputBuffer(&img8, w8, h8);
        FILE* f8 = fopen(out8.c_str(), "wb");
        if (!f8) { perror(out8.c_str()); return 6; }
        fwrite(img8, 1, (size_t)w8 * h8 * 3, f8);
        fclose(f8);
        delete[] img8;
    }

    double sum[3] = {0, 0, 0};
    long nan = 0;
    for (long i = 0; i < (long)w * h; i++)
        for (int k = 0; k < 3; k++) {
            float v = img[3 * i + k];
            if (v != v) nan++; else sum[k] += v;
        }
    // steady state = all tile-steps after the first (the first draw pays the one-time llvmpipe shader JIT)
    double first = stepTimes.empty() ? 0 : stepTimes[0], rest = 0;
    for (size_t i = 1; i < stepTimes.size(); i++) rest += stepTimes[i];
    const RenderOptions& o = GlobalState.scene->renderOptions;
    int tilesPerSample = (int)(ceil((float)w / o.tileWidth) * ceil((float)h / o.tileHeight));
    double steadyPerStep = stepTimes.size() > 1 ? rest / (stepTimes.size() - 1) : first;
    double samplesPerStep = (double)w * h / tilesPerSample;
    const char* threads = getenv("LP_NUM_THREADS");
    if (timingJson) {
        printf("{\"impl\": \"llvmpipe\", \"gl_renderer\": \"%s\", \"gl_version\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %d, "
               "\"tile_steps\": %zu, \"load_s\": %.3f, \"init_s\": %.3f, \"first_step_s\": %.4f, \"steady_s_per_step\": %.6f, "
               "\"samples_per_s_steady\": %.1f, \"render_s\": %.3f, \"lp_num_threads\": \"%s\", \"nproc\": %ld, "
               "\"mean_rgb\": [%.8g, %.8g, %.8g], \"nan\": %ld, \"gl_error\": %u}\n",
               glRenderer.c_str(), glVersion.c_str(), w, h, spp, stepTimes.size(), tLoad1 - tLoad0, t1 - t0, first, steadyPerStep,
               samplesPerStep / steadyPerStep, t2 - t1, threads ? threads : "default", sysconf(_SC_NPROCESSORS_ONLN),
               sum[0] / (w * h), sum[1] / (w * h), sum[2] / (w * h), nan, glGetError());
    } else {
        printf("%s | %s\n%dx%d %d spp: init %.2f s, first step %.2f s, steady %.4f s/step => %.0f samples/s; mean rgb %.6f %.6f %.6f; NaN %ld; glGetError 0x%x\n",
               glRenderer.c_str(), glVersion.c_str(), w, h, spp, t1 - t0, first, steadyPerStep, samplesPerStep / steadyPerStep,
               sum[0] / (w * h), sum[1] / (w * h), sum[2] / (w * h), nan, glGetError());
    }
    if (shadersOverride.empty()) {
        std::string cmd = std::string("rm -rf ") + tmpl;
        if (system(cmd.c_str()) != 0) {}
    }
    return 0;
}
