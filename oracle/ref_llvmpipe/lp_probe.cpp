// lp_probe — TEST INFRASTRUCTURE.  Evaluates one GLSL expression per texel on the same Mesa llvmpipe the tier-1 oracle
// (lf_ref_llvmpipe) runs on, so that the oracle's restatement of GLSL built-ins (sin/cos/exp/log/pow/acos/atan, LINEAR
// RGBA8 / RGB32F filtering) can be pinned bit for bit against what the reference's shaders actually compute there.
// GL context: the same display-less GLX pbuffer route as lf_ref_llvmpipe.cpp.  Nothing in the product links this.
//
//   lp_probe --in args.f32 --n N --expr 'vec4(sin(a.x), cos(a.x), 0, 0)' --out res.f32
//            [--tex8 W H L file.rgba8]      sampler2DArray tex8  (GL_RGBA8, LINEAR/LINEAR, REPEAT: Renderer.cpp:152-161)
//            [--texf W H file.rgb32f]       sampler2D texf       (GL_RGB32F, LINEAR/LINEAR: Renderer.cpp:166-172)
//            [--pre 'GLSL declarations']      inserted before main(); may call `vec4 A(int k)`
//            [--stride K]                     K vec4 per item: `a` = A(0), the others through A(1) .. A(K-1); args.f32 then holds N*K vec4
// args.f32 holds N vec4 (16 bytes each) = `a` in the expression; res.f32 receives N vec4.
#include <GL/gl3w.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

struct _XDisplay;
extern "C" _XDisplay* FakeOpenDisplay(void);
typedef void* GLXFBConfig;
typedef void* GLXContext;
typedef unsigned long GLXPbuffer;
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

static std::vector<unsigned char> slurp(const std::string& path) {
    std::vector<unsigned char> s;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { perror(path.c_str()); exit(2); }
    unsigned char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.insert(s.end(), buf, buf + n);
    fclose(f);
    return s;
}

static GLuint compile(GLenum type, const std::string& src) {
    GLuint s = glCreateShader(type);
    const char* p = src.c_str();
    glShaderSource(s, 1, &p, nullptr);
    glCompileShader(s);
    GLint ok = 0;
    glGetShaderiv(s, GL_COMPILE_STATUS, &ok);
    if (!ok) {
        char log[4096];
        glGetShaderInfoLog(s, sizeof log, nullptr, log);
        fprintf(stderr, "shader compile failed:\n%s\n%s\n", log, src.c_str());
        exit(4);
    }
    return s;
}

int main(int argc, char** argv) {
    std::string in, out = "res.f32", expr = "a", pre, tex8File, texfFile;
    long n = 0;
    int stride = 1;
    int t8w = 0, t8h = 0, t8l = 0, tfw = 0, tfh = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--in") in = next();
        else if (a == "--n") n = atol(next().c_str());
        else if (a == "--expr") expr = next();
        else if (a == "--pre") pre = next();
        else if (a == "--stride") stride = atoi(next().c_str());
        else if (a == "--out") out = next();
        else if (a == "--tex8") { t8w = atoi(next().c_str()); t8h = atoi(next().c_str()); t8l = atoi(next().c_str()); tex8File = next(); }
        else if (a == "--texf") { tfw = atoi(next().c_str()); tfh = atoi(next().c_str()); texfFile = next(); }
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (in.empty() || n <= 0) { fprintf(stderr, "usage: lp_probe --in args.f32 --n N --expr E --out res.f32\n"); return 2; }
    std::vector<unsigned char> args = slurp(in);
    if (stride < 1 || (long)args.size() < n * 16 * stride) { fprintf(stderr, "input too short\n"); return 2; }
    const int W = 1024;
    const int H = (int)((n + W - 1) / W);
    const int IW = 1024, IH = (int)((n * stride + IW - 1) / IW);       // argument texture: item i at texels [i * stride, (i + 1) * stride)
    args.resize((size_t)IW * IH * 16, 0);

    void* gl = dlopen("libGL.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!gl) { fprintf(stderr, "dlopen libGL.so.1: %s\n", dlerror()); return 3; }
    auto gpa = (void* (*)(const char*))dlsym(gl, "glXGetProcAddress");
    auto choose = (GLXFBConfig * (*)(_XDisplay*, int, const int*, int*)) dlsym(gl, "glXChooseFBConfig");
    auto pbuf = (GLXPbuffer(*)(_XDisplay*, GLXFBConfig, const int*))dlsym(gl, "glXCreatePbuffer");
    auto mkcur = (int (*)(_XDisplay*, GLXPbuffer, GLXPbuffer, GLXContext))dlsym(gl, "glXMakeContextCurrent");
    auto ctxattr = (GLXContext(*)(_XDisplay*, GLXFBConfig, GLXContext, int, const int*))gpa("glXCreateContextAttribsARB");
    _XDisplay* d = FakeOpenDisplay();
    int at[] = {0x8010, 0x5, 0x8011, 0x1, 8, 8, 9, 8, 10, 8, 0};
    int nc = 0;
    GLXFBConfig* c = choose(d, 0, at, &nc);
    if (!c || nc < 1) { fprintf(stderr, "no GLX fbconfig\n"); return 3; }
    int ca[] = {0x2091, 3, 0x2092, 3, 0x9126, 1, 0};
    GLXContext ctx = ctxattr(d, c[0], 0, 1, ca);
    int pa[] = {0x8041, 64, 0x8040, 64, 0};
    GLXPbuffer pb = pbuf(d, c[0], pa);
    if (!ctx || !mkcur(d, pb, pb, ctx)) { fprintf(stderr, "glXMakeContextCurrent failed\n"); return 3; }
    if (gl3wInit() != 0) { fprintf(stderr, "gl3wInit failed\n"); return 3; }

    GLuint inTex, outTex, fbo, vao;
    glGenTextures(1, &inTex);
    glActiveTexture(GL_TEXTURE0);
    glBindTexture(GL_TEXTURE_2D, inTex);
    glTexImage2D(GL_TEXTURE_2D, 0, GL_RGBA32F, IW, IH, 0, GL_RGBA, GL_FLOAT, args.data());
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_NEAREST);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_NEAREST);
    glGenTextures(1, &outTex);
    glActiveTexture(GL_TEXTURE3);
    glBindTexture(GL_TEXTURE_2D, outTex);
    glTexImage2D(GL_TEXTURE_2D, 0, GL_RGBA32F, W, H, 0, GL_RGBA, GL_FLOAT, nullptr);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_NEAREST);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_NEAREST);
    glGenFramebuffers(1, &fbo);
    glBindFramebuffer(GL_FRAMEBUFFER, fbo);
    glFramebufferTexture2D(GL_FRAMEBUFFER, GL_COLOR_ATTACHMENT0, GL_TEXTURE_2D, outTex, 0);
    if (glCheckFramebufferStatus(GL_FRAMEBUFFER) != GL_FRAMEBUFFER_COMPLETE) { fprintf(stderr, "fbo incomplete\n"); return 3; }

    if (!tex8File.empty()) {
        std::vector<unsigned char> px = slurp(tex8File);
        if (px.size() < (size_t)t8w * t8h * t8l * 4) { fprintf(stderr, "tex8 too short\n"); return 2; }
        GLuint t;
        glGenTextures(1, &t);
        glActiveTexture(GL_TEXTURE1);
        glBindTexture(GL_TEXTURE_2D_ARRAY, t);
        glTexImage3D(GL_TEXTURE_2D_ARRAY, 0, GL_RGBA8, t8w, t8h, t8l, 0, GL_RGBA, GL_UNSIGNED_BYTE, px.data());
        glTexParameteri(GL_TEXTURE_2D_ARRAY, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
        glTexParameteri(GL_TEXTURE_2D_ARRAY, GL_TEXTURE_MIN_FILTER, GL_LINEAR);
    }
    if (!texfFile.empty()) {
        std::vector<unsigned char> px = slurp(texfFile);
        if (px.size() < (size_t)tfw * tfh * 12) { fprintf(stderr, "texf too short\n"); return 2; }
        GLuint t;
        glGenTextures(1, &t);
        glActiveTexture(GL_TEXTURE2);
        glBindTexture(GL_TEXTURE_2D, t);
        glPixelStorei(GL_UNPACK_ALIGNMENT, 1);
        glTexImage2D(GL_TEXTURE_2D, 0, GL_RGB32F, tfw, tfh, 0, GL_RGB, GL_FLOAT, px.data());
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR);
    }

    const std::string vs =
        "#version 330\nvoid main() { vec2 p = vec2((gl_VertexID & 1) * 4 - 1, (gl_VertexID & 2) * 2 - 1); gl_Position = vec4(p, 0, 1); }\n";
    const std::string fs = "#version 330\nprecision highp float;\nuniform sampler2D inTex;\nuniform sampler2DArray tex8;\nuniform sampler2D texf;\n"
                           "out vec4 res;\n"
                           "vec4 A(int k) { int j = (int(gl_FragCoord.y) * 1024 + int(gl_FragCoord.x)) * " + std::to_string(stride) + " + k; return texelFetch(inTex, ivec2(j % 1024, j / 1024), 0); }\n" +
                           pre + "\nvoid main() {\n  vec4 a = A(0);\n  res = " + expr + ";\n}\n";
    GLuint prog = glCreateProgram();
    glAttachShader(prog, compile(GL_VERTEX_SHADER, vs));
    glAttachShader(prog, compile(GL_FRAGMENT_SHADER, fs));
    glLinkProgram(prog);
    GLint ok = 0;
    glGetProgramiv(prog, GL_LINK_STATUS, &ok);
    if (!ok) { fprintf(stderr, "link failed\n"); return 4; }
    glUseProgram(prog);
    glUniform1i(glGetUniformLocation(prog, "inTex"), 0);
    glUniform1i(glGetUniformLocation(prog, "tex8"), 1);
    glUniform1i(glGetUniformLocation(prog, "texf"), 2);
    glGenVertexArrays(1, &vao);
    glBindVertexArray(vao);
    glViewport(0, 0, W, H);
    glDisable(GL_BLEND);
    glDisable(GL_DEPTH_TEST);
    glDrawArrays(GL_TRIANGLES, 0, 3);
    glFinish();
    std::vector<float> res((size_t)W * H * 4);
    glPixelStorei(GL_PACK_ALIGNMENT, 1);
    glReadPixels(0, 0, W, H, GL_RGBA, GL_FLOAT, res.data());
    FILE* f = fopen(out.c_str(), "wb");
    if (!f) { perror(out.c_str()); return 6; }
    fwrite(res.data(), 16, (size_t)n, f);
    fclose(f);
    GLenum e = glGetError();
    if (e) fprintf(stderr, "glGetError 0x%x\n", e);
    return 0;
}
