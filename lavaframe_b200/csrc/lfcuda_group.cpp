// lfcuda_group.cpp — several GPUs behind ONE renderer in ONE process (include/lfcuda.h, "multi-GPU in one process").
//
// The reference constructs a single renderer (`new TiledRenderer(...)`, LavaFrame/Main.cpp:91); a drop-in that uses the whole 8 x B200
// box must therefore fan out below that one object.  A group owns one lfcuda context and one host worker thread per device.  The path
// shards by sample index exactly as in the multi-process case (`frame` is the RNG seed, globals.glsl:116-120): a call for
// frames first, first + stride, ... deals them round-robin to the devices, every device accumulates its share into its own buffer, and
// no device ever waits for another while rendering.  The only exchange is at read-out: the post-process kernel of the group's first
// device reads every device's accumulation buffer directly - peers through CUDA peer access, i.e. plain loads over NVLink / NVSwitch -
// adds them and tonemaps in the same pass (lf_kernels.cu k_post_sum).  The summed image is never materialised unless asked for
// (lfcuda_group_read_accum), and the local buffers stay local, so rendering simply continues after a read-out.
// (The one-process-per-GPU deployment, bench.py under torchrun, sums with NCCL instead: lfcuda_reduce.)
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "lf_ctx_internal.h"
#include "lf_kernels.h"

using namespace lf;

namespace {

thread_local std::string g_group_create_error;

// One host thread per device: it owns every call into that device's context, so the contexts never see concurrent callers and the
// launch sequences of the devices are issued in parallel (25 launches per bounce loop and batch; serial issue from one thread would
// stagger the devices by the launch cost of all their predecessors).
class Worker {
public:
    Worker() : th([this] { run(); }) {}
    ~Worker() {
        { std::lock_guard<std::mutex> l(m); stop = true; }
        cv.notify_all();
        th.join();
    }
    void post(std::function<void()> f) {
        { std::lock_guard<std::mutex> l(m); jobs.push(std::move(f)); pending++; }
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> l(m);
        done.wait(l, [this] { return pending == 0; });
    }

private:
    void run() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(m);
                cv.wait(l, [this] { return stop || !jobs.empty(); });
                if (jobs.empty()) return;
                f = std::move(jobs.front());
                jobs.pop();
            }
            f();
            { std::lock_guard<std::mutex> l(m); pending--; }
            done.notify_all();
        }
    }
    std::mutex m;
    std::condition_variable cv, done;
    std::queue<std::function<void()>> jobs;
    int pending = 0;
    bool stop = false;
    std::thread th;
};

}  // namespace

struct lfcuda_group {
    std::vector<int> devices;
    std::vector<lfcuda_ctx*> ctx;
    std::vector<Worker*> workers;
    std::vector<int> rc;                 // result of the last job per device
    std::string err;
    float* d_sum = nullptr; size_t sum_floats = 0;   // the summed accumulation, on device 0, only for lfcuda_group_read_accum
    bool peers_enabled = false;
};

namespace {

int gfail(lfcuda_group* g, int code, const std::string& msg) {
    if (g) g->err = msg; else g_group_create_error = msg;
    return code;
}

// Run fn(i, ctx_i) on every device's worker and wait; the first failure is reported with its device.
int for_all(lfcuda_group* g, const std::function<int(int, lfcuda_ctx*)>& fn) {
    const int n = (int)g->ctx.size();
    for (int i = 0; i < n; i++) g->workers[i]->post([g, i, &fn] { g->rc[i] = fn(i, g->ctx[i]); });
    for (int i = 0; i < n; i++) g->workers[i]->wait();
    for (int i = 0; i < n; i++)
        if (g->rc[i] != 0) return gfail(g, g->rc[i], "device " + std::to_string(g->devices[i]) + ": " + lfcuda_last_error(g->ctx[i]));
    return 0;
}

// Device 0 of the group reads the other devices' accumulation buffers in its post-process kernel: peer access, once.
int enable_peers(lfcuda_group* g) {
    if (g->peers_enabled) return 0;
    const int root = g->devices[0];
    cudaError_t e = cudaSetDevice(root);
    for (size_t i = 1; i < g->devices.size() && e == cudaSuccess; i++) {
        const int d = g->devices[i];
        if (d == root) continue;                       // a second context on the same device: already addressable
        int can = 0;
        e = cudaDeviceCanAccessPeer(&can, root, d);
        if (e != cudaSuccess) break;
        if (!can) return gfail(g, LFCUDA_ECUDA, "device " + std::to_string(root) + " cannot access device " + std::to_string(d) + " as a peer (no NVLink / PCIe P2P path)");
        e = cudaDeviceEnablePeerAccess(d, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    }
    if (e != cudaSuccess) return gfail(g, LFCUDA_ECUDA, std::string("peer access: ") + cudaGetErrorString(e));
    g->peers_enabled = true;
    return 0;
}

// All devices finish their queued frames, then device 0 runs the fused sum + post-process pass over every accumulation buffer.
int read_common(lfcuda_group* g, float inv, int tonemap, float* out_f, uint8_t* out_u8, float* out_sum) {
    int r = for_all(g, [](int, lfcuda_ctx* c) { return lfcuda_synchronize(c); });
    if (r) return r;
    if ((r = enable_peers(g))) return r;
    std::vector<CtxView> v(g->ctx.size());
    for (size_t i = 0; i < g->ctx.size(); i++)
        if (!ctx_view(g->ctx[i], &v[i])) return gfail(g, LFCUDA_EINVAL, "render parameters not set (lfcuda_group_set_params)");
    const CtxView& root = v[0];
    const float* ptrs[kMaxGroupDevices];
    for (size_t i = 0; i < v.size(); i++) ptrs[i] = v[i].accum;
    cudaError_t e = cudaSetDevice(root.device);
    if (e == cudaSuccess && out_sum && g->sum_floats != root.accum_floats) {
        if (g->d_sum) cudaFree(g->d_sum);
        g->d_sum = nullptr; g->sum_floats = 0;
        e = cudaMalloc((void**)&g->d_sum, root.accum_floats * sizeof(float));
        if (e == cudaSuccess) g->sum_floats = root.accum_floats;
    }
    if (e != cudaSuccess) return gfail(g, LFCUDA_ECUDA, std::string("group read-out: ") + cudaGetErrorString(e));
    ctx_count_launch(g->ctx[0]);
    launch_post_sum(root.stream, ptrs, (int)v.size(), out_f ? root.out_f : nullptr, out_u8 ? root.out_u8 : nullptr, out_sum ? g->d_sum : nullptr,
                    root.width, root.height, inv, tonemap, root.post);
    e = cudaGetLastError();
    if (e == cudaSuccess && out_f) e = cudaMemcpyAsync(out_f, root.out_f, root.accum_floats * sizeof(float), cudaMemcpyDeviceToHost, root.stream);
    if (e == cudaSuccess && out_u8) e = cudaMemcpyAsync(out_u8, root.out_u8, root.accum_floats, cudaMemcpyDeviceToHost, root.stream);
    if (e == cudaSuccess && out_sum) e = cudaMemcpyAsync(out_sum, g->d_sum, root.accum_floats * sizeof(float), cudaMemcpyDeviceToHost, root.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(root.stream);
    if (e != cudaSuccess) return gfail(g, LFCUDA_ECUDA, std::string("group read-out: ") + cudaGetErrorString(e));
    return 0;
}

}  // namespace

extern "C" {

int lfcuda_group_create(lfcuda_group** out, const int32_t* devices, int32_t ndev) {
    if (!out) return gfail(nullptr, LFCUDA_EINVAL, "out is NULL");
    *out = nullptr;
    if (!devices || ndev < 1 || ndev > kMaxGroupDevices) return gfail(nullptr, LFCUDA_EINVAL, "a group has 1 to " + std::to_string(kMaxGroupDevices) + " devices");
    lfcuda_group* g = new lfcuda_group;
    g->devices.assign(devices, devices + ndev);
    g->rc.assign(ndev, 0);
    for (int i = 0; i < ndev; i++) {
        lfcuda_ctx* c = nullptr;
        int r = lfcuda_create(&c, devices[i]);
        if (r != 0) {
            std::string msg = lfcuda_last_error(nullptr);
            for (lfcuda_ctx* p : g->ctx) lfcuda_destroy(p);
            delete g;
            return gfail(nullptr, r, msg);
        }
        g->ctx.push_back(c);
    }
    for (int i = 0; i < ndev; i++) g->workers.push_back(new Worker);
    *out = g;
    return 0;
}

void lfcuda_group_destroy(lfcuda_group* g) {
    if (!g) return;
    for (size_t i = 0; i < g->ctx.size(); i++) g->workers[i]->post([g, i] { lfcuda_destroy(g->ctx[i]); });
    for (Worker* w : g->workers) { w->wait(); delete w; }
    if (g->d_sum) { cudaSetDevice(g->devices[0]); cudaFree(g->d_sum); }
    delete g;
}

const char* lfcuda_group_last_error(const lfcuda_group* g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }
int lfcuda_group_size(const lfcuda_group* g) { return g ? (int)g->ctx.size() : 0; }
lfcuda_ctx* lfcuda_group_ctx(lfcuda_group* g, int32_t index) { return (g && index >= 0 && index < (int)g->ctx.size()) ? g->ctx[index] : nullptr; }

int lfcuda_group_upload_scene(lfcuda_group* g, const LfSceneView* scene) {
    if (!g || !scene) return LFCUDA_EINVAL;
    return for_all(g, [scene](int, lfcuda_ctx* c) { return lfcuda_upload_scene(c, scene); });   // re-pack + upload on every device, in parallel
}
int lfcuda_group_update_instances(lfcuda_group* g, const float* transforms, int32_t num_instances, const float* materials, int32_t num_materials,
                                  const float* tlas_nodes, int32_t first_node, int32_t num_tlas_nodes) {
    if (!g) return LFCUDA_EINVAL;
    return for_all(g, [=](int, lfcuda_ctx* c) { return lfcuda_update_instances(c, transforms, num_instances, materials, num_materials, tlas_nodes, first_node, num_tlas_nodes); });
}
int lfcuda_group_update_instances_device(lfcuda_group* g, const float* transforms, int32_t num_instances, const float* materials, int32_t num_materials,
                                         const int32_t* instance_material_ids) {
    if (!g) return LFCUDA_EINVAL;
    return for_all(g, [=](int, lfcuda_ctx* c) { return lfcuda_update_instances_device(c, transforms, num_instances, materials, num_materials, instance_material_ids); });
}
int lfcuda_group_set_params(lfcuda_group* g, const LfParams* p) {
    if (!g || !p) return LFCUDA_EINVAL;
    return for_all(g, [p](int, lfcuda_ctx* c) { return lfcuda_set_params(c, p); });
}
int lfcuda_group_set_camera(lfcuda_group* g, const LfCamera* cam) {
    if (!g || !cam) return LFCUDA_EINVAL;
    return for_all(g, [cam](int, lfcuda_ctx* c) { return lfcuda_set_camera(c, cam); });
}
int lfcuda_group_set_post(lfcuda_group* g, const LfPostParams* p) {
    if (!g) return LFCUDA_EINVAL;
    return for_all(g, [p](int, lfcuda_ctx* c) { return lfcuda_set_post(c, p); });
}
int lfcuda_group_clear(lfcuda_group* g) {
    if (!g) return LFCUDA_EINVAL;
    return for_all(g, [](int, lfcuda_ctx* c) { return lfcuda_clear(c); });
}
int lfcuda_group_synchronize(lfcuda_group* g) {
    if (!g) return LFCUDA_EINVAL;
    return for_all(g, [](int, lfcuda_ctx* c) { return lfcuda_synchronize(c); });
}

// Frames first, first + stride, ... (nframes of them) of tile (tile_x, tile_y), dealt round-robin: device i renders every ndev-th one
// starting with the i-th.  Returns when every device has ENQUEUED its share (the devices then run concurrently).
int lfcuda_group_render_frames(lfcuda_group* g, int32_t first_frame, int32_t nframes, int32_t frame_stride, int32_t tile_x, int32_t tile_y) {
    if (!g) return LFCUDA_EINVAL;
    if (nframes < 0) return gfail(g, LFCUDA_EINVAL, "nframes < 0");
    const int n = (int)g->ctx.size();
    return for_all(g, [=](int i, lfcuda_ctx* c) {
        const int mine = nframes > i ? (nframes - i + n - 1) / n : 0;
        if (mine == 0) return 0;
        return lfcuda_render_frames(c, first_frame + i * frame_stride, mine, frame_stride * n, tile_x, tile_y);
    });
}

int lfcuda_group_read_output(lfcuda_group* g, float inv_sample_counter, int32_t tonemap_index, float* rgb_out) {
    if (!g || !rgb_out) return LFCUDA_EINVAL;
    return read_common(g, inv_sample_counter, tonemap_index, rgb_out, nullptr, nullptr);
}
int lfcuda_group_read_output_u8(lfcuda_group* g, float inv_sample_counter, int32_t tonemap_index, uint8_t* rgb_out) {
    if (!g || !rgb_out) return LFCUDA_EINVAL;
    return read_common(g, inv_sample_counter, tonemap_index, nullptr, rgb_out, nullptr);
}
int lfcuda_group_read_accum(lfcuda_group* g, float* rgb_out) {
    if (!g || !rgb_out) return LFCUDA_EINVAL;
    return read_common(g, 1.0f, 0, nullptr, nullptr, rgb_out);
}

}  // extern "C"
