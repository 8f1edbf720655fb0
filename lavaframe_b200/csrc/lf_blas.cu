// lf_blas.cu — lfcuda_build_blas: the bottom-level BVH of one mesh built ON THE DEVICE, node for node what the reference's host builder
// produces (Mesh::BuildBVH, LavaFrame/Mesh.cpp:93-111; RadeonRays::SplitBvh(2.0f, 64, 0, 0.001f, 0), split_bvh.cpp:11-289;
// BvhTranslator::ProcessBLASNodes, bvh_translator.cpp:35-60).
//
// The algorithm is the text of lf_blas_build.h (see there for how the sequential builder becomes order-independent data-parallel steps); this
// file is its CUDA executor: every step is one launch of k_blas_step<Step> (grid-stride, one work item per thread, accumulation by atomics),
// the level loop runs on the host with ONE 4-byte read-back per level (the number of nodes that split).  A level of C2's 869 880-triangle mesh
// is about 16 launches; the tree has 26 levels.  Nothing here runs on the CPU but that loop; there is no host fallback.
#include <chrono>
#include <cstdio>

#include <cuda_runtime.h>

#include "lf_blas_build.h"
#include "lf_ctx_internal.h"
#include "lfcuda.h"

namespace {

using namespace lf::blas;

template <class F>
__global__ void __launch_bounds__(256) k_blas_step(F f, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) f(i);
}

struct DevExec {
    cudaStream_t stream;
    int max_blocks;
    int launches = 0;
    cudaError_t err = cudaSuccess;
    template <class F>
    void run(int n, const F& f) {
        if (n <= 0 || err != cudaSuccess) return;
        int blocks = (n + 255) / 256;
        if (blocks > max_blocks) blocks = max_blocks;
        k_blas_step<F><<<blocks, 256, 0, stream>>>(f, n);
        launches++;
    }
    int read_int(const int* p) {
        int v = 0;
        if (err != cudaSuccess) return 0;
        err = cudaMemcpyAsync(&v, p, sizeof(int), cudaMemcpyDeviceToHost, stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(stream);
        return err == cudaSuccess ? v : 0;
    }
};

struct Arena {                       // every device array of one build comes out of ONE allocation (a cudaMalloc / cudaFree pair costs more than a level)
    char* base = nullptr;
    size_t used = 0, cap = 0;
    template <class T>
    T* get(size_t count) {           // pass 1 (base == nullptr): sizes only
        const size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
        T* p = base ? (T*)(base + used) : nullptr;
        used += bytes;
        return p;
    }
    cudaError_t commit() { cap = used; used = 0; return cudaMalloc((void**)&base, cap); }
    ~Arena() { if (base) cudaFree(base); }
};

int blas_fail(int code, const char* what, cudaError_t e) {
    char buf[512];
    snprintf(buf, sizeof buf, "lfcuda_build_blas: %s%s%s", what, e != cudaSuccess ? ": " : "", e != cudaSuccess ? cudaGetErrorString(e) : "");
    return lf::ctx_fail(nullptr, code, buf);
}

}  // namespace

extern "C" int lfcuda_build_blas(int32_t device, const float* prim_bounds, int32_t num_prims, float traversal_cost, int32_t num_bins,
                                 float* out_nodes, int32_t* out_indices, LfBlasInfo* info) {
    if (!prim_bounds || !out_nodes || !out_indices || !info) return blas_fail(LFCUDA_EINVAL, "NULL argument", cudaSuccess);
    if (num_prims < 1) return blas_fail(LFCUDA_EINVAL, "num_prims < 1", cudaSuccess);
    if (num_prims > (1 << 30)) return blas_fail(LFCUDA_ELIMIT, "more than 2^30 primitives", cudaSuccess);
    if (num_bins < 2 || num_bins > kMaxBins) return blas_fail(LFCUDA_EINVAL, "num_bins outside 2..64", cudaSuccess);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return blas_fail(LFCUDA_ECUDA, "no CUDA device available; liblfcuda has no CPU fallback", e);
    if (device < 0 || device >= ndev) return blas_fail(LFCUDA_EINVAL, "device out of range", cudaSuccess);
    int prev_device = -1;
    cudaGetDevice(&prev_device);                                   // the caller's current device is put back before returning
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_device};
    if ((e = cudaSetDevice(device)) != cudaSuccess) return blas_fail(LFCUDA_ECUDA, "cudaSetDevice", e);
    const auto t0 = std::chrono::steady_clock::now();
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaStream_t stream = nullptr;
    if ((e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)) != cudaSuccess) return blas_fail(LFCUDA_ECUDA, "cudaStreamCreate", e);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    int rc = 0;
    {
        const int n = num_prims;
        const int lcap = level_capacity(n);
        Arena A;
        State S{};
        S.n = n; S.nbins = num_bins; S.tc = traversal_cost;
        float* d_in = nullptr;
        cudaError_t ea = cudaSuccess;
        for (int pass = 0; pass < 2 && ea == cudaSuccess; pass++) {       // sizes, then pointers
            d_in = A.get<float>(6 * (size_t)n);
            S.in_bounds = d_in;
            for (int k = 0; k < 2; k++) { S.lo[k] = A.get<float4>(n); S.hi[k] = A.get<float4>(n); S.node_of[k] = A.get<int>(n); S.lev[k] = A.get<LevelNode>(lcap); }
            S.flag = A.get<int>((size_t)n + 1); S.scan = A.get<int>((size_t)n + 1); S.chunk = A.get<int>(chunk_capacity(n));
            S.pairL = A.get<int>(n); S.pairR = A.get<int>(n);
            S.nflag = A.get<int>(lcap); S.nscan = A.get<int>(lcap);
            S.g = A.get<GNode>(2 * (size_t)n);
            S.bin_cap = lcap < 65536 ? lcap : 65536;                       // 5 376 bytes of bins per inner node; wider levels go in batches
            S.bins = A.get<float>((size_t)S.bin_cap * 3 * kBinFields * kMaxBins);
            S.misc = A.get<int>(4);
            S.out_nodes = A.get<float>(9 * 2 * (size_t)n);
            S.out_indices = A.get<int>(n);
            if (pass == 0) ea = A.commit();
        }
        if (ea != cudaSuccess) rc = blas_fail(ea == cudaErrorMemoryAllocation ? LFCUDA_ENOMEM : LFCUDA_ECUDA, "device allocation", ea);
        if (!rc && (e = cudaMemcpyAsync(d_in, prim_bounds, 6 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, stream)) != cudaSuccess)
            rc = blas_fail(LFCUDA_ECUDA, "upload of the primitive bounds", e);
        if (!rc) {
            DevExec ex{stream, sms * 16};
            cudaEventRecord(ev0, stream);
            Result R = build(ex, S);
            cudaEventRecord(ev1, stream);
            if (ex.err == cudaSuccess) ex.err = cudaGetLastError();
            if (ex.err == cudaSuccess) ex.err = cudaMemcpyAsync(out_nodes, S.out_nodes, 9 * (size_t)R.num_nodes * sizeof(float), cudaMemcpyDeviceToHost, stream);
            if (ex.err == cudaSuccess) ex.err = cudaMemcpyAsync(out_indices, S.out_indices, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, stream);
            if (ex.err == cudaSuccess) ex.err = cudaStreamSynchronize(stream);
            if (ex.err != cudaSuccess) rc = blas_fail(LFCUDA_ECUDA, "build steps", ex.err);
            else {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev0, ev1);
                info->num_nodes = R.num_nodes; info->num_indices = n; info->height = R.height; info->negative_zero = R.negative_zero;
                info->levels = R.levels; info->launches = ex.launches; info->build_ms = ms;
                info->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
            }
        }
        cudaStreamSynchronize(stream);
    }
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    cudaStreamDestroy(stream);
    return rc;
}
