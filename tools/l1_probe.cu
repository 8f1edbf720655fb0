// l1_probe.cu — development aid: what does a divergent 64-byte record fetch cost on the L1 data pipe of sm_100a?
// Each lane walks a pseudo-random chain of 64-byte "nodes" (the traversal's access pattern) in a table that sits in
// L1 (32 KB), in L2 (8 MB) or beyond L2 (1 GB), with several ways of issuing the loads:
//   0  4 x LDG.128 per lane
//   1  2 x LDG.256 per lane
//   2  lane pairs share the loads: instruction k fetches the node of lane (2p + k), 32 bytes per lane (one line per
//      pair and instruction), then 8 SHFL.BFLY hand the halves to their owners
//   3  as 2 without the shuffles (wrong data, load cost only)
//   4  1 x LDG.256 per lane (half a node)
//   5  1 x LDG.128 per lane
//   6  as 0 with only the first 13 lanes of each warp active (the measured SIMT efficiency of the leaf phase)
//   7  2 x LDG.256 + 8 SHFL.BFLY (shuffle cost on top of variant 1)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l1_probe tools/l1_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

struct f8 { float4 lo, hi; };
__device__ __forceinline__ f8 ldg8(const float4* p) {
    f8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 sx(float4 v) {
    v.x = __shfl_xor_sync(0xffffffffu, v.x, 1); v.y = __shfl_xor_sync(0xffffffffu, v.y, 1);
    v.z = __shfl_xor_sync(0xffffffffu, v.z, 1); v.w = __shfl_xor_sync(0xffffffffu, v.w, 1);
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(128, 8) k_probe(const float4* __restrict__ nodes, unsigned mask, int steps, unsigned* sink) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned idx = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    float acc = 0.f;
    if (MODE == 6 && lane >= 13) return;
    for (int s = 0; s < steps; s++) {
        unsigned n = (idx >> 7) & mask;                 // node index, different in every lane
        const float4* p = nodes + (size_t)4 * n;
        float4 a, b, c, d;
        if (MODE == 0 || MODE == 6) { a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2); d = __ldg(p + 3); }
        else if (MODE == 1 || MODE == 7) { f8 x = ldg8(p), y = ldg8(p + 2); a = x.lo; b = x.hi; c = y.lo; d = y.hi; }
        else if (MODE == 2 || MODE == 3) {
            unsigned pn = __shfl_xor_sync(0xffffffffu, n, 1);
            unsigned even = (lane & 1u) ? pn : n, odd = (lane & 1u) ? n : pn;
            f8 x = ldg8(nodes + (size_t)4 * even + 2 * (lane & 1u));   // node of the even lane, my half
            f8 y = ldg8(nodes + (size_t)4 * odd + 2 * (lane & 1u));    // node of the odd lane, my half
            if (MODE == 2) {
                f8 mine = (lane & 1u) ? y : x, theirs = (lane & 1u) ? x : y;
                float4 g0 = sx(theirs.lo), g1 = sx(theirs.hi);
                a = (lane & 1u) ? g0 : mine.lo; b = (lane & 1u) ? g1 : mine.hi;
                c = (lane & 1u) ? mine.lo : g0; d = (lane & 1u) ? mine.hi : g1;
            } else { a = x.lo; b = x.hi; c = y.lo; d = y.hi; }
        }
        else if (MODE == 4) { f8 x = ldg8(p); a = x.lo; b = x.hi; c = a; d = b; }
        else { a = __ldg(p); b = c = d = a; }
        if (MODE == 7) { a = sx(a); b = sx(b); }
        float v = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w)) + ((c.x + c.y) + (c.z + c.w)) + ((d.x + d.y) + (d.z + d.w));
        acc += v;
        idx = idx * 1664525u + 1013904223u + __float_as_uint(d.w);   // next node depends on the data (a walk)
    }
    if (acc == 123.456f) *sink = 1;
}

template <int MODE>
static void run(const char* name, const float4* nodes, unsigned nnodes, unsigned* sink, int sms) {
    int steps = 2000;
    int blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_probe<MODE><<<blocks, 128>>>(nodes, nnodes - 1, 200, sink);
    cudaEventRecord(e0);
    k_probe<MODE><<<blocks, 128>>>(nodes, nnodes - 1, steps, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double lanes = (MODE == 6) ? 13.0 / 32.0 : 1.0;
    double visits = (double)blocks * 128 * lanes * steps;
    double clk = 1.965e9;
    printf("  %-44s %8.3f ms  %7.2f G node-visits/s  %6.3f visits/clk/SM  %7.0f GB/s (64 B)\n", name, ms, visits / ms / 1e6,
           visits / (ms * 1e-3) / clk / sms, visits * 64 / ms / 1e6);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, sms);
    unsigned* sink;
    cudaMalloc(&sink, 4);
    const size_t sizes[3] = {32u << 10, 8u << 20, 1u << 30};
    const char* names[3] = {"32 KB table (L1 hits)", "8 MB table (L2 hits)", "1 GB table (HBM)"};
    for (int t = 0; t < 3; t++) {
        size_t bytes = sizes[t];
        float4* nodes;
        cudaMalloc(&nodes, bytes);
        cudaMemset(nodes, 0, bytes);
        unsigned nnodes = (unsigned)(bytes / 64);
        printf("%s\n", names[t]);
        run<0>("4 x LDG.128", nodes, nnodes, sink, sms);
        run<1>("2 x LDG.256", nodes, nnodes, sink, sms);
        run<2>("pair-shared 2 x LDG.256 + 8 SHFL", nodes, nnodes, sink, sms);
        run<3>("pair-shared 2 x LDG.256, no SHFL", nodes, nnodes, sink, sms);
        run<4>("1 x LDG.256 (32 B)", nodes, nnodes, sink, sms);
        run<5>("1 x LDG.128 (16 B)", nodes, nnodes, sink, sms);
        run<6>("4 x LDG.128, 13 of 32 lanes", nodes, nnodes, sink, sms);
        run<7>("2 x LDG.256 + 8 SHFL", nodes, nnodes, sink, sms);
        cudaFree(nodes);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
