// tests/hostcheck/lf_blascheck.cpp — TEST INFRASTRUCTURE, not a product path.
//
// The data-parallel BLAS build of lavaframe_b200/csrc/lf_blas_build.h (the text lf_blas.cu launches as CUDA kernels) compiled for the host:
// every step runs as a loop over its work items - forwards, backwards or in a scrambled order - so that tests/test_blas_build.py can check,
// without a GPU, (1) that the level-synchronous restatement builds the reference's tree node for node (against the node arrays RadeonRays'
// SplitBvh produced for the committed scene packs) and (2) that no step depends on the order of its items.  Nothing in the product calls it.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lf_blas_build.h"

namespace {
using namespace lf::blas;
struct HostExec {
    int order;
    template <class F>
    void run(int n, const F& f) {
        if (n <= 0) return;
        if (order == 0) { for (int i = 0; i < n; i++) f(i); }
        else if (order == 1) { for (int i = n - 1; i >= 0; i--) f(i); }
        else {                                   // a fixed permutation: stride coprime with n
            long long stride = 7919; while (n % stride == 0 || stride % 2 == 0 && n % 2 == 0) stride += 2;
            auto gcd = [](long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; };
            while (gcd(stride, n) != 1) stride++;
            long long k = 12345 % n;
            for (int i = 0; i < n; i++) { f((int)k); k = (k + stride) % n; }
        }
    }
    int read_int(const int* p) { return *p; }
};
}  // namespace

extern "C" __attribute__((visibility("default")))
int lfblascheck_build(const float* bounds, int n, float tc, int nbins, int order, int bin_cap, float* out_nodes, int* out_indices, int* info) {
    if (n < 1 || nbins < 2 || nbins > kMaxBins) return -1;
    State S{};
    S.n = n; S.nbins = nbins; S.tc = tc; S.in_bounds = bounds;
    std::vector<float4> lo0(n), lo1(n), hi0(n), hi1(n);
    std::vector<int> no0(n), no1(n), flag(n + 1), scan(n + 1), chunk(chunk_capacity(n)), pairL(n), pairR(n), nflag(level_capacity(n)), nscan(level_capacity(n)), misc(4);
    std::vector<LevelNode> l0(level_capacity(n)), l1(level_capacity(n));
    std::vector<GNode> g(2 * (size_t)n);
    if (bin_cap < 1) bin_cap = 1;
    std::vector<float> bins((size_t)bin_cap * 3 * kBinFields * kMaxBins);
    S.lo[0] = lo0.data(); S.lo[1] = lo1.data(); S.hi[0] = hi0.data(); S.hi[1] = hi1.data();
    S.node_of[0] = no0.data(); S.node_of[1] = no1.data();
    S.flag = flag.data(); S.scan = scan.data(); S.chunk = chunk.data(); S.pairL = pairL.data(); S.pairR = pairR.data();
    S.lev[0] = l0.data(); S.lev[1] = l1.data(); S.nflag = nflag.data(); S.nscan = nscan.data();
    S.g = g.data(); S.bins = bins.data(); S.bin_cap = bin_cap; S.misc = misc.data();
    S.out_nodes = out_nodes; S.out_indices = out_indices;
    HostExec ex{order};
    Result R = build(ex, S);
    info[0] = R.num_nodes; info[1] = R.height; info[2] = R.negative_zero; info[3] = R.levels;
    return 0;
}
