// Probe 5: what bounds the persistent 256 -> 256 conv (audio bottleneck: reads A once, writes A once, streams the 256 KB
// weight image from L2 once per 128-row tile)?  The production kernel template is timed as is and with one ingredient removed:
//   full     GlnActLoader + StoreEpi4                       nostore  epilogue without the global stores
//   noload   producers without the global loads             (build with -DRTFS_PROBE_W_ONCE: weight slabs copied once only)
#include <cstdio>
#include <vector>
#include "../../rtfs_net_b200/csrc/gemm_tcp.cuh"
using namespace rtfs;
#ifndef REPS
#define REPS 5
#endif
#ifndef WARM
#define WARM 2
#endif

struct NullStoreEpi4 {
    float* C;
    using Pre = NoPre;
    DEVINL void init(int, int) {}
    DEVINL void prep(int) {}
    DEVINL Pre load(int, int) const { return Pre{}; }
    DEVINL void store4(int row, int col, float4 v, const Pre&) {
        if (v.x == 12345.678f) C[0] = v.x + row + col;
    }
    DEVINL void finish(float*) {}
    DEVINL void finish_group(float*, int, int, int) {}
};
struct NoLoadLoader {  // same interface as GlnActLoader's raw()/xform(), always the same 4 KB of A (L1/L2 resident)
    const float* A;
    static constexpr int kExtra = 0;
    DEVINL void init_p(int, int, float*, int, int) {}
    DEVINL const float* raw(long long row, int k) const { return A + (row & 3) * 256 + k; }
    DEVINL float4 xform(float4 x, int, int) const { return x; }
    DEVINL float4 load(int, int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
};

template <class F>
void timeit(const char* name, double bytes, F f) {
    cudaEvent_t s, e;
    cudaEventCreate(&s);
    cudaEventCreate(&e);
    for (int i = 0; i < WARM; ++i) f();
    cudaEventRecord(s);
    for (int i = 0; i < REPS; ++i) f();
    cudaEventRecord(e);
    cudaEventSynchronize(e);
    float ms;
    cudaEventElapsedTime(&ms, s, e);
    ms /= REPS;
    cudaDeviceSynchronize();
    printf("%-28s %.3f ms  %.0f GB/s (%s)\n", name, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int B = 32, P = 251 * 129, M = B * P;
    float *A, *C, *W, *g, *b;
    double* sums;
    cudaMalloc(&A, (size_t)M * 256 * 4);
    cudaMalloc(&C, (size_t)M * 256 * 4);
    cudaMalloc(&W, 256 * 256 * 4);
    cudaMalloc(&g, 1024);
    cudaMalloc(&b, 1024);
    cudaMalloc(&sums, B * 16);
    cudaMemset(A, 0, (size_t)M * 256 * 4);
    cudaMemset(W, 0, 256 * 256 * 4);
    cudaMemset(g, 0, 1024);
    cudaMemset(b, 0, 1024);
    std::vector<double> h(2 * B, 1.0);
    cudaMemcpy(sums, h.data(), B * 16, cudaMemcpyHostToDevice);
    GlnActLoader<256, 1> al{A, GlnRef{sums, g, b, 1.0 / (double)(P * 256)}, P, B};
    StoreEpi4 ep{C, 256, nullptr};
    NullStoreEpi4 ne{C};
    NoLoadLoader nl{A};
    const double a = (double)M * 1024;
#define CFG(PF_, NP_, NSA_)                                                                                                              \
    timeit("full    PF=" #PF_ " NPROD=" #NP_ " NSA=" #NSA_, 2 * a, [&] { launch_gemm_tcp<256, 256, NSA_, 3, false, PF_, 2, NP_>(al, W, ep, M, 0); }); \
    timeit("nostore PF=" #PF_ " NPROD=" #NP_ " NSA=" #NSA_, a, [&] { launch_gemm_tcp<256, 256, NSA_, 3, false, PF_, 2, NP_>(al, W, ne, M, 0); });
    CFG(4, 512, 4)
    CFG(2, 512, 4)
    CFG(1, 512, 4)
    CFG(4, 256, 4)
    CFG(2, 256, 4)
    CFG(4, 512, 3)
    timeit("noload (write A)", a, [&] { launch_gemm_tcp<256, 256, 4, 3, false, 4, 2, 512>(nl, W, ep, M, 0); });
    timeit("neither (MMA + W stream)", a, [&] { launch_gemm_tcp<256, 256, 4, 3, false, 4, 2, 512>(nl, W, ne, M, 0); });
    return 0;
}
