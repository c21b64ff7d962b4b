// Probe 3: one persistent 512-thread CTA per SM streams 128-row x 1 KB tiles in 32-float K chunks (the GEMM producers'
// access shape: thread = (row r0 + 64 i, 16-byte piece kq)), either with cp.async into a shared-memory ring
// (wait_group, D chunks in flight) or with plain loads into registers (D chunks in flight).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int N> __device__ void cpwait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }
template <int D, bool ASYNC>
__global__ void __launch_bounds__(512) k(const float* __restrict__ in, float* __restrict__ out, int ntiles) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int tid = threadIdx.x, kq = tid & 7, r0 = tid >> 3;
    const int my = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x, G = my * 8;
    float4 acc = make_float4(0, 0, 0, 0);
    if (ASYNC) {
        auto issue = [&](int g) {
            if (g < G) {
                const int it = g >> 3, kc = g & 7;
                const long long row = (long long)(blockIdx.x + it * gridDim.x) * 128 + r0;
                unsigned char* dst = sm + (g % (D + 2)) * 16384 + tid * 32;
                for (int i = 0; i < 2; ++i) {
                    unsigned s = (unsigned)__cvta_generic_to_shared(dst + i * 16);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(in + (row + 64 * i) * 256 + kc * 32 + kq * 4));
                }
            }
            asm volatile("cp.async.commit_group;");
        };
        for (int g = 0; g < D; ++g) issue(g);
        for (int g = 0; g < G; ++g) {
            issue(g + D);
            cpwait<D>();
            const float4* p = reinterpret_cast<const float4*>(sm + (g % (D + 2)) * 16384 + tid * 32);
            acc.x += p[0].x + p[1].x;
        }
    } else {
        float4 v[D][2];
        auto ld = [&](int g, float4 (&d)[2]) {
            const int it = g >> 3, kc = g & 7;
            const long long row = (long long)(blockIdx.x + it * gridDim.x) * 128 + r0;
            for (int i = 0; i < 2; ++i) d[i] = *reinterpret_cast<const float4*>(in + (row + 64 * i) * 256 + kc * 32 + kq * 4);
        };
#pragma unroll
        for (int g = 0; g < D; ++g) ld(g, v[g]);
        for (int g0 = 0; g0 < G; g0 += D) {
#pragma unroll
            for (int j = 0; j < D; ++j) {
                acc.x += v[j][0].x + v[j][1].x;
                if (g0 + j + D < G) ld(g0 + j + D, v[j]);
            }
        }
    }
    if (acc.x == 12345.678f) out[0] = acc.x;
}
template <int D, bool ASYNC>
void run(const float* in, float* out, int ntiles) {
    const int smem = ASYNC ? (D + 2) * 16384 : 0;
    cudaFuncSetAttribute(k<D, ASYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    for (int i = 0; i < 2; ++i) k<D, ASYNC><<<148, 512, 200 * 1024>>>(in, out, ntiles);
    cudaEventRecord(s);
    for (int i = 0; i < 5; ++i) k<D, ASYNC><<<148, 512, 200 * 1024>>>(in, out, ntiles);
    cudaEventRecord(e);
    cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e); ms /= 5;
    printf("%s D=%d chunks in flight (%3d KB/SM): %.3f ms  %.0f GB/s (%s)\n", ASYNC ? "cp.async" : "ld->reg ", D, D * 16, ms,
           (double)ntiles * 131072 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    (void)smem;
}
int main() {
    const int ntiles = 8096;
    float *in, *out;
    cudaMalloc(&in, (size_t)ntiles * 131072);
    cudaMalloc(&out, 1024);
    cudaMemset(in, 0, (size_t)ntiles * 131072);
    { unsigned* hbuf = (unsigned*)malloc(64u << 20); unsigned s = 12345u; for (size_t i = 0; i < (16u << 20); ++i) { s = s * 1664525u + 1013904223u; hbuf[i] = (s >> 9) | 0x3f800000u; } for (size_t off = 0; off < (size_t)ntiles * 131072; off += (64u << 20)) { size_t n = (size_t)ntiles * 131072 - off; if (n > (64u << 20)) n = 64u << 20; cudaMemcpy((char*)in + off, hbuf, n, cudaMemcpyHostToDevice); } free(hbuf); }
    run<2, true>(in, out, ntiles); run<4, true>(in, out, ntiles); run<8, true>(in, out, ntiles);
    run<2, false>(in, out, ntiles); run<4, false>(in, out, ntiles); run<8, false>(in, out, ntiles);
    return 0;
}
