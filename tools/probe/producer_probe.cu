// Probe 4: which ingredient of the GEMM producer loop costs HBM bandwidth?  One persistent CTA per SM, 16 producer warps
// (+1 consumer warp), register prefetch of D=4 chunks (the ld->reg D=4 case of cpasync_probe.cu = ~5.65 TB/s), plus:
//   mode 0: nothing else (baseline)                         mode 1: + STS.128 of the chunk into a 4-stage smem ring
//   mode 2: mode 1 + fence.proxy.async per chunk            mode 3: mode 2 + full/empty mbarrier hand-shake with a consumer thread
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(su32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(su32(b)), "r"(par) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(544) k(const float* __restrict__ in, float* __restrict__ out, int ntiles) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + 4 * 16640);
    uint64_t* empty = full + 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { for (int s = 0; s < 4; ++s) { mbar_init(full + s, 16); mbar_init(empty + s, 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int my = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x, G = my * 8;
    if (warp < 16) {
        const int kq = tid & 7, r0 = tid >> 3;
        unsigned char* dst0 = sm + kq * 2064 + r0 * 16;
        constexpr int D = 4;
        float4 v[D][2];
        float accx = 0.f;
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
                const int g = g0 + j, s = g & 3;
                if (MODE >= 3 && g >= 4) mbar_wait(empty + s, ((g >> 2) - 1) & 1);
                if (MODE >= 1) {
                    for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(dst0 + s * 16640 + i * 1024) = v[j][i];
                } else accx += v[j][0].x + v[j][1].x;
                if (MODE >= 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (MODE >= 3) { __syncwarp(); if (lane == 0) mbar_arrive(full + s); }
                if (g + D < G) ld(g + D, v[j]);
            }
        }
        if (accx == 12345.678f) out[0] = accx;
    } else if (MODE >= 3 && lane == 0) {
        for (int g = 0; g < G; ++g) {
            const int s = g & 3;
            mbar_wait(full + s, (g >> 2) & 1);
            mbar_arrive(empty + s);
        }
    }
}
template <int MODE>
void run(const float* in, float* out, int ntiles) {
    const int smem = 4 * 16640 + 128;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    for (int i = 0; i < 2; ++i) k<MODE><<<148, 544, 200 * 1024>>>(in, out, ntiles);
    cudaEventRecord(s);
    for (int i = 0; i < 5; ++i) k<MODE><<<148, 544, 200 * 1024>>>(in, out, ntiles);
    cudaEventRecord(e);
    cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e); ms /= 5;
    printf("mode %d: %.3f ms  %.0f GB/s (%s)\n", MODE, ms, (double)ntiles * 131072 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    (void)smem;
}
int main() {
    const int ntiles = 8096;
    float *in, *out;
    cudaMalloc(&in, (size_t)ntiles * 131072);
    cudaMalloc(&out, 1024);
    cudaMemset(in, 0, (size_t)ntiles * 131072);
    run<0>(in, out, ntiles); run<1>(in, out, ntiles); run<2>(in, out, ntiles); run<3>(in, out, ntiles);
    return 0;
}
