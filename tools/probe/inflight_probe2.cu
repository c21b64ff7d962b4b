// Probe 2: HBM read bandwidth with ONE CTA per SM as a function of the CTA's warp count (NT threads) and loads in flight.
#include <cstdio>
#include <cuda_runtime.h>
template <int U, int NT>
__global__ void __launch_bounds__(NT) k(const float4* __restrict__ in, float4* __restrict__ out, int ntiles) {
    float4 acc = make_float4(0, 0, 0, 0);
    constexpr int PER = 8192 / NT;  // float4 per thread per 128 KB tile
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const float4* p = in + (long long)tile * 8192 + threadIdx.x;
#pragma unroll 1
        for (int j = 0; j < PER; j += U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = p[(j + u) * NT];
#pragma unroll
            for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
    }
    if (acc.x == 12345.678f) out[0] = acc;
}
template <int U, int NT>
void run(const float4* in, float4* out, int ntiles) {
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k<U, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    for (int i = 0; i < 2; ++i) k<U, NT><<<148, NT, smem>>>(in, out, ntiles);
    cudaEventRecord(s);
    for (int i = 0; i < 5; ++i) k<U, NT><<<148, NT, smem>>>(in, out, ntiles);
    cudaEventRecord(e);
    cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e); ms /= 5;
    printf("1 CTA/SM, %4d threads, U=%2d (in flight/SM %4d KB): %.3f ms  %.0f GB/s (%s)\n", NT, U, U * NT * 16 / 1024, ms,
           (double)ntiles * 131072 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int ntiles = 8096;
    float4 *in, *out;
    cudaMalloc(&in, (size_t)ntiles * 131072);
    cudaMalloc(&out, 1024);
    cudaMemset(in, 0, (size_t)ntiles * 131072);
    run<4, 256>(in, out, ntiles); run<16, 256>(in, out, ntiles);
    run<4, 512>(in, out, ntiles); run<8, 512>(in, out, ntiles); run<16, 512>(in, out, ntiles);
    run<2, 1024>(in, out, ntiles); run<4, 1024>(in, out, ntiles); run<8, 1024>(in, out, ntiles);
    return 0;
}
