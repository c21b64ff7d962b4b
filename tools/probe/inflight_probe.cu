// Probe: HBM read bandwidth vs bytes in flight per SM.  Each 256-thread CTA reads its 128 KB tile with U independent
// 16-byte loads per thread in flight (then repeats for the next tile, persistent); occupancy is limited with dynamic smem.
#include <cstdio>
#include <cuda_runtime.h>
template <int U>
__global__ void __launch_bounds__(256) k(const float4* __restrict__ in, float4* __restrict__ out, int ntiles) {
    float4 acc = make_float4(0, 0, 0, 0);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const float4* p = in + (long long)tile * 8192 + threadIdx.x;
#pragma unroll 1
        for (int j = 0; j < 32; j += U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = p[(j + u) * 256];
#pragma unroll
            for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
    }
    if (acc.x == 12345.678f) out[0] = acc;
}
template <int U>
void run(const float4* in, float4* out, int ntiles, int ctas_per_sm) {
    const int smem = ctas_per_sm >= 8 ? 0 : (200 * 1024 / ctas_per_sm);
    cudaFuncSetAttribute(k<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    const int grid = 148 * ctas_per_sm;
    for (int i = 0; i < 2; ++i) k<U><<<grid, 256, smem>>>(in, out, ntiles);
    cudaEventRecord(s);
    for (int i = 0; i < 5; ++i) k<U><<<grid, 256, smem>>>(in, out, ntiles);
    cudaEventRecord(e);
    cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e); ms /= 5;
    printf("U=%2d ctas/SM=%d  in flight/SM = %4d KB : %.3f ms  %.0f GB/s (%s)\n", U, ctas_per_sm, U * 4 * ctas_per_sm, ms,
           (double)ntiles * 131072 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int ntiles = 8096;
    float4 *in, *out;
    cudaMalloc(&in, (size_t)ntiles * 131072);
    cudaMalloc(&out, 1024);
    cudaMemset(in, 0, (size_t)ntiles * 131072);
    for (int c : {1, 2, 4, 8}) { run<4>(in, out, ntiles, c); run<8>(in, out, ntiles, c); run<16>(in, out, ntiles, c); run<32>(in, out, ntiles, c); }
    return 0;
}
