// Probe: achieved HBM bandwidth of a (M x 256) fp32 read [+ write] as a function of the per-warp access shape.
//   mode 0: warp instruction = 512 contiguous bytes of one row; a CTA walks its 128-row tile row by row
//   mode 1: warp instruction = 4 rows x 128 B; column block outer (the GEMM epilogue's pattern: each 1 KB row is
//           touched in 8 pieces spread over time)
//   mode 2: same 4 x 128 B shape, but all 8 column blocks of a row group back to back (row group outer)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(const float4* __restrict__ in, float4* __restrict__ out, int M, int mode, int write) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long row0 = (long long)blockIdx.x * 128;
    float4 acc = make_float4(0, 0, 0, 0);
    if (mode == 0) {
        for (int it = 0; it < 32; ++it) {  // 128 rows x 2 halves / 8 warps
            const int idx = it * 8 + warp;  // 0..255
            const long long row = row0 + (idx >> 1);
            const int c = (idx & 1) * 32 + lane;  // float4 index within the row (64 per row)
            if (row < M) {
                float4 v = in[row * 64 + c];
                if (write) out[row * 64 + c] = make_float4(v.x * 2, v.y * 2, v.z * 2, v.w * 2);
                else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
            }
        }
    } else {
        const int q = warp & 3, hlf = warp >> 2, rsub = lane >> 3, c4 = lane & 7;
        for (int a = 0; a < 4; ++a)
            for (int p = 0; p < 8; ++p) {
                const int cb = mode == 1 ? a : (p & 3), pp = mode == 1 ? p : (a * 2 + (p >> 2));
                const long long row = row0 + q * 32 + pp * 4 + rsub;
                const int c = (hlf * 128 + cb * 32) / 4 + c4;
                if (row < M) {
                    float4 v = in[row * 64 + c];
                    if (write) out[row * 64 + c] = make_float4(v.x * 2, v.y * 2, v.z * 2, v.w * 2);
                    else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
                }
            }
    }
    if (!write && acc.x == 12345.678f) out[0] = acc;
}
int main() {
    const int M = 32 * 251 * 129;
    float4 *in, *out;
    cudaMalloc(&in, (size_t)M * 1024);
    cudaMalloc(&out, (size_t)M * 1024);
    cudaMemset(in, 0, (size_t)M * 1024);
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    for (int write = 0; write < 2; ++write)
        for (int mode = 0; mode < 3; ++mode) {
            for (int i = 0; i < 2; ++i) k<<<(M + 127) / 128, 256>>>(in, out, M, mode, write);
            cudaEventRecord(s);
            for (int i = 0; i < 5; ++i) k<<<(M + 127) / 128, 256>>>(in, out, M, mode, write);
            cudaEventRecord(e);
            cudaEventSynchronize(e);
            float ms; cudaEventElapsedTime(&ms, s, e); ms /= 5;
            printf("write=%d mode=%d: %.3f ms  %.0f GB/s (%s)\n", write, mode, ms, (double)M * 1024 * (1 + write) / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
