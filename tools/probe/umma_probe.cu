// Probe: how does tcgen05.mma address a K-major SWIZZLE_NONE operand whose descriptor start is advanced by
// t*16 bytes (t rows inside an 8-row core matrix)?  D[r][0] reports which slab row the tensor core read for row r.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../rtfs_net_b200/csrc/gemm_tc.cuh"
using namespace rtfs;

constexpr int ROWS = 144, LBO = ROWS * 16 + 16;

__global__ void probe(float* out, int shift, int mode) {
    __shared__ __align__(128) unsigned char slab[2 * LBO + 256];
    __shared__ __align__(128) float bmat[2 * 16 * 4];  // [kq][n=16][4]
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc<32>(&slot);
    if (tid == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
    for (int i = tid; i < 2 * ROWS; i += blockDim.x) {
        const int kq = i / ROWS, r = i % ROWS;
        float4 v = make_float4(kq == 0 ? (float)r : 1000.f + r, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(slab + kq * LBO + r * 16) = v;
    }
    for (int i = tid; i < 2 * 16 * 4; i += blockDim.x) bmat[i] = 0.f;
    __syncthreads();
    if (tid == 0) { bmat[(0 * 16 + 0) * 4 + 0] = 1.f; bmat[(1 * 16 + 1) * 4 + 0] = 1.f; }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        uint32_t a = smem_u32(slab);
        uint64_t da;
        if (mode == 0) da = umma_desc(a + shift * 16, LBO, 128);
        else da = umma_desc(a + shift * 16, LBO, 128) | ((uint64_t)(shift & 7) << 49);  // base_offset field
        const uint64_t db = umma_desc(smem_u32(bmat), 16 * 16, 128);
        umma_tf32(tmem, da, db, umma_idesc_tf32(128, 16), 0u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);  // 32 columns (only 16 valid, alloc is 32)
    out[(warp * 32 + lane) * 2 + 0] = __uint_as_float(v[0]);
    out[(warp * 32 + lane) * 2 + 1] = __uint_as_float(v[1]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<32>(tmem);
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * 2 * 4);
    float h[256];
    for (int mode = 0; mode < 2; ++mode)
        for (int shift = 0; shift <= 9; ++shift) {
            probe<<<1, 128>>>(d, shift, mode);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("mode %d shift %d (%s): rows read for r=0..17 kq0:", mode, shift, cudaGetErrorString(e));
            for (int r = 0; r < 18; ++r) printf(" %g", h[2 * r]);
            printf(" | kq1:");
            for (int r = 0; r < 10; ++r) printf(" %g", h[2 * r + 1] - 1000.f);
            printf(" | r=120..127:");
            for (int r = 120; r < 128; ++r) printf(" %g", h[2 * r]);
            printf("\n");
        }
    return 0;
}
