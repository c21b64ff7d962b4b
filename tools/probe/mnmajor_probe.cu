// Probe: which shared-memory word does tcgen05.mma read for B[k][n] under an MN-major SWIZZLE_128B descriptor?
// B area word i holds i (split into low / high 10 bits over two runs: tf32 keeps 10 mantissa bits).  A = unit row e_k0.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../rtfs_net_b200/csrc/att_core_tc.cuh"
using namespace rtfs;
constexpr int LBOA = 128 * 16 + 16;
__global__ void probe(float* out, int k0, int part, uint32_t lbo, uint32_t sbo, int bmajor, int N, int ltype) {
    extern __shared__ unsigned char dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn) + 1023) & ~(uintptr_t)1023);
    float* bsm = reinterpret_cast<float*>(base);             // 64 KB
    unsigned char* asl = base + 65536;                       // 2 x LBOA
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc<256>(&slot);
    if (tid == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
    for (int i = tid; i < 16384; i += blockDim.x) bsm[i] = part == 0 ? (float)(i & 1023) : (float)(i >> 10);
    for (int i = tid; i < 2 * LBOA / 4; i += blockDim.x) reinterpret_cast<float*>(asl)[i] = 0.f;
    __syncthreads();
    if (tid == 0) *reinterpret_cast<float*>(asl + (k0 / 4) * LBOA + 0 * 16 + (k0 % 4) * 4) = 1.f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint64_t da = umma_desc(smem_u32(asl), LBOA, 128);
        const uint64_t db = (umma_desc_sw128(smem_u32(bsm), lbo, sbo) & ~(7ull << 61)) | ((uint64_t)ltype << 61);
        const uint32_t idesc = umma_idesc_tf32(128, N) | (bmajor ? (1u << 16) : 0u);
        umma_tf32(tmem, da, db, idesc, 0u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (warp == 0) {
        for (int cb = 0; cb < N / 32; ++cb) {
            uint32_t v[32];
            tmem_ld32(tmem + cb * 32, v);
            if (tid == 0)
                for (int i = 0; i < 32; ++i) out[cb * 32 + i] = __uint_as_float(v[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}
int main() {
    float* d;
    cudaMalloc(&d, 256 * 4);
    float lo[256], hi[256];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 72000);
    const int N = 128;
    struct Cfg { uint32_t lbo, sbo; int bm, lt; } cfgs[] = {{4096, 512, 1, 1}, {512, 4096, 1, 1}, {4096, 1024, 1, 0}};
    for (auto c : cfgs)
        for (int k0 = 0; k0 < 8; k0 += 1) {
            probe<<<1, 128, 72000>>>(d, k0, 0, c.lbo, c.sbo, c.bm, N, c.lt);
            cudaMemcpy(lo, d, sizeof(lo), cudaMemcpyDeviceToHost);
            probe<<<1, 128, 72000>>>(d, k0, 1, c.lbo, c.sbo, c.bm, N, c.lt);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(hi, d, sizeof(hi), cudaMemcpyDeviceToHost);
            printf("ltype %d bmajor %d lbo %u sbo %u k0 %d (%s): word read for n=0..:", c.lt, c.bm, c.lbo, c.sbo, k0, cudaGetErrorString(e));
            for (int n = 0; n < N; ++n) if (n < 12 || (n % 8 == 0)) printf(" [%d]%d", n, (int)hi[n] * 1024 + (int)lo[n]);
            printf("\n");
        }
    return 0;
}
