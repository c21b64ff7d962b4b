// Probe for the fused dual-path RNN's layer-0 GEMM (dprnn_fused.cuh): what bounds it -- the tensor pipe or the weight stream?
//   mode 0: UMMA only   (operands resident, NMMA tcgen05.mma of 128 x N x 8 TF32 issued back to back, committed every `per` MMAs)
//   mode 1: stream only (the 512 KB layer-0 weight image through a D-stage ring of U-byte units, cp.async.bulk, nothing consumes)
//   mode 2: both        (the production loop: wait unit, issue its MMAs, commit to the unit's barrier)
// Every CTA (one per SM, `grid` of them) reports clock64 cycles for ONE pass (32 slabs of 16 KB = one layer-0 GEMM).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../rtfs_net_b200/csrc/dprnn_fused.cuh"
using namespace rtfs;

template <int N, int U, int D>
__global__ void __launch_bounds__(256, 1) probe(const float* wimg, long long* out, int mode, int reps, int el, int nprod, int cper) {
    extern __shared__ __align__(128) unsigned char sm[];
    constexpr int ROWS = 7 + N + 7, LBO = ROWS * 16 + 16, HBUF = 16 * LBO;
    unsigned char* hbuf = sm;
    unsigned char* ring = sm + ((HBUF + 127) / 128) * 128;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + D * U);
    uint64_t* done = full + D;
    uint64_t* fin = done + D;
    uint32_t* slot = reinterpret_cast<uint32_t*>(fin + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < (HBUF + D * U) / 4; i += 256) reinterpret_cast<float*>(sm)[i] = 0.f;
    if (tid == 32) {
        for (int s = 0; s < D; ++s) {
            mbar_init(full + s, 1);
            mbar_init(done + s, 1);
        }
        mbar_init(fin, 1);
        fence_mbar_init();
    }
    if (tid < 32) tmem_alloc<512>(slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    constexpr int UNITS = 32 * 16384 / U;       // units per pass
    constexpr int MPU = (U / 4096) * 1;          // MMAs per unit: 4 KB of weights (128 features x 8 K) per MMA
    constexpr uint32_t IDESC = umma_idesc_tf32(128, N);
    const uint32_t hb = smem_u32(hbuf), rg = smem_u32(ring);
    const int total = UNITS * reps;
    long long t0 = 0, t1 = 0;
    // producers: lane 0 of warps 1..nprod, units g = p mod nprod (D % nprod == 0: a slot is always refilled by the same thread)
    if ((tid & 31) == 0 && (tid >> 5) >= 1 && (tid >> 5) <= nprod && mode != 0) {
        for (int g = (tid >> 5) - 1; g < total; g += nprod) {
            const int s = g % D;
            if (g >= D) mbar_wait(done + s, ((g / D) - 1) & 1);
            mbar_expect_tx(full + s, U);
            bulk_g2s(ring + s * U, wimg + (size_t)(g % UNITS) * (U / 4), U, full + s);
        }
    }
    // el = 1: issue from the elect.sync lane of warp 0 (uniform-register UTCHMMA / UTCBAR); cper = units per mma_done commit
    bool issuer = tid == 0;
    if (el && tid < 32) issuer = elect_one();
    if (tid < 32 && el ? issuer : tid == 0) {
        t0 = clock64();
        for (int g = 0; g < total; ++g) {
            const int s = g % D;
            if (mode != 0) {
                mbar_wait(full + s, (g / D) & 1);
                tc_fence_after();
            }
            if (mode != 1) {
#pragma unroll
                for (int m = 0; m < MPU; ++m) {
                    const uint64_t db = umma_desc(hb + ((m & 1) * 2) * LBO + 7 * 16, LBO, 128);
                    umma_tf32(tmem + ((m >> 1) & 1) * N, umma_desc(rg + s * U + (m % (U / 4096)) * 4096, 2048, 128), db, IDESC, 1u);
                }
                if (cper == 1) umma_commit(done + s);
                else if ((g % cper) == cper - 1)
                    for (int j = 0; j < cper; ++j) umma_commit(done + (g - j) % D);
            } else {
                mbar_arrive(done + s);
            }
        }
        if (mode != 1) {
            umma_commit(fin);
            mbar_wait(fin, 0);
        }
        t1 = clock64();
        out[blockIdx.x] = (t1 - t0) / reps;
    }
    __syncthreads();
    if (tid < 32) tmem_dealloc<512>(tmem);
}


template <int N, int U, int D>
void run(const float* w, int grid, int mode, int reps, const char* what, int el = 0, int nprod = 1, int cper = 1) {
    constexpr int ROWS = 7 + N + 7, LBO = ROWS * 16 + 16, HBUF = 16 * LBO;
    const int smem = ((HBUF + 127) / 128) * 128 + D * U + (2 * D + 1) * 8 + 64;
    cudaFuncSetAttribute(probe<N, U, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long* out;
    cudaMalloc(&out, sizeof(long long) * grid);
    for (int i = 0; i < 2; ++i) probe<N, U, D><<<grid, 256, 200 * 1024>>>(w, out, mode, reps, el, nprod, cper);  // 200 KB: one CTA per SM
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1ll << 60, sum = 0;
    for (long long v : h) {
        mx = v > mx ? v : mx;
        mn = v < mn ? v : mn;
        sum += v;
    }
    printf("%s prod=%d commit/%d N=%3d unit=%5d B depth=%d (%3d KB in flight) grid=%3d %-11s: cycles per 512 KB pass  min %6lld  avg %6lld  max %6lld  -> %5.1f B/clk/SM  %s (smem %d)\n", el ? "elect" : "tid0 ", nprod, cper, N, U, D,
           D * U / 1024, grid, what, mn, sum / grid, mx, 524288.0 / (double)(sum / grid), e == cudaSuccess ? "" : cudaGetErrorString(e), smem);
    cudaFree(out);
}

int main(int argc, char** argv) {
    float* w;
    cudaMalloc(&w, 32 * 16384);
    cudaMemset(w, 0, 32 * 16384);
    const char* names[3] = {"UMMA only", "stream only", "both"};
    if (argc > 1) {  // round-2 questions: elect.sync issue, several producers, commit granularity (grid 148)
        for (int mode : {0, 2}) {
            run<128, 8192, 8>(w, 148, mode, 4, names[mode], 0, 4, 1);
            run<128, 8192, 8>(w, 148, mode, 4, names[mode], 1, 4, 1);
            run<128, 8192, 8>(w, 148, mode, 4, names[mode], 1, 4, 2);
            run<128, 8192, 8>(w, 148, mode, 4, names[mode], 1, 4, 4);
            run<128, 8192, 4>(w, 148, mode, 4, names[mode], 1, 4, 1);
            run<128, 8192, 16>(w, 148, mode, 4, names[mode], 1, 4, 1);
            run<128, 16384, 4>(w, 148, mode, 4, names[mode], 1, 4, 1);
            run<128, 16384, 8>(w, 148, mode, 4, names[mode], 1, 4, 1);
            run<256, 16384, 4>(w, 148, mode, 4, names[mode], 1, 4, 1);
            run<256, 16384, 8>(w, 148, mode, 4, names[mode], 1, 4, 1);
        }
        run<128, 8192, 8>(w, 148, 1, 4, names[1], 1, 4, 1);
        run<128, 8192, 16>(w, 148, 1, 4, names[1], 1, 4, 1);
        return 0;
    }
    for (int grid : {1, 148}) {
        for (int mode = 0; mode < 3; ++mode) {
            run<256, 16384, 5>(w, grid, mode, 4, names[mode]);
            run<128, 16384, 5>(w, grid, mode, 4, names[mode]);
            run<128, 8192, 5>(w, grid, mode, 4, names[mode]);
            run<128, 8192, 10>(w, grid, mode, 4, names[mode]);
            run<128, 4096, 10>(w, grid, mode, 4, names[mode]);
            run<128, 16384, 8>(w, grid, mode, 4, names[mode]);
        }
    }
    return 0;
}
