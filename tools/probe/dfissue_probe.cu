// Probe 2 for the fused dual-path RNN: (A) what one thread can ISSUE -- cycles per tcgen05.mma (128 x N x 8, TF32) for the
// production issue pattern (descriptor rebuilt per MMA, commit per unit) against precomputed descriptors, and with two issuing
// threads; (B) cp.async.bulk global->shared throughput of one SM against unit size and the number of issuing threads.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../rtfs_net_b200/csrc/dprnn_fused.cuh"
using namespace rtfs;

// variant 0: production pattern, commit every `per` MMAs; 1: descriptors advanced by integer adds, one commit at the end;
// 2: as 1 from two threads (warps 0 and 2), half of the MMAs each, separate accumulators
template <int N>
__global__ void __launch_bounds__(128, 1) issue_probe(long long* out, int variant, int per, int nmma) {
    extern __shared__ __align__(128) unsigned char sm[];
    constexpr int ROWS = 7 + N + 7, LBO = ROWS * 16 + 16, HBUF = 16 * LBO;
    unsigned char* hbuf = sm;
    unsigned char* ring = sm + ((HBUF + 127) / 128) * 128;  // 64 KB of "weights"
    uint64_t* fin = reinterpret_cast<uint64_t*>(ring + 65536);
    uint32_t* slot = reinterpret_cast<uint32_t*>(fin + 4);
    const int tid = threadIdx.x;
    for (int i = tid; i < (HBUF + 65536) / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 0.f;
    if (tid == 32) {
        mbar_init(fin, 1);
        mbar_init(fin + 1, 1);
        mbar_init(fin + 2, 1);
        fence_mbar_init();
    }
    if (tid < 32) tmem_alloc<512>(slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    constexpr uint32_t IDESC = umma_idesc_tf32(128, N);
    const uint32_t hb = smem_u32(hbuf), rg = smem_u32(ring);
    if (variant == 0 && tid == 0) {
        const long long t0 = clock64();
        for (int g = 0; g < nmma / per; ++g) {
            const uint32_t st = rg + (g & 3) * 16384;
#pragma unroll 1
            for (int m = 0; m < per; ++m) {
                const uint64_t db = umma_desc(hb + ((m & 1) * 2) * LBO + (7 + (g & 7)) * 16, LBO, 128);
                umma_tf32(tmem + ((m >> 1) & 1) * N, umma_desc(st + (m & 3) * 4096, 2048, 128), db, IDESC, 1u);
            }
            umma_commit(fin + 2);  // nobody waits on it
        }
        const long long t1 = clock64();
        umma_commit(fin);
        mbar_wait(fin, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    if (variant >= 1 && (tid == 0 || (variant == 2 && tid == 64))) {
        const int who = tid == 0 ? 0 : 1;
        const int mine = variant == 2 ? nmma / 2 : nmma;
        uint64_t da = umma_desc(rg, 2048, 128), db = umma_desc(hb + 7 * 16, LBO, 128);
        const uint32_t acc = tmem + who * N;
        const long long t0 = clock64();
#pragma unroll 4
        for (int m = 0; m < mine; ++m) {
            umma_tf32(acc, da + (uint64_t)((m & 15) * 256), db + (uint64_t)(m & 7), IDESC, 1u);
        }
        const long long t1 = clock64();
        umma_commit(fin + who);
        mbar_wait(fin + who, 0);
        const long long t2 = clock64();
        out[2 * who] = t1 - t0;
        out[2 * who + 1] = t2 - t0;
    }
    // variant 3: the production pattern (descriptor rebuilt per MMA, commit every `per`) issued by the lane elect.sync picks inside
    // warp-uniform code; variant 4: as 1 under elect.sync
    if (variant >= 3 && tid < 32) {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        if (pred) {
            const long long t0 = clock64();
            if (variant == 3) {
                for (int g = 0; g < nmma / per; ++g) {
                    const uint32_t st = rg + (g & 3) * 16384;
#pragma unroll 1
                    for (int m = 0; m < per; ++m) {
                        const uint64_t db = umma_desc(hb + ((m & 1) * 2) * LBO + (7 + (g & 7)) * 16, LBO, 128);
                        umma_tf32(tmem + ((m >> 1) & 1) * N, umma_desc(st + (m & 3) * 4096, 2048, 128), db, IDESC, 1u);
                    }
                    umma_commit(fin + 2);
                }
            } else {
                uint64_t da = umma_desc(rg, 2048, 128), db = umma_desc(hb + 7 * 16, LBO, 128);
#pragma unroll 4
                for (int m = 0; m < nmma; ++m) umma_tf32(tmem, da + (uint64_t)((m & 15) * 256), db + (uint64_t)(m & 7), IDESC, 1u);
            }
            const long long t1 = clock64();
            umma_commit(fin);
            mbar_wait(fin, 0);
            const long long t2 = clock64();
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    __syncthreads();
    if (tid < 32) tmem_dealloc<512>(tmem);
}

template <int N>
void run_issue(int variant, int per, int nmma) {
    cudaFuncSetAttribute(issue_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long* out;
    cudaMalloc(&out, 64);
    cudaMemset(out, 0, 64);
    for (int i = 0; i < 2; ++i) issue_probe<N><<<1, 128, 200 * 1024>>>(out, variant, per, nmma);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4];
    cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
    printf("issue N=%3d variant %d commit/%d: %d MMAs issued in %6lld cycles (%5.1f per MMA), complete after %6lld (%5.1f per MMA)", N, variant, per, nmma, h[0],
           (double)h[0] / (variant == 2 ? nmma / 2 : nmma), h[1], (double)h[1] / nmma);
    if (variant == 2) printf(" | thread 2: issued %lld complete %lld", h[2], h[3]);
    printf(" %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out);
}

// P producer threads (lane 0 of warps 1..P), each with its own ring of D units of U bytes; a consumer thread per ring releases
// units as they land.  Reports bytes per clock of the SM.
DEVINL bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__global__ void __launch_bounds__(512, 1) bulk_probe(const float* w, long long* out, int P, int U, int D, int units_each, int elect) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 196608);  // [P][2][D]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < P * 2 * D; ++i) mbar_init(bars + i, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const long long t0 = clock64();
    if (elect && warp >= 1 && warp <= P) {  // producer: warp-uniform code, the elected lane issues
        const int p = warp - 1;
        uint64_t* full = bars + p * 2 * D;
        uint64_t* done = full + D;
        unsigned char* ring = sm + p * (196608 / P);
        for (int g = 0; g < units_each; ++g) {
            const int s = g % D;
            if (g >= D) mbar_wait(done + s, ((g / D) - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(full + s, U);
                bulk_g2s(ring + s * U, w + ((size_t)(g * P + p) * (U / 4)) % (131072), U, full + s);
            }
            __syncwarp();
        }
    }
    if (!elect && lane == 0 && warp >= 1 && warp <= P) {  // producer
        const int p = warp - 1;
        uint64_t* full = bars + p * 2 * D;
        uint64_t* done = full + D;
        unsigned char* ring = sm + p * (196608 / P);
        for (int g = 0; g < units_each; ++g) {
            const int s = g % D;
            if (g >= D) mbar_wait(done + s, ((g / D) - 1) & 1);
            mbar_expect_tx(full + s, U);
            bulk_g2s(ring + s * U, w + ((size_t)(g * P + p) * (U / 4)) % (131072), U, full + s);
        }
    }
    if (lane == 0 && warp >= 8 && warp < 8 + P) {  // consumer
        const int p = warp - 8;
        uint64_t* full = bars + p * 2 * D;
        uint64_t* done = full + D;
        for (int g = 0; g < units_each; ++g) {
            const int s = g % D;
            mbar_wait(full + s, (g / D) & 1);
            mbar_arrive(done + s);
        }
        out[blockIdx.x * 8 + p] = clock64() - t0;
    }
}

void run_bulk(const float* w, int grid, int P, int U, int D, int elect = 0) {
    cudaFuncSetAttribute(bulk_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long* out;
    cudaMalloc(&out, sizeof(long long) * grid * 8);
    const int total_bytes = 4 << 20;
    const int units_each = total_bytes / U / P;
    for (int i = 0; i < 2; ++i) bulk_probe<<<grid, 512, 200 * 1024>>>(w, out, P, U, D, units_each, elect);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(grid * 8);
    cudaMemcpy(h.data(), out, sizeof(long long) * grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int b = 0; b < grid; ++b)
        for (int p = 0; p < P; ++p) mx = h[b * 8 + p] > mx ? h[b * 8 + p] : mx;
    printf("bulk%s grid=%3d producers=%d unit=%5d depth=%d (%3d KB in flight): %7lld cycles for 4 MB -> %5.1f B/clk/SM, %5.0f cycles per copy %s\n", elect ? " (elect)" : "", grid, P, U, D,
           P * U * D / 1024, mx, (double)total_bytes / mx, (double)mx / units_each, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out);
}

int main(int argc, char** argv) {
    const int sel = argc > 1 ? atoi(argv[1]) : -1;  // -1: everything that is known to run; 3 / 4: the elect.sync issue variants; 5: elect bulk
    if (sel == 3) {
        run_issue<128>(3, 4, 128);
        return 0;
    }
    if (sel == 4) {
        run_issue<128>(4, 1, 128);
        return 0;
    }
    float* w;
    cudaMalloc(&w, 1 << 20);
    cudaMemset(w, 0, 1 << 20);
    if (sel == 5) {
        run_bulk(w, 1, 1, 16384, 4, 1);
        return 0;
    }
    for (int per : {4, 2, 1}) run_issue<256>(0, per, 128);
    for (int per : {4, 2, 1}) run_issue<128>(0, per, 128);
    run_issue<256>(1, 0, 128);
    run_issue<128>(1, 0, 128);
    run_issue<64>(1, 0, 128);
    run_issue<128>(2, 0, 128);
    run_issue<64>(2, 0, 128);
    for (int grid : {1, 148})
        for (int P : {1, 2, 4})
            for (int U : {4096, 8192, 16384, 32768}) {
                const int D = 196608 / P / U < 4 ? 196608 / P / U : 4;
                if (D >= 2) run_bulk(w, grid, P, U, D);
            }
    return 0;
}
