// Gateway + projection of the RTFS block (tdanet.py:34-49,107-108) as a persistent 1024-thread tcgen05 kernel:
//     r = PReLU(wg[c]*x + bg[c])  (never stored) ;  p_pre = Wp . r + bp  (256 -> 64) ;  per-sample (sum, sumsq) of p_pre
// The stage reads a (B*T*F) x 256 fp32 tensor once and writes a quarter of that: it is a pure HBM read stream.
// Measured on this part (tools/probe/inflight_probe2.cu) the read bandwidth follows the number of warps that issue
// loads (8 warps/SM ~4 TB/s, 32 warps ~6.1 TB/s), so the schedule is: ONE CTA per SM, 32 warps, each thread owns
// 8 float4 of the 128-row tile and keeps the NEXT tile's 8 loads in flight (registers) while the current tile is
// transformed into the UMMA K-major slab, multiplied (32 x tcgen05.mma M=128 N=64 K=8 by one thread, two TMEM
// accumulators ping-pong) and the previous tile's accumulator is drained by warps 0-3 (transposed through shared
// memory -> coalesced float4 stores + gLN statistics).  The 64 KB weight image stays resident in shared memory.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace rtfs {

constexpr int GW_NT = 1024;
constexpr int GW_A_BYTES = 64 * TC_LBO_A;                 // 64 K-pieces x (128 rows x 16 B + pad) = 132096
constexpr int GW_W_BYTES = 64 * 256 * 4;                  // 65536
constexpr int GW_STG_BYTES = 4 * 32 * TC_STG_LD * 4;      // 4 epilogue warps x [32][36] floats = 18432
constexpr int GW_SMEM = GW_A_BYTES + GW_W_BYTES + GW_STG_BYTES + 256;

struct GwArgs {
    const float* x;      // [M][256]
    const float* wg;     // gateway dw1x1 [256]
    const float* bg;
    const float* slope;  // PReLU [1]
    const float* wimg;   // projection weight image [64 pieces][64][4] (tf32)
    const float* bias;   // [64]
    float* out;          // [M][64]
    double* sums;        // [B][2]
    int M, P, B, ntiles;
};

__global__ void __launch_bounds__(GW_NT, 1) gateproj_wide_kernel(GwArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* a_slab = smem_raw;
    unsigned char* w_slab = smem_raw + GW_A_BYTES;
    float* stg_all = reinterpret_cast<float*>(w_slab + GW_W_BYTES);
    uint64_t* w_ready = reinterpret_cast<uint64_t*>(smem_raw + GW_A_BYTES + GW_W_BYTES + GW_STG_BYTES);
    uint64_t* mma_done = w_ready + 1;  // [2] one per accumulator
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 2);
    float* scratch = reinterpret_cast<float*>(tmem_slot + 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) tmem_alloc<128>(tmem_slot);
    if (tid == 32) {
        mbar_init(w_ready, 1);
        mbar_init(mma_done, 1);
        mbar_init(mma_done + 1, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        mbar_expect_tx(w_ready, GW_W_BYTES);
        bulk_g2s(w_slab, a.wimg, GW_W_BYTES, w_ready);
    }

    // thread -> (K piece kq of 4 channels, rows r0 + 16 i): a warp reads 512 contiguous bytes of one row
    const int kq = tid & 63, r0 = tid >> 6;
    const float4 wgv = ldg4(a.wg + 4 * kq), bgv = ldg4(a.bg + 4 * kq);
    const float slope = __ldg(a.slope);
    unsigned char* a_dst = a_slab + kq * TC_LBO_A + r0 * 16;
    constexpr uint32_t IDESC = umma_idesc_tf32(128, 64);

    float4 v[8];
    auto load_tile = [&](int tile) {
        const long long rbase = (long long)tile * 128 + r0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long row = rbase + 16 * i;
            v[i] = row < a.M ? ldg4(a.x + row * 256 + 4 * kq) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if ((int)blockIdx.x < a.ntiles) load_tile(blockIdx.x);

    int it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles + (int)gridDim.x; tile += gridDim.x, ++it) {
        const bool have = tile < a.ntiles;       // a tile to multiply this round
        const bool drain = it > 0;               // the previous round's accumulator to write out
        if (!have && !drain) break;
        if (have) {
            // the MMAs of tile it-1 (the last readers of the slab) have completed
            if (it > 0) mbar_wait(mma_done + ((it - 1) & 1), ((it - 1) >> 1) & 1);
            const long long rbase = (long long)tile * 128 + r0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 y;
                y.x = tf32r(prelu(fmaf(wgv.x, v[i].x, bgv.x), slope));
                y.y = tf32r(prelu(fmaf(wgv.y, v[i].y, bgv.y), slope));
                y.z = tf32r(prelu(fmaf(wgv.z, v[i].z, bgv.z), slope));
                y.w = tf32r(prelu(fmaf(wgv.w, v[i].w, bgv.w), slope));
                if (rbase + 16 * i >= a.M) y = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(a_dst + i * (16 * 16)) = y;
            }
            if (tile + (int)gridDim.x < a.ntiles) load_tile(tile + gridDim.x);  // next tile's loads fly during MMA + drain
            fence_proxy_async();
        } else if (it > 0) {
            mbar_wait(mma_done + ((it - 1) & 1), ((it - 1) >> 1) & 1);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (have && tid == 0) {
            if (it == 0) mbar_wait(w_ready, 0);
            const uint32_t ab = smem_u32(a_slab), wb = smem_u32(w_slab);
#pragma unroll 8
            for (int k8 = 0; k8 < 32; ++k8)
                umma_tf32(tmem + (it & 1) * 64, umma_desc(ab + 2 * k8 * TC_LBO_A, TC_LBO_A, 128), umma_desc(wb + 2 * k8 * 1024, 1024, 128), IDESC,
                          k8 > 0 ? 1u : 0u);
            umma_commit(mma_done + (it & 1));
        }
        if (drain && warp < 4) {
            // ---- write out tile it-1 (its MMAs were waited for above): warps 0-3 = TMEM lane quarters
            const int ptile = tile - gridDim.x, prow0 = ptile * 128;
            const int bfirst = prow0 / a.P, split = (bfirst + 1) * a.P;
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
            float* stg = stg_all + warp * (32 * TC_STG_LD);
            const int rsub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {  // 16 columns at a time: the next tile's 8 loads stay live in registers
                    uint32_t t16[16];
                    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(((it - 1) & 1) * 64 + cb * 32 + hh * 16), t16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<float4*>(stg + lane * TC_STG_LD + 16 * hh + 4 * i) =
                            make_float4(__uint_as_float(t16[4 * i]), __uint_as_float(t16[4 * i + 1]), __uint_as_float(t16[4 * i + 2]), __uint_as_float(t16[4 * i + 3]));
                }
                __syncwarp();
                const float4 bi = ldg4(a.bias + cb * 32 + c4);
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int r = p * 4 + rsub, row = prow0 + warp * 32 + r;
                    if (row < a.M) {
                        float4 y = *reinterpret_cast<const float4*>(stg + r * TC_STG_LD + c4);
                        y = add4(y, bi);
                        *reinterpret_cast<float4*>(a.out + (long long)row * 64 + cb * 32 + c4) = y;
                        const float s = (y.x + y.y) + (y.z + y.w), q = (y.x * y.x + y.y * y.y) + (y.z * y.z + y.w * y.w);
                        if (row < split) {
                            s0 += s;
                            q0 += q;
                        } else {
                            s1 += s;
                            q1 += q;
                        }
                    }
                }
                __syncwarp();
            }
            group_stats_atomic(s0, q0, a.sums + 2 * bfirst, scratch, tid, 128, 1);
            group_stats_atomic(s1, q1, (bfirst + 1 < a.B) ? a.sums + 2 * (bfirst + 1) : nullptr, scratch, tid, 128, 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem);
}

inline cudaError_t launch_gateproj_wide(GwArgs a, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gateproj_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GW_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    a.ntiles = (a.M + 127) / 128;
    const int grid = a.ntiles < 148 ? a.ntiles : 148;
    gateproj_wide_kernel<<<grid, GW_NT, GW_SMEM, st>>>(a);
    return cudaGetLastError();
}

}  // namespace rtfs
