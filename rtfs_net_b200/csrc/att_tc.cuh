// tcgen05 version of the two conv + PReLU + LayerNormalization4D stages of the TF self-attention
// (reference: MultiHeadSelfAttention2D.forward, layers/attention.py:149-189; ConvActNorm = 1x1 conv -> PReLU ->
// LayerNormalization4D over (E, F), layers/conv_layers.py:201-205, layers/normalizations.py:20-37):
//   MODE 0 : the 12 Q/K/V head convs of TWO frames as one 128 x 96 x 64 UMMA (rows = (frame, f), columns = [Q 4x4 | K 4x4 | V 4x16]),
//            PReLU per conv, LN over (E, F) per (frame, conv), per-head token rows q/k (B,H,Tc,256), v (B,H,Tc,1024), tf32-rounded
//   MODE 1 : concat-projection 64 -> 64 of two frames (128 x 64 x 64), PReLU, LN over (C, F), + residual
// One persistent CTA of 128 threads per SM slot (2 per SM): thread r owns TMEM lane r = row (frame r>>6, f = r&63), so the
// LN statistics of a (frame, conv) are a reduction over the 64 lanes of two warps (shuffles + one shared-memory exchange,
// two passes: mean, then centred sum of squares), and every thread writes E contiguous floats of its token row
// (a warp writes 512 B .. 2 KB contiguous).  The 24 / 16 KB weight image stays resident in shared memory.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace rtfs {

constexpr int AC_LBO = 128 * 16 + 16;    // activation slab: 16 K-pieces x (128 rows x 16 B + pad)
constexpr int AC_A_BYTES = 16 * AC_LBO;  // 33024
// CTAs per SM: the per-CTA chain (load -> MMA -> three TMEM passes) is latency-bound, so as many as registers (MODE 0:
// 168) and shared memory (MODE 1: the transpose staging aliases the activation slab) allow
template <int MODE>
__host__ __device__ constexpr int ac_ctas() { return MODE == 0 ? 3 : 4; }
template <int MODE>
__host__ __device__ constexpr int ac_a_area() { return MODE == 1 ? 4 * 32 * 68 * 4 : AC_A_BYTES; }  // 34816 / 33024

struct AttConvArgs {
    const float* x;      // (B*Tc, 64 f, 64 c)
    const float* resid;  // MODE 1
    const float* wimg;   // [16 pieces][N][4] tf32
    const float* bias;   // [N]
    const float* slope;  // [groups]
    const float* gamma;  // per group, [f*E+e] order, groups concatenated in column order
    const float* beta;
    float* q;
    float* k;
    float* v;
    float* out;  // MODE 1 (B*Tc, 64, 64)
    int B, Tc, H, nframes;
};

template <int N, int MODE>
constexpr int att_conv_smem() {
    return ac_a_area<MODE>() + N * 64 * 4 + 2 * 4 * 16 * 4 + 64;
}

// group g of MODE 0: columns [c0, c0 + E)
DEVINL void ac_group(int g, int& c0, int& E) {
    if (g < 4) {
        c0 = g * 4;
        E = 4;
    } else if (g < 8) {
        c0 = 16 + (g - 4) * 4;
        E = 4;
    } else {
        c0 = 32 + (g - 8) * 16;
        E = 16;
    }
}

template <int N, int MODE>
__global__ void __launch_bounds__(128, ac_ctas<MODE>()) att_conv_ln_tc_kernel(AttConvArgs a) {
    constexpr int NG = MODE == 0 ? 12 : 1;
    constexpr int TCOLS = N <= 64 ? 64 : 128;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* a_slab = smem_raw;
    unsigned char* w_slab = smem_raw + ac_a_area<MODE>();
    float* part = reinterpret_cast<float*>(w_slab + N * 64 * 4);  // [pass 2][warp 4][16]
    float* stg_all = reinterpret_cast<float*>(smem_raw);  // MODE 1: 4 x [32][68], aliases the slab (free once the MMAs have completed)
    uint64_t* bars = reinterpret_cast<uint64_t*>(part + 2 * 4 * 16);
    uint64_t* w_ready = bars;
    uint64_t* mma_done = bars + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) tmem_alloc<TCOLS>(tmem_slot);
    if (tid == 32) {
        mbar_init(w_ready, 1);
        mbar_init(mma_done, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        mbar_expect_tx(w_ready, N * 64 * 4);
        bulk_g2s(w_slab, a.wimg, N * 64 * 4, w_ready);
    }
    constexpr uint32_t IDESC = umma_idesc_tf32(128, N);
    const int kq = tid & 15, r0 = tid >> 4;  // A staging: piece kq, rows r0 + 8 i
    const int fr = warp >> 1, f = tid & 63;  // epilogue: frame of the pair, frequency bin
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    const int npairs = (a.nframes + 1) >> 1;

    int it = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
        // ---- stage the two frames (128 rows x 64 channels), tf32-rounded, in the UMMA K-major slab
        const long long row_base = (long long)pair * 128;
        const long long nrows = (long long)a.nframes * 64;
        float4 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const long long row = row_base + r0 + 8 * i;
            v[i] = row < nrows ? ldg4(a.x + row * 64 + kq * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float4 y = make_float4(tf32r(v[i].x), tf32r(v[i].y), tf32r(v[i].z), tf32r(v[i].w));
            *reinterpret_cast<float4*>(a_slab + kq * AC_LBO + (r0 + 8 * i) * 16) = y;
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (warp == 0 && elect_one()) {  // issued straight from uniform registers (gemm_tc.cuh: elect_one)
            if (it == 0) mbar_wait(w_ready, 0);
            const uint32_t ab = smem_u32(a_slab), wb = smem_u32(w_slab);
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8)
                umma_tf32(tmem, umma_desc(ab + 2 * k8 * AC_LBO, AC_LBO, 128), umma_desc(wb + 2 * k8 * (N * 16), N * 16, 128), IDESC, k8 > 0 ? 1u : 0u);
            umma_commit(mma_done);
        }
        mbar_wait(mma_done, it & 1);
        tc_fence_after();

        const int frame = pair * 2 + fr;
        const bool fvalid = frame < a.nframes;
        // ---- pass 1: per (frame, group) mean of PReLU(conv + bias)
        float mean[NG], rstd[NG];
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
            for (int cb = 0; cb < N / 16; ++cb) {
                uint32_t t16[16];
                tmem_ld16(tl + cb * 16, t16);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (MODE == 0) {
                    // 16 columns = 4 groups of 4 (Q, K) or one group of 16 (V)
                    if (cb < 2) {
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg) {
                            const int g = cb * 4 + gg;
                            const float sl = __ldg(a.slope + g);
                            float s = 0.f;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float y = prelu(__uint_as_float(t16[gg * 4 + e]) + __ldg(a.bias + cb * 16 + gg * 4 + e), sl);
                                s += pass == 0 ? y : (y - mean[g]) * (y - mean[g]);
                            }
                            s = warp_sum(s);
                            if (lane == 0) part[(pass * 4 + warp) * 16 + g] = s;
                        }
                    } else {
                        const int g = 8 + (cb - 2);
                        const float sl = __ldg(a.slope + g);
                        float s = 0.f;
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const float y = prelu(__uint_as_float(t16[e]) + __ldg(a.bias + cb * 16 + e), sl);
                            s += pass == 0 ? y : (y - mean[g]) * (y - mean[g]);
                        }
                        s = warp_sum(s);
                        if (lane == 0) part[(pass * 4 + warp) * 16 + g] = s;
                    }
                } else {
                    const float sl = __ldg(a.slope);
                    float s = 0.f;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float y = prelu(__uint_as_float(t16[e]) + __ldg(a.bias + cb * 16 + e), sl);
                        s += pass == 0 ? y : (y - mean[0]) * (y - mean[0]);
                    }
                    s = warp_sum(s);
                    if (lane == 0) {
                        if (cb == 0) part[(pass * 4 + warp) * 16] = s;
                        else part[(pass * 4 + warp) * 16] += s;
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                int c0 = 0, E = 64;
                if (MODE == 0) ac_group(g, c0, E);
                const float tot = part[(pass * 4 + 2 * fr) * 16 + g] + part[(pass * 4 + 2 * fr + 1) * 16 + g];
                if (pass == 0) mean[g] = tot / (float)(64 * E);
                else rstd[g] = 1.f / sqrtf(tot / (float)(64 * E) + RTFS_EPS);
            }
        }
        // ---- pass 3: normalise and write
        if (MODE == 0) {
            const int b = frame / a.Tc, tt = frame - b * a.Tc;
            int goff = 0;
#pragma unroll
            for (int cb = 0; cb < N / 16; ++cb) {
                uint32_t t16[16];
                tmem_ld16(tl + cb * 16, t16);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cb < 2) {
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) {
                        const int g = cb * 4 + gg, h = gg;
                        const float sl = __ldg(a.slope + g);
                        float* dst = (cb == 0 ? a.q : a.k) + (((long long)b * a.H + h) * a.Tc + tt) * 256 + f * 4;
                        const float4 gm = ldg4(a.gamma + goff + f * 4), be = ldg4(a.beta + goff + f * 4);
                        float y[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) y[e] = (prelu(__uint_as_float(t16[gg * 4 + e]) + __ldg(a.bias + cb * 16 + gg * 4 + e), sl) - mean[g]) * rstd[g];
                        if (fvalid)
                            *reinterpret_cast<float4*>(dst) = make_float4(tf32r(y[0] * gm.x + be.x), tf32r(y[1] * gm.y + be.y), tf32r(y[2] * gm.z + be.z), tf32r(y[3] * gm.w + be.w));
                        goff += 256;
                    }
                } else {
                    const int g = 8 + (cb - 2), h = cb - 2;
                    const float sl = __ldg(a.slope + g);
                    float* dst = a.v + (((long long)b * a.H + h) * a.Tc + tt) * 1024 + f * 16;
#pragma unroll
                    for (int e4 = 0; e4 < 4; ++e4) {
                        const float4 gm = ldg4(a.gamma + goff + f * 16 + e4 * 4), be = ldg4(a.beta + goff + f * 16 + e4 * 4);
                        float y[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            y[e] = (prelu(__uint_as_float(t16[e4 * 4 + e]) + __ldg(a.bias + cb * 16 + e4 * 4 + e), sl) - mean[g]) * rstd[g];
                        if (fvalid)
                            *reinterpret_cast<float4*>(dst + e4 * 4) = make_float4(tf32r(y[0] * gm.x + be.x), tf32r(y[1] * gm.y + be.y), tf32r(y[2] * gm.z + be.z), tf32r(y[3] * gm.w + be.w));
                    }
                    goff += 1024;
                }
            }
        } else {
            // transpose the PReLU'd rows through shared memory so that gamma / beta / residual loads and the output store
            // are row-contiguous (lane = 4-column piece, 2 rows per instruction)
            const float sl = __ldg(a.slope);
            float* stg = stg_all + warp * (32 * 68);
#pragma unroll
            for (int cb = 0; cb < N / 16; ++cb) {
                uint32_t t16[16];
                tmem_ld16(tl + cb * 16, t16);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                    const int c = cb * 16 + e4 * 4;
                    const float4 bi = ldg4(a.bias + c);
                    *reinterpret_cast<float4*>(stg + lane * 68 + c) =
                        make_float4(prelu(__uint_as_float(t16[e4 * 4]) + bi.x, sl), prelu(__uint_as_float(t16[e4 * 4 + 1]) + bi.y, sl),
                                    prelu(__uint_as_float(t16[e4 * 4 + 2]) + bi.z, sl), prelu(__uint_as_float(t16[e4 * 4 + 3]) + bi.w, sl));
                }
            }
            __syncwarp();
            const int rsub = lane >> 4, c = (lane & 15) * 4;
            const float mu = mean[0], rs = rstd[0];
#pragma unroll 4
            for (int p = 0; p < 16; ++p) {
                const int r = p * 2 + rsub;                 // row of this warp's 32
                const int ff = (warp & 1) * 32 + r;         // frequency bin
                const long long o = ((long long)frame * 64 + ff) * 64 + c;
                if (fvalid) {
                    const float4 y = *reinterpret_cast<const float4*>(stg + r * 68 + c);
                    const float4 gm = ldg4(a.gamma + ff * 64 + c), be = ldg4(a.beta + ff * 64 + c), rr = ldg4(a.resid + o);
                    *reinterpret_cast<float4*>(a.out + o) = make_float4((y.x - mu) * rs * gm.x + be.x + rr.x, (y.y - mu) * rs * gm.y + be.y + rr.y,
                                                                        (y.z - mu) * rs * gm.z + be.z + rr.z, (y.w - mu) * rs * gm.w + be.w + rr.w);
                }
            }
            __syncwarp();
        }
        tc_fence_before();
        __syncthreads();  // TMEM accumulator, slab and `part` are free for the next pair
        tc_fence_after();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TCOLS>(tmem);
}

template <int N, int MODE>
inline cudaError_t launch_att_conv_tc(const AttConvArgs& a, cudaStream_t st) {
    auto kern = att_conv_ln_tc_kernel<N, MODE>;
    const int smem = att_conv_smem<N, MODE>();
    static SmemCfg cfg;  // per instantiation, per device
    if (cudaError_t e = ensure_smem(kern, smem, cfg); e != cudaSuccess) return e;
    const int npairs = (a.nframes + 1) / 2;
    const int grid = npairs < sm_count() * ac_ctas<MODE>() ? npairs : sm_count() * ac_ctas<MODE>();
    kern<<<grid, 128, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace rtfs
