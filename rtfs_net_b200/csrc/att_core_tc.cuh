// tcgen05 + tensor-map TMA attention core of the TF self-attention (reference: MultiHeadSelfAttention2D.forward,
// layers/attention.py:173-183: softmax(Q K^T / sqrt(E*F)) V per (batch, head), tokens = time frames).
//
// One CTA per (128-query tile, batch*head).  Everything that moves is moved by the TMA engine from 3-D tensor maps
// (feature, token, batch*head) with SWIZZLE_128B boxes, so the shared-memory tiles ARE the canonical UMMA operand layouts
// and no thread touches Q, K or V:
//   * S = Q K^T      : K-major A (queries) and B (keys) boxes of 32 features x 128/256 tokens, 8 chunks x 4 UMMAs (K = 8),
//                      accumulator S (128 x KP fp32) in TMEM;
//   * softmax        : thread r owns TMEM lane r = query row r: max and exp straight out of TMEM (no shuffles), the
//                      un-normalised probabilities go back to shared memory tf32-rounded as the K-major SW128 A operand of the
//                      second contraction, 1/sum stays in a register for the epilogue;
//   * O = P V        : V is consumed as an MN-major B operand (token rows of 32 features = one 128-byte swizzle line), so the
//                      (token, feature) rows written by the Q/K/V kernel need no transpose: 4 chunks of 256 value columns,
//                      accumulators double-buffered in TMEM (the first aliases S) so the epilogue of one chunk overlaps the
//                      UMMAs of the next;
//   * epilogue       : tcgen05.ld -> x 1/sum -> per-warp shared-memory transpose -> 64-byte row pieces of the (B,Tc,64,64)
//                      channels-last tensor (channel = head*16 + j).
// Out-of-range tokens of the last tile are zero-filled by the TMA unit (tensor-map bounds); masked to -inf before the softmax.
// Warp roles: 0-3 softmax + epilogue (TMEM lane quarters), 4 TMA producer, 5 MMA issuer.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_tc.cuh"
#include "dprnn_fused.cuh"  // mbar_arrive

namespace rtfs {

// ---------------------------------------------------------------------------------- tensor-map TMA helpers
typedef CUresult (*PFN_encodeTiled_rtfs)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_encodeTiled_rtfs tmap_encoder() {
    static PFN_encodeTiled_rtfs fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (PFN_encodeTiled_rtfs)p;
    }();
    return fn;
}
// fp32 tensor (inner, rows, batch) with rows of `inner` contiguous floats; box = (box_inner, box_rows, 1), 128-byte swizzle
inline bool tmap_rows3d(CUtensorMap* m, const float* base, int inner, int rows, int batch, int box_inner, int box_rows,
                        CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    PFN_encodeTiled_rtfs enc = tmap_encoder();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)inner * 4, (cuuint64_t)inner * 4 * (cuuint64_t)rows};
    cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// TMA tensor load (SASS: UTMALDG), completion on an mbarrier
DEVINL void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
DEVINL void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// UMMA shared-memory descriptors for SWIZZLE_128B tiles (cute::UMMA::SmemDescriptor, version 1, layout type 2):
//   K-major : rows of 128 bytes (32 tf32 along K), 8-row groups SBO = 1024 bytes apart; a K step of 8 advances the start by 32 B
DEVINL uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major TF32 operands exist in ONE shared-memory layout (layout type 1, SWIZZLE_128B_BASE32B; the plain 128-byte swizzle
// reads as zeros -- tools/probe/mnmajor_probe.cu): lines of 128 bytes (32 tf32 along N) per K index whose 32-byte chunks are
// XORed with (line & 3), 4-line groups SBO apart, 32-wide N atoms LBO apart:
//   byte(k, n) = (n/32) LBO + (k/4) SBO + (k%4) 128 + (((n%32)/8) ^ (k%4)) 32 + (n%8) 4
// which is what the TMA unit writes for a (32 floats x rows) box under CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (SBO = 512).
DEVINL uint64_t umma_desc_mn32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
// instruction descriptor: fp32 accumulate, TF32 x TF32, A K-major, B MN-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32_bmn(int M, int N) { return umma_idesc_tf32(M, N) | (1u << 16); }

// ---------------------------------------------------------------------------------- kernel
struct AttnTcArgs {
    float* o;  // (B, Tc, 64, 64), channel = h*16 + j
    int Tc, H;
    float scale_log2e;  // log2(e) / sqrt(E*F)
    float* dbg_s;       // probe only (tools/probe/attcore_probe.cu): raw scores [bh][query][KP], else null
};

constexpr int ATC_KVK = 32;  // keys per V stage
template <int NK>
struct AtcCfg {
    static constexpr int KP = 128 * NK;                       // padded keys
    static constexpr int P_BYTES = 128 * KP * 4;              // probabilities, (KP/32) blocks of [128 rows][128 B]
    static constexpr int NSV = NK == 1 ? 3 : 2;               // V stages
    static constexpr int V_STAGE = ATC_KVK * 256 * 4;         // 8 atoms x [32 keys][128 B]
    static constexpr int NSQ = 3;                             // Q/K stages (alias P + the V ring)
    static constexpr int QK_STAGE = (128 + KP) * 128;         // [128 queries][128 B] + [KP keys][128 B]
    static constexpr int STG_BYTES = 4 * 32 * TC_STG_LD * 4;  // epilogue transpose, per warp [32][36]
    static constexpr int MAIN = P_BYTES + NSV * V_STAGE > NSQ * QK_STAGE ? P_BYTES + NSV * V_STAGE : NSQ * QK_STAGE;
    static constexpr int SMEM = 1024 + MAIN + STG_BYTES + 256;
};

template <int NK>
__global__ void __launch_bounds__(192, 1) attn_core_tc_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                                                              const __grid_constant__ CUtensorMap mv, AttnTcArgs a) {
    using C = AtcCfg<NK>;
    constexpr int KP = C::KP, NSV = C::NSV, NSQ = C::NSQ;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* p_sm = base;                  // P
    unsigned char* v_sm = base + C::P_BYTES;     // V ring
    unsigned char* qk_sm = base;                 // Q/K ring (first phase only)
    float* stg_all = reinterpret_cast<float*>(base + C::MAIN);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + C::MAIN + C::STG_BYTES);
    uint64_t* fullq = bars;            // [NSQ]
    uint64_t* emptyq = bars + 3;       // [NSQ]
    uint64_t* fullv = bars + 6;        // [NSV]
    uint64_t* emptyv = bars + 9;       // [NSV]
    uint64_t* s_ready = bars + 12;
    uint64_t* p_ready = bars + 13;     // 128 arrivals
    uint64_t* o_ready = bars + 14;     // [2]
    uint64_t* o_free = bars + 16;      // [2], 128 arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * 128, bh = blockIdx.y;
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    if (tid == 32) {
        for (int i = 0; i < 3; ++i) {
            mbar_init(fullq + i, 1);
            mbar_init(emptyq + i, 1);
            mbar_init(fullv + i, 1);
            mbar_init(emptyv + i, 1);
        }
        mbar_init(s_ready, 1);
        mbar_init(p_ready, 128);
        mbar_init(o_ready, 1);
        mbar_init(o_ready + 1, 1);
        mbar_init(o_free, 128);
        mbar_init(o_free + 1, 128);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one()) {
            tma_prefetch_desc(&mq);
            tma_prefetch_desc(&mk);
            tma_prefetch_desc(&mv);
            for (int c = 0; c < 8; ++c) {
                const int s = c % NSQ, u = c / NSQ;
                mbar_wait(emptyq + s, (u & 1) ^ 1);
                mbar_expect_tx(fullq + s, C::QK_STAGE);
                unsigned char* st = qk_sm + s * C::QK_STAGE;
                tma_load_3d(st, &mq, c * 32, q0, bh, fullq + s);
#pragma unroll
                for (int kb = 0; kb < NK; ++kb) tma_load_3d(st + 128 * 128 + kb * 128 * 128, &mk, c * 32, kb * 128, bh, fullq + s);
            }
            // the V ring aliases the Q/K ring: wait until every Q K^T MMA has read its operands
            mbar_wait(s_ready, 0);
            const int nkb = (a.Tc + 31) >> 5;  // 32-key blocks that hold at least one token
            int i = 0;
            for (int nc = 0; nc < 4; ++nc)
                for (int kb = 0; kb < nkb; ++kb, ++i) {
                    const int s = i % NSV, u = i / NSV;
                    mbar_wait(emptyv + s, (u & 1) ^ 1);
                    mbar_expect_tx(fullv + s, C::V_STAGE);
                    unsigned char* st = v_sm + s * C::V_STAGE;
#pragma unroll
                    for (int at = 0; at < 8; ++at) tma_load_3d(st + at * (ATC_KVK * 128), &mv, nc * 256 + at * 32, kb * ATC_KVK, bh, fullv + s);
                }
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------ MMA issuer (uniform-register issue: gemm_tc.cuh elect_one)
        if (elect_one()) {
            constexpr uint32_t IDESC_S = umma_idesc_tf32(128, KP);
            constexpr uint32_t IDESC_O = umma_idesc_tf32_bmn(128, 256);
            for (int c = 0; c < 8; ++c) {
                const int s = c % NSQ, u = c / NSQ;
                mbar_wait(fullq + s, u & 1);
                tc_fence_after();
                const uint32_t qa = smem_u32(qk_sm + s * C::QK_STAGE), ka = qa + 128 * 128;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_tf32(tmem, umma_desc_sw128(qa + ks * 32, 16, 1024), umma_desc_sw128(ka + ks * 32, 16, 1024), IDESC_S, (c | ks) ? 1u : 0u);
                umma_commit(emptyq + s);
            }
            umma_commit(s_ready);
            mbar_wait(p_ready, 0);  // P is in shared memory, S (aliased by the first O buffer) has been read
            tc_fence_after();
            const int nkb = (a.Tc + 31) >> 5;
            int i = 0;
            for (int nc = 0; nc < 4; ++nc) {
                const int buf = nc & 1;
                if (nc >= 2) {
                    mbar_wait(o_free + buf, ((nc >> 1) - 1) & 1);
                    tc_fence_after();
                }
                for (int kb = 0; kb < nkb; ++kb, ++i) {
                    const int s = i % NSV, u = i / NSV;
                    mbar_wait(fullv + s, u & 1);
                    tc_fence_after();
                    const uint32_t pa = smem_u32(p_sm + kb * (128 * 128)), va = smem_u32(v_sm + s * C::V_STAGE);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_tf32(tmem + buf * 256, umma_desc_sw128(pa + kk * 32, 16, 1024), umma_desc_mn32(va + kk * 1024, ATC_KVK * 128, 512),
                                  IDESC_O, (kb | kk) ? 1u : 0u);
                    umma_commit(emptyv + s);
                }
                umma_commit(o_ready + buf);
            }
        }
    } else {
        // ------------------------------------------------------------ softmax + epilogue: thread = query row
        const int r = tid;  // TMEM lane
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        const int Tc = a.Tc, nkb = (Tc + 31) >> 5;
        mbar_wait(s_ready, 0);
        tc_fence_after();
        float m = -INFINITY;
#pragma unroll 1
        for (int cb = 0; cb < nkb; ++cb) {
            uint32_t v[32];
            tmem_ld32(tl + cb * 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (cb * 32 + i < Tc) m = fmaxf(m, __uint_as_float(v[i]));
            if (a.dbg_s != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; ++i) a.dbg_s[((long long)bh * gridDim.x * 128 + q0 + r) * KP + cb * 32 + i] = __uint_as_float(v[i]);
            }
        }
        const float ms = m * a.scale_log2e;
        float sum = 0.f;
#pragma unroll 1
        for (int cb = 0; cb < nkb; ++cb) {
            uint32_t v[32];
            tmem_ld32(tl + cb * 32, v);
            unsigned char* prow = p_sm + cb * (128 * 128) + r * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float e[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int key = cb * 32 + q * 4 + i;
                    const float p = key < Tc ? tf32r(exp2f(fmaf(__uint_as_float(v[q * 4 + i]), a.scale_log2e, -ms))) : 0.f;
                    e[i] = p;
                    sum += p;
                }
                *reinterpret_cast<float4*>(prow + ((q ^ (r & 7)) << 4)) = make_float4(e[0], e[1], e[2], e[3]);
            }
        }
        const float inv = 1.f / sum;
        fence_proxy_async();  // the generic-proxy writes of P must be visible to the tensor core's async-proxy reads
        tc_fence_before();
        mbar_arrive(p_ready);

        const int b = bh / a.H, h = bh - b * a.H;
        float* stg = stg_all + warp * (32 * TC_STG_LD);
        const int rsub = lane >> 3, piece = (lane >> 2) & 1, quad = lane & 3;
        for (int nc = 0; nc < 4; ++nc) {
            const int buf = nc & 1;
            mbar_wait(o_ready + buf, (nc >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cb = 0; cb < 8; ++cb) {
                uint32_t v[32];
                tmem_ld32(tl + buf * 256 + cb * 32, v);
#ifdef ATC_PROBE_RAW
                if (a.dbg_s != nullptr && nc == 0 && cb == 0) {
                    float* d = a.dbg_s + ((long long)bh * gridDim.x * 128 + q0 + r) * KP;
                    d[0] = sum;
                    for (int i = 0; i < 32; ++i) d[1 + i] = __uint_as_float(v[i]);
                }
#endif
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(stg + lane * TC_STG_LD + q * 4) =
                        make_float4(__uint_as_float(v[q * 4]) * inv, __uint_as_float(v[q * 4 + 1]) * inv, __uint_as_float(v[q * 4 + 2]) * inv,
                                    __uint_as_float(v[q * 4 + 3]) * inv);
                __syncwarp();
                const int f = nc * 16 + cb * 2 + piece;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = it * 4 + rsub, rg = q0 + warp * 32 + row;
                    if (rg < Tc) {
                        const float4 y = *reinterpret_cast<const float4*>(stg + row * TC_STG_LD + piece * 16 + quad * 4);
                        *reinterpret_cast<float4*>(a.o + (((long long)b * Tc + rg) * 64 + f) * 64 + h * 16 + quad * 4) = y;
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(o_free + buf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// q, k: (B*H, Tc, 256) ; v: (B*H, Tc, 1024) ; all tf32-rounded by the producing kernel.  Tc <= 256 (the caller runs the
// mma.sync kernel on longer sequences); a missing tensor-map driver entry point is an error, not a fallback.
inline cudaError_t launch_attn_core_tc(const float* q, const float* k, const float* v, float* o, int B, int H, int Tc, cudaStream_t st,
                                       float* dbg_s = nullptr) {
    if (Tc > 256) return cudaErrorNotSupported;
    const int NK = Tc > 128 ? 2 : 1, BH = B * H;
    CUtensorMap mq, mk, mv;
    if (!tmap_rows3d(&mq, q, 256, Tc, BH, 32, 128) || !tmap_rows3d(&mk, k, 256, Tc, BH, 32, 128) || !tmap_rows3d(&mv, v, 1024, Tc, BH, 32, ATC_KVK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
        return cudaErrorUnknown;
    AttnTcArgs a{o, Tc, H, 1.4426950408889634f / sqrtf(4.f * 64.f), dbg_s};
    const dim3 grid((Tc + 127) / 128, BH);
    if (NK == 1) {
        static SmemCfg cfg;
        if (cudaError_t e = ensure_smem(attn_core_tc_kernel<1>, AtcCfg<1>::SMEM, cfg); e != cudaSuccess) return e;
        attn_core_tc_kernel<1><<<grid, 192, AtcCfg<1>::SMEM, st>>>(mq, mk, mv, a);
    } else {
        static SmemCfg cfg;
        if (cudaError_t e = ensure_smem(attn_core_tc_kernel<2>, AtcCfg<2>::SMEM, cfg); e != cudaSuccess) return e;
        attn_core_tc_kernel<2><<<grid, 192, AtcCfg<2>::SMEM, st>>>(mq, mk, mv, a);
    }
    return cudaGetLastError();
}

}  // namespace rtfs
