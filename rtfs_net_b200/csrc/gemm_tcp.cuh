// Persistent, warp-specialised tcgen05 GEMM for the full-resolution 1x1 convolutions of the RTFS-Net forward
// (audio bottleneck, gateway+projection, residual conv, S^3 mask): the HBM-bound contractions whose A operand
// and epilogue each stream a (B*T*F) x 256 fp32 tensor.  Same math, operand layouts, loader and epilogue functors
// as gemm_tc.cuh; what changes is the schedule: one CTA per SM loops over 128-row tiles with three groups of
// warps running concurrently, so the loads of tile i+1, the MMAs of tile i and the epilogue traffic of tile i-1
// are all in flight at once:
//
//   warps 0-7   epilogue : tmem_full[acc] -> tcgen05.ld -> shared-memory transpose -> functor (batched global loads,
//                          coalesced float4 stores) -> tmem_empty[acc]          (two TMEM accumulators, ping-pong)
//   warps 8-15  producers: loader functor (fused prologue) -> TF32 -> UMMA K-major slab in an NSA-stage ring ->
//                          full_a[stage]; a stage is recycled when the MMAs that read it commit to empty_a[stage]
//   warp 16     MMA      : one thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) x 4 per K chunk
//   warp 17     weights  : WRES: the whole image (<= 64 KB) is bulk-copied once and stays resident;
//                          otherwise 32-wide K slabs stream through an NSW-stage ring (cp.async.bulk + mbarrier)
#pragma once
#include <cstdio>
#include <type_traits>
#include "common.cuh"
#include "gemm_tc.cuh"

namespace rtfs {

template <class AL, class = void>
struct loader_tile_invariant {
    static constexpr bool value = false;
};
template <class AL>
struct loader_tile_invariant<AL, decltype((void)AL::kTileInvariant)> {
    static constexpr bool value = AL::kTileInvariant;
};

#ifndef TCP_REGSPLIT
#define TCP_REGSPLIT 0   // setmaxnreg split (works; measured 0.74 vs 0.71 ms for the residual conv: the 72-register producers lose more than the epilogue gains)
#endif
#ifndef TCP_AFFINE
#define TCP_AFFINE 0   // block-pointer epilogue addressing (gemm_tc.cuh): its 64 live load registers spill under the same cap
#endif
#ifndef TCP_BATCHED
#define TCP_BATCHED 0  // batched raw loads in the per-tile producer: 64 live registers, spills under the 96-register cap of 18 warps
#endif
constexpr int TCP_EPI = 256;  // epilogue threads (warps 0-7); then NPROD producer threads, the MMA warp and the weight warp

DEVINL void mbar_arrive_cta(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BN, int KTOT, int NSA, int NSW, bool WRES>
__host__ __device__ constexpr int tcp_smem_bytes(int extra_floats) {
    return NSA * TC_A_STAGE + (WRES ? (KTOT / TC_KC) : NSW) * BN * 128 + TC_STG_BYTES + 2 * extra_floats * 4 + 512;  // + ep_smem_bytes<EP> (launch)
}

// EPS: epilogue functor exposes finish_group(scratch, gtid, nthr, barid) (gLN statistics) instead of finish()
// Producer modes (template parameter ASYNC):
//   0  per-tile register pipeline through the loader functor's load() (any loader; PF chunks in flight, drains at tile ends)
//   1  raw chunk landed in its ring stage with 16-byte cp.async, transformed in place -- measured SLOWER: LDGSTS streams at
//      ~3.4 TB/s on this part whatever the depth (tools/probe/cpasync_probe.cu), plain loads at 5.7-5.9 TB/s
//   2  flat raw register stream: PF chunks of RAW values in flight per thread, running ahead across tile boundaries;
//      the loader's xform() is applied when a chunk is stored (loaders with raw()/xform(): gLN-act, gateway, PReLU)
#ifdef RTFS_TCP_TRACE  // build with RTFS_NVCC_EXTRA=-DRTFS_TCP_TRACE: clock64 stamps of epilogue warp 0 of CTA 0 in its third tile
__device__ long long g_tcp_trace[64];
#define TCP_STAMP(i)                                                                            \
    do {                                                                                        \
        if (blockIdx.x == 0 && tid == 0 && it == 2) g_tcp_trace[i] = clock64();                 \
    } while (0)
#else
#define TCP_STAMP(i)
#endif
template <int BN, int KTOT, int NSA, int NSW, bool WRES, int PF, int ASYNC, int NPROD, class AL, class EP>
__global__ void __launch_bounds__(TCP_EPI + NPROD + 64, 1) gemm_tcp_kernel(AL al, const float* __restrict__ Wimg, EP ep, int M, int ntiles) {
    constexpr int TCP_PROD = NPROD;             // 256 or 512 producer threads
    constexpr int RPT = 1024 / NPROD;           // A rows per producer thread per chunk
    constexpr int RS = NPROD / 8;               // row stride between them
    constexpr int MMA_WARP = (TCP_EPI + NPROD) / 32;
    static_assert(NPROD == 256 || NPROD == 512, "8 or 16 producer warps");
    constexpr int NK = KTOT / TC_KC;
    constexpr int WBYTES = BN * 128;
    constexpr int NWS = WRES ? NK : NSW;  // weight slabs held in shared memory
    static_assert(2 * BN <= 512, "two accumulators must fit the 512 TMEM columns");
    static_assert(PF >= 1 && PF <= NK && PF <= 4, "register prefetch depth");
    static_assert(BN % 64 == 0, "N must be a multiple of 64");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* a_stage = smem_raw;
    unsigned char* w_stage = a_stage + NSA * TC_A_STAGE;
    float* stg_all = reinterpret_cast<float*>(w_stage + NWS * WBYTES);
    float* extra = stg_all + TC_STG_BYTES / 4;  // two loader tables (ping-pong by tile)
    float* ep_tab = extra + 2 * AL::kExtra;  // epilogue-owned table (fused mask + decoder: the decoder filter)
    uint64_t* bars = reinterpret_cast<uint64_t*>(ep_tab + ep_smem_bytes<EP>::value / 4);
    uint64_t* full_a = bars;                 // [NSA] count TCP_PROD
    uint64_t* empty_a = full_a + NSA;        // [NSA] count 1 (tcgen05.commit)
    uint64_t* full_w = empty_a + NSA;        // [NWS] count 1 + tx
    uint64_t* empty_w = full_w + NWS;        // [NWS] count 1 (streamed only)
    uint64_t* tmem_full = empty_w + NWS;     // [2] count 1
    uint64_t* tmem_empty = tmem_full + 2;    // [2] count TCP_EPI
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* scratch = reinterpret_cast<float*>(tmem_slot + 4);  // 16 floats

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) tmem_alloc<(2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512)>(tmem_slot);
    if (tid == 32) {
        for (int s = 0; s < NSA; ++s) {
            mbar_init(full_a + s, TCP_PROD / 32);  // one elected arrive per producer warp
            mbar_init(empty_a + s, 1);
        }
        for (int s = 0; s < NWS; ++s) {
            mbar_init(full_w + s, 1);
            mbar_init(empty_w + s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full + s, 1);
            mbar_init(tmem_empty + s, TCP_EPI / 32);
        }
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t IDESC = umma_idesc_tf32(TC_BM, BN);

    // Register split (8-warp producer configuration only): 18 warps cap every thread at 96 registers, which spills the
    // residual epilogue's 64 in-flight load registers; the producer warpgroups hand 24 registers per thread to the
    // epilogue warpgroups (setmaxnreg works on aligned groups of 4 warps; the MMA / weight warps keep their 96).
    constexpr bool REGSPLIT = TCP_REGSPLIT && NPROD == 256;
    // 16-producer-warp configuration (26 warps, 72 registers each): epilogues that hold global inputs in registers (S^3 mask)
    // ask for 104 through EP::kTcpEpiRegs, the producers drop to 56
    constexpr bool REGSPLIT16 = NPROD == 512 && ep_epi_regs<EP>::value > 0;
    if (warp < 8) {
        if constexpr (REGSPLIT) asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
        if constexpr (REGSPLIT16) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ep_epi_regs<EP>::value));
        // ================================================================= epilogue
        const int q = warp & 3, hlf = warp >> 2;
        float* stg = stg_all + warp * (32 * TC_STG_LD);
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;
        int it = 0;
        constexpr bool AFF = ep_affine<EP>::value && TCP_AFFINE && REGSPLIT;
        // ROLL: the epilogue's global inputs of block n+1 (next 32 columns, or the first block of this CTA's next tile) are
        // requested right after the matching rows of block n are stored -- same registers, but a warp no longer exposes one
        // memory round trip per block (4 per tile for the mask epilogue, which made the epilogue warps the critical path)
        constexpr bool ROLL = ep_roll<EP>::value && !AFF;
        constexpr bool FUSED = ep_fused_rows<EP>::value;
        if constexpr (FUSED) {
            ep.bind_smem(ep_tab, tid, TCP_EPI);
            named_bar_sync(3, TCP_EPI);
        }
        typename std::conditional<AFF, typename ep_pre<EP>::type, typename EP::Pre>::type pre[8];
        if constexpr (ROLL) {
            if ((int)blockIdx.x < ntiles) {
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int row = blockIdx.x * TC_BM + q * 32 + rsub + p * 4;
                    pre[p] = ep.load(row < M ? row : M - 1, hlf * (BN / 2) + c4);
                }
            }
        }
        if constexpr (ep_prefetch<EP>::value) {
            if ((int)blockIdx.x < ntiles) ep.prefetch_tile(blockIdx.x * TC_BM, M, tid);
        }
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int acc = it & 1, row0 = tile * TC_BM;
            ep.init(row0, M);
            if constexpr (ep_prefetch<EP>::value) {  // the next tile's epilogue inputs on their way to L2 while this one is processed
                if (tile + (int)gridDim.x < ntiles) ep.prefetch_tile((tile + (int)gridDim.x) * TC_BM, M, tid);
            }
            TCP_STAMP(0);
            mbar_wait(tmem_full + acc, (it >> 1) & 1);
            tc_fence_after();
            TCP_STAMP(1);
#pragma unroll 1
            for (int cb = 0; cb < BN / 64; ++cb) {
                const int col0 = hlf * (BN / 2) + cb * 32;
                // all global loads of this 32x32 block are issued first and stay in flight while the
                // accumulator block is read from TMEM and transposed through shared memory
                const int rowq0 = row0 + q * 32;
                if constexpr (AFF) ep.prep_block(rowq0, rsub, col0 + c4, M);
                else ep.prep(col0 + c4);
#ifdef RTFS_PROBE_NO_STAGE
                float pre_dummy = 0.f;
#endif
                if constexpr (!ROLL) {
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        const int row = rowq0 + rsub + p * 4;
                        if constexpr (AFF) {
                            if (row < M) pre[p] = ep.load_p(p);
                        } else {
                            pre[p] = ep.load(row < M ? row : M - 1, col0 + c4);
                        }
                    }
                }
                TCP_STAMP(2 + 5 * cb);  // loads issued
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t v[16];
                    tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + col0) + 16 * hh, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#ifdef RTFS_PROBE_NO_STAGE
                    pre_dummy += __uint_as_float(v[0]) + __uint_as_float(v[7]) + __uint_as_float(v[15]);
#else
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<float4*>(stg + lane * TC_STG_LD + 16 * hh + 4 * i) =
                            make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
#endif
                }
                __syncwarp();
                TCP_STAMP(3 + 5 * cb);  // accumulator block staged
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int r = p * 4 + rsub;
#ifdef RTFS_PROBE_NO_STAGE  // tools/probe/tcp_ablate.cu: timing without the staging reads (results meaningless)
                    const float4 x = make_float4(pre_dummy, pre_dummy, pre_dummy, pre_dummy);
#else
                    const float4 x = *reinterpret_cast<const float4*>(stg + r * TC_STG_LD + c4);
#endif
                    const int row = rowq0 + r;
                    if (row < M) {
                        if constexpr (FUSED) ep.store4s(row, col0 + c4, x, pre[p], stg + r * TC_STG_LD + c4);
                        else if constexpr (AFF) ep.store_p(p, row, x, pre[p]);
                        else ep.store4(row, col0 + c4, x, pre[p]);
                    }
                    if constexpr (ROLL) {
                        const bool lastb = cb + 1 == BN / 64;
                        const int ntile = lastb ? tile + (int)gridDim.x : tile;
                        if (ntile < ntiles) {
                            const int nrow = ntile * TC_BM + q * 32 + r;
                            pre[p] = ep.load(nrow < M ? nrow : M - 1, (lastb ? hlf * (BN / 2) : col0 + 32) + c4);
                        }
                    }
                }
                __syncwarp();
                TCP_STAMP(4 + 5 * cb);  // epilogue inputs arrived, block transformed
                if constexpr (FUSED) {  // (running the previous block's reduction under this block's loads was measured twice: the 8 Pre
                                        // register sets live across it spill at 104 and at 120 epilogue registers, 1.24 / 1.09 vs 0.74 ms)
                    ep.block_reduce(stg, lane, col0);
                    __syncwarp();
                }
                TCP_STAMP(5 + 5 * cb);  // fused row reduction done
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(tmem_empty + acc);  // this warp's TMEM reads of the accumulator are complete
            if constexpr (FUSED) ep.tile_done(stg_all, warp, lane, row0 + q * 32, M);
            ep.finish_group(scratch, tid, TCP_EPI, 1);
            TCP_STAMP(30);
        }
    } else if (warp < MMA_WARP) {
        if constexpr (REGSPLIT) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        // setmaxnreg moves registers inside the CTA's own launch allocation (26 warps x 72): 8 E + 16 p + 2 x 72 <= 26 x 72, i.e. the
        // producers get p = 108 - E / 2 (E = 104 -> 56, E = 120 -> 48; more deadlocks the epilogue's TRY_ALLOC).  At a register
        // prefetch depth of 4 chunks the 56-register producers spilled (SASS: STL / LDL only behind USETMAXREG.DEALLOC): such kernels
        // are launched with PF = 2.
        if constexpr (REGSPLIT16) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(108 - ep_epi_regs<EP>::value / 2));
        // ================================================================= A producers
        const int ptid = tid - TCP_EPI;
        const int kq = ptid & 7;
        unsigned char* a_dst0 = a_stage + kq * TC_LBO_A + (ptid >> 3) * 16;
        if constexpr (ASYNC == 1) {
            constexpr int D = NSA - 2 - (NSA >= 7 ? 2 : 0);  // chunks in flight ahead of the transform (slack for the commit -> empty round trip)
            static_assert(NSA >= 3, "async producers need >= 3 stages");
            const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            const int G = my_tiles * NK;
            auto issue = [&](int g) {
                if (g < G) {
                    const int it = g / NK, kc = g - it * NK;
                    const long long rbase = (long long)(blockIdx.x + it * gridDim.x) * TC_BM + (ptid >> 3);
                    const int s = g % NSA;
                    if (g >= NSA) mbar_wait(empty_a + s, ((g / NSA) - 1) & 1);
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        const long long row = rbase + RS * i;
                        const bool valid = row < M;
                        cp_async16(a_dst0 + (size_t)s * TC_A_STAGE + i * (RS * 16), al.raw(valid ? row : 0, kc * TC_KC + kq * 4), valid);
                    }
                }
                cp_async_commit();
            };
#pragma unroll
            for (int g = 0; g < D; ++g) issue(g);
            for (int g = 0; g < G; ++g) {
                issue(g + D);
                cp_async_wait<D>();  // this thread's pieces of chunk g have landed
                const int it = g / NK, kc = g - it * NK;
                const int row0 = (blockIdx.x + it * gridDim.x) * TC_BM;
                if (kc == 0) {
                    al.init_p(row0, M, extra + (it & 1) * AL::kExtra, ptid, TCP_PROD);
                    if (AL::kExtra > 0) named_bar_sync(2, TCP_PROD);
                }
                const int s = g % NSA;
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    float4* slot = reinterpret_cast<float4*>(a_dst0 + (size_t)s * TC_A_STAGE + i * (RS * 16));
                    const int row = row0 + (ptid >> 3) + RS * i;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row < M) v = al.xform(*slot, row, kc * TC_KC + kq * 4);
                    v.x = tf32r(v.x);
                    v.y = tf32r(v.y);
                    v.z = tf32r(v.z);
                    v.w = tf32r(v.w);
                    *slot = v;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cta(full_a + s);
            }
            cp_async_wait<0>();
        } else if constexpr (ASYNC == 2) {
            // The producers are ISSUE-bound, not latency-bound (tools/probe/tcp_ablate.cu: ~125 instructions per warp and
            // chunk, 16 warps -> the chunk period was ~1500 cycles whatever the memory did), so the loop is written
            // tile by tile with the K chunk as a compile-time index: ring slot arithmetic, table offsets and global
            // addresses become immediates off two per-tile row pointers, and the row-validity test is hoisted per tile.
            static_assert(NK % PF == 0, "flat register stream: PF must divide the number of K chunks");
            const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            float4 raw[PF][RPT];
            const float* pc[RPT];  // this thread's rows of the current tile (K offset kq*4), and of the next one
            const float* pn[RPT];
            bool vc[RPT], vn[RPT];
            auto rows_of = [&](int it, const float* (&p)[RPT], bool (&v)[RPT]) {
                const long long rbase = (long long)(blockIdx.x + it * gridDim.x) * TC_BM + (ptid >> 3);
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const long long row = rbase + RS * i;
                    v[i] = it < my_tiles && row < M;
                    p[i] = al.raw(v[i] ? row : 0, kq * 4);
                }
            };
            rows_of(0, pc, vc);
#pragma unroll
            for (int c = 0; c < PF; ++c) {
#pragma unroll
                for (int i = 0; i < RPT; ++i) raw[c][i] = vc[i] ? ldg4(pc[i] + c * TC_KC) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            int s = 0;          // ring slot of the next chunk, and the parity its empty barrier is waited with
            uint32_t eph = 1;   // (first pass over the ring: nothing to wait for)
            bool first_pass = true;
            for (int it = 0; it < my_tiles; ++it) {
                const int row0 = (blockIdx.x + it * gridDim.x) * TC_BM;
                rows_of(it + 1, pn, vn);
                if (!loader_tile_invariant<AL>::value || it == 0) {
                    al.init_p(row0, M, extra + (loader_tile_invariant<AL>::value ? 0 : (it & 1)) * AL::kExtra, ptid, TCP_PROD);
                    if (AL::kExtra > 0) named_bar_sync(2, TCP_PROD);
                }
#pragma unroll
                for (int kc = 0; kc < NK; ++kc) {
                    if (!first_pass) mbar_wait(empty_a + s, eph);
                    // rows past M are transformed like any other (zeros in): their accumulator rows are never stored, and
                    // transforming all rows before the first store lets the per-channel table loads be shared between them
                    float4 v[RPT];
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        v[i] = al.xform(raw[kc % PF][i], row0 + (ptid >> 3) + RS * i, kc * TC_KC + kq * 4);
                        v[i].x = tf32r_fast(v[i].x);
                        v[i].y = tf32r_fast(v[i].y);
                        v[i].z = tf32r_fast(v[i].z);
                        v[i].w = tf32r_fast(v[i].w);
                    }
#pragma unroll
                    for (int i = 0; i < RPT; ++i) *reinterpret_cast<float4*>(a_dst0 + (size_t)s * TC_A_STAGE + i * (RS * 16)) = v[i];
                    fence_proxy_async();  // before the next loads are issued
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta(full_a + s);
                    if (++s == NSA) {
                        s = 0;
                        eph = first_pass ? 0u : eph ^ 1u;
                        first_pass = false;
                    }
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        if (kc + PF < NK) raw[kc % PF][i] = vc[i] ? ldg4(pc[i] + (kc + PF) * TC_KC) : make_float4(0.f, 0.f, 0.f, 0.f);
                        else raw[kc % PF][i] = vn[i] ? ldg4(pn[i] + (kc + PF - NK) * TC_KC) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    pc[i] = pn[i];
                    vc[i] = vn[i];
                }
            }
        } else {
        int it = 0, ga = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int row0 = tile * TC_BM;
            al.init_p(row0, M, extra + (it & 1) * AL::kExtra, ptid, TCP_PROD);
            if (AL::kExtra > 0) named_bar_sync(2, TCP_PROD);  // the tile's loader table is complete
            if constexpr (loader_batched<AL>::value && TCP_BATCHED) {
                // several operand loads per element (TF-AR combine): one chunk's raw loads are issued as a batch and
                // transformed afterwards (see gemm_tc_kernel), the next chunk's batch flies behind the stores
                typename AL::Raw raw[RPT];
#pragma unroll
                for (int i = 0; i < RPT; ++i) raw[i] = al.raw_load(i, kq * 4);
#pragma unroll
                for (int kc = 0; kc < NK; ++kc, ++ga) {
                    const int s = ga % NSA, use = ga / NSA;
                    if (use > 0) mbar_wait(empty_a + s, (use - 1) & 1);
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        float4 v = al.xform_raw(raw[i], i, kc * TC_KC + kq * 4);
                        v.x = tf32r_fast(v.x);
                        v.y = tf32r_fast(v.y);
                        v.z = tf32r_fast(v.z);
                        v.w = tf32r_fast(v.w);
                        *reinterpret_cast<float4*>(a_dst0 + (size_t)s * TC_A_STAGE + i * (RS * 16)) = v;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta(full_a + s);
                    if (kc + 1 < NK) {
#pragma unroll
                        for (int i = 0; i < RPT; ++i) raw[i] = al.raw_load(i, (kc + 1) * TC_KC + kq * 4);
                    }
                }
                continue;
            }
            float4 areg[PF][RPT];
#pragma unroll
            for (int c = 0; c < PF; ++c) {
#pragma unroll
                for (int i = 0; i < RPT; ++i) areg[c][i] = al.load(i, c * TC_KC + kq * 4);
            }
#pragma unroll
            for (int kc = 0; kc < NK; ++kc, ++ga) {
                const int s = ga % NSA, use = ga / NSA;
                if (use > 0) mbar_wait(empty_a + s, (use - 1) & 1);
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    float4 v = areg[kc % PF][i];
                    v.x = tf32r(v.x);
                    v.y = tf32r(v.y);
                    v.z = tf32r(v.z);
                    v.w = tf32r(v.w);
                    *reinterpret_cast<float4*>(a_dst0 + (size_t)s * TC_A_STAGE + i * (RS * 16)) = v;
                }
                fence_proxy_async();  // before the next loads are issued
                __syncwarp();
                if (lane == 0) mbar_arrive_cta(full_a + s);
                if (kc + PF < NK) {
#pragma unroll
                    for (int i = 0; i < RPT; ++i) areg[kc % PF][i] = al.load(i, (kc + PF) * TC_KC + kq * 4);
                }
            }
        }
        }
    } else if (warp == MMA_WARP) {
        // ================================================================= MMA issuer
        if (elect_one()) {
            int it = 0, ga = 0;
#ifdef RTFS_PROBE_TIMING  // tools/probe/tcp_ablate.cu: cycles the MMA thread waits on each barrier kind
            long long t_te = 0, t_fw = 0, t_fa = 0, t0 = clock64(), tq;
#define PROBE_T(acc_, stmt) tq = clock64(); stmt; acc_ += clock64() - tq
#else
#define PROBE_T(acc_, stmt) stmt
#endif
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                if (it >= 2) { PROBE_T(t_te, mbar_wait(tmem_empty + acc, ((it >> 1) - 1) & 1)); }
                tc_fence_after();
#pragma unroll 1
                for (int kc = 0; kc < NK; ++kc, ++ga) {
                    const int s = ga % NSA;
                    const int ws = WRES ? kc : ga % NSW;
                    PROBE_T(t_fw, mbar_wait(full_w + ws, WRES ? 0 : (ga / NSW) & 1));
                    PROBE_T(t_fa, mbar_wait(full_a + s, (ga / NSA) & 1));
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(a_stage + (size_t)s * TC_A_STAGE);
                    const uint32_t w_base = smem_u32(w_stage + (size_t)ws * WBYTES);
#pragma unroll
                    for (int k8 = 0; k8 < TC_KC / 8; ++k8) {
                        const uint64_t da = umma_desc(a_base + 2 * k8 * TC_LBO_A, TC_LBO_A, 128);
                        const uint64_t db = umma_desc(w_base + 2 * k8 * (BN * 16), BN * 16, 128);
                        umma_tf32(tmem + acc * BN, da, db, IDESC, (kc > 0 || k8 > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_a + s);
                    if (!WRES) umma_commit(empty_w + ws);
                }
                umma_commit(tmem_full + acc);
            }
#ifdef RTFS_PROBE_TIMING
            if (blockIdx.x < 1) printf("cta %d mma thread: total %lld cycles, wait tmem_empty %lld, full_w %lld, full_a %lld (tiles %d)\n", blockIdx.x, clock64() - t0, t_te, t_fw, t_fa, it);
#endif
        }
    } else {
        // ================================================================= weight loader
        if (elect_one()) {
            if (WRES) {
                for (int c = 0; c < NK; ++c) {
                    mbar_expect_tx(full_w + c, WBYTES);
                    bulk_g2s(w_stage + (size_t)c * WBYTES, Wimg + (size_t)c * (BN * TC_KC), WBYTES, full_w + c);
                }
            } else {
                int gw = 0;
                for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                    for (int kc = 0; kc < NK; ++kc, ++gw) {
                        const int ws = gw % NSW, use = gw / NSW;
                        if (use > 0) mbar_wait(empty_w + ws, (use - 1) & 1);
#ifdef RTFS_PROBE_W_ONCE  // tools/probe/tcp_ablate.cu: timing without the weight stream (stale slabs, results meaningless)
                        if (gw >= NSW) {
                            mbar_arrive_cta(full_w + ws);
                            continue;
                        }
#endif
                        mbar_expect_tx(full_w + ws, WBYTES);
                        bulk_g2s(w_stage + (size_t)ws * WBYTES, Wimg + (size_t)kc * (BN * TC_KC), WBYTES, full_w + ws);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<(2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512)>(tmem);
}

template <int BN, int KTOT, int NSA, int NSW, bool WRES, int PF, int ASYNC, int NPROD, class AL, class EP>
inline cudaError_t launch_gemm_tcp(const AL& al, const float* Wimg, const EP& ep, int M, cudaStream_t st) {
    auto kern = gemm_tcp_kernel<BN, KTOT, NSA, NSW, WRES, PF, ASYNC, NPROD, AL, EP>;
    const int smem = tcp_smem_bytes<BN, KTOT, NSA, NSW, WRES>(AL::kExtra) + ep_smem_bytes<EP>::value;
    static SmemCfg cfg;  // per instantiation, per device
    if (cudaError_t e = ensure_smem(kern, smem, cfg); e != cudaSuccess) return e;
    const int ntiles = (M + TC_BM - 1) / TC_BM;
    const int grid = ntiles < sm_count() ? ntiles : sm_count();
    kern<<<grid, TCP_EPI + NPROD + 64, smem, st>>>(al, Wimg, ep, M, ntiles);
    return cudaGetLastError();
}

}  // namespace rtfs
