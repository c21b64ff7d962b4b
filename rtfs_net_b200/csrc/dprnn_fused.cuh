// Fused dual-path RNN (reference: DualPathRNN.forward, layers/rnn_layers.py:136-162; SRU recurrence per the
// third-party `sru` package).  One kernel per path replaces prep + 5 GEMMs + 4 scans:
//
//   g (B,Tc,Fc,64) --LN over C--> n --unfold(8) . W0--> U0 --scan--> h0 --W1--> U1 --scan--> ... h3
//     --ConvTranspose1d(64,64,8) + bias + residual--> g'
//
// A tile = whole sequences packed back to back: 128 positions (2 x 64 frequency bins of one frame each / 1 x 125 frames of one
// bin) for sequences of up to 128 steps, two such tiles in flight per persistent CTA (template DUAL, see the kernel's comment);
// 256 positions, one tile per CTA, for longer sequences.  Everything between the load of g and the store of g' stays on chip:
//   * the activation slab (n, then h0..h3 in place) lives in shared memory in the UMMA K-major no-swizzle
//     layout with 7 guard rows either side; nn.Unfold(8) and the k=8 transposed conv are the SAME slab read
//     through descriptors whose start address is advanced by `tap` rows (16 bytes each) -- no im2col;
//   * the SRU GEMMs run transposed (features on the 128 TMEM lanes, positions on the columns):
//         acc0 lanes 0-63 = candidate, 64-127 = reset pre-activation ; acc1 lanes 0-63 = forget pre-activation,
//         64-127 = highway projection (layer 0)
//     so the serial recurrence reads its own TMEM lane, 16 time steps per tcgen05.ld: warps (4s+0, 4s+1) run
//     the c-recurrence of sequence s (forward / backward halves), warps (4s+2, 4s+3) then form h in parallel;
//   * weights stream as 52 whole 16 KB slabs per tile (one cp.async.bulk, one wait, one commit each) through a ring that layer 0
//     extends into the idle c buffer, dealt to four producer warps of which each slot has exactly one (DfC, df_slab_slot);
//   * tcgen05.mma / commit are issued by the elect.sync lane of warp 0 with descriptors advanced by integer adds (gemm_tc.cuh:
//     elect_one), and every waiting warp parks on its mbarrier (suspend-time hint);
//   * the transposed conv runs with positions on the lanes, so its epilogue (bias + residual + store) is the
//     coalesced transposed-through-shared-memory store of gemm_tc.cuh.
// Measurements behind these choices: profiles/r02_df_probes.txt (tools/probe/dfgemm_probe.cu, dfissue_probe.cu); DESIGN.md 4.2.
#pragma once
#include <cstdlib>
#include "common.cuh"
#include "gemm_tc.cuh"

namespace rtfs {

constexpr int DF_NP = 256;                 // largest tile (positions); sequences longer than this take the unfused path
constexpr int DF_SLAB = 16384;             // one weight slab of the image (weights.py: dprnn_fused_image) = one bulk copy
constexpr int DF_NSLAB = 32 + 3 * 4 + 8;   // layer 0 | layers 1-3 | transposed conv
constexpr int DF_NPROD = 4;                // weight-producer warps

// Tile configuration.  NP = 256: one CTA of 512 threads per SM.  NP = 128 (sequences of up to 128 steps): 256 threads and 102 KB of
// shared memory per pipeline, so two tiles are resident per SM (2 x 256 TMEM columns) and one tile's serial recurrence / load /
// store phases run under the other tile's UMMAs.
template <int NP>
struct DfC {
    static constexpr int NT = 2 * NP;              // threads: 4 warps per 64 positions
    static constexpr int NG = NT / 128;            // warp groups (one sequence each)
    static constexpr int ROWS = 7 + NP + 7;        // slab rows (guard rows hold zeros)
    static constexpr int LBO = ROWS * 16 + 16;     // bytes between 4-channel pieces (4336 / 2288: 112 mod 128)
    static constexpr int HBUF = 16 * LBO;
    static constexpr int CS = NP * 64 * 4;         // c_t of every (position, column)
    // Weight ring.  A bulk copy lands ~1000 cycles after it is issued and costs its issuing thread 300-500 cycles whatever its size, and
    // a wait that has to suspend takes ~200 cycles to wake: the issuing thread must find its weights already landed.  So weights move
    // as whole 16 KB slabs, dealt to four producer warps.  The ring proper holds NRING slabs; the c buffer directly below it is as
    // large and idle until the first recurrence, so layer 0 -- 512 KB of weights per tile -- cycles through c buffer + ring
    // (NSL = 2 NRING slab slots), the later layers through the ring.
    static constexpr int NRING = CS / DF_SLAB;     // 2 (NP = 128) or 4
    static constexpr int NSL = 2 * NRING;
    static constexpr int SMEM = HBUF + CS + NRING * DF_SLAB + 2 * NP * 4 + 256 + 8 * 16 * 8;  // one pipeline
    static constexpr int MINB = NP == 256 ? 1 : 2;
    static_assert(CS == NRING * DF_SLAB && (NRING == 2 || NRING == 4), "the c buffer mirrors the ring");
};
// Slab slot ss (0 .. NRING-1: c buffer, NRING .. NSL-1: ring; also the index of its full_w / mma_done barriers), wait parity and owning
// producer of weight slab i of tile `it` of the pipeline.  Every slot is refilled by ONE producer, in order: a parity wait is only
// sound while the waiter is at most one phase behind, which a single in-order owner is by construction.
//   NP = 128: ONE 4-slot ring over the whole tile, ss = (i + 2) & 3 -- slabs 0, 1 of the tile, of every later layer and of the
//     transposed conv fall on the ring half (fetched ahead, while the c buffer holds c values or stages the epilogue), slabs 2, 3 of a
//     later layer on the c-buffer half (requested the moment that layer's GEMM phase begins: the h phase has just released the
//     buffer; they used to wait for slabs 0, 1 to be multiplied first).  52 slabs = 13 uses per slot and tile: parities flip per tile.
//   NP = 256 (one tile per CTA): layer 0 through all 8 slots, the later layers through the ring's 4 (a whole layer fits).
template <int NP>
DEVINL void df_slab_slot(int i, int it, int& ss, uint32_t& par, int& owner) {
    using C = DfC<NP>;
    if (NP == 128) {
        static_assert(NP != 128 || (DF_NSLAB % 4 == 0 && (DF_NSLAB / 4) % 2 == 1), "13 uses per slot and tile");
        ss = (i + 2) & 3;
        par = (uint32_t)((i >> 2) + it) & 1u;
        owner = ss;
    } else if (i < 32) {
        ss = (i + C::NRING) & (C::NSL - 1);
        par = (uint32_t)(i / C::NSL) & 1u;
        owner = i & 3;  // = the owner of the later layers' slot when ss is a ring slot
    } else {
        const int v = i - 32;
        ss = C::NRING + (v & (C::NRING - 1));
        par = (uint32_t)(v / C::NRING) & 1u;  // an even number of layer-0 uses came before
        owner = v & (C::NRING - 1);
    }
}

struct DfArgs {
    const float* g_in;    // (B,Tc,Fc,64) when first == 0
    const float* d1_pre;  // first == 1: g = gLN(d1_pre) + pool (tdanet.py:117-118), also written to g_first
    const float* pool;
    GlnRef gln;
    float* g_first;
    const float* ln_gamma;  // [64]
    const float* ln_beta;
    const float* wimg;      // DF_NSLAB slabs of 4096 floats (weights.py: dprnn_fused_image)
    const float* wc[4];     // [128] = v_f | v_r per layer
    const float* bias[4];   // [128] = b_f | b_r
    const float* ct_bias;   // [64]
    float* g_out;
    int B, Tc, Fc;
    int time_path;  // 0: sequences = (b,t), scanned axis f ; 1: sequences = (b,f), scanned axis t
    int first;
    int S, L;       // sequence length, SRU steps (S - 7)
    int nseq_total, nseq_tile, n_other;
    int yield_sms;   // a side-stream kernel is in flight (forked VP block): run one-tile CTAs, which leave SMs as they finish
    long long* dbg;  // optional: [tiles][16] clock64 stamps of thread 0 at the phase boundaries
};

DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr float DF_NL2E = -1.4426950408889634f;
// MUFU approximations without the denormal fix-ups of exp2f / division: 2^t flushes to 0 / overflows to inf
// at the ends, which is exactly sigmoid's 1 / 0 limit
DEVINL float df_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
DEVINL float df_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
DEVINL float df_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Gate arithmetic of the recurrences (template parameter GATE of the kernel):
//   0: sigmoid(x) = 1 / (1 + 2^(-x log2 e))   MUFU.EX2 + MUFU.RCP (2 quarter-rate ops per gate)
//   1: sigmoid(x) = 0.5 + 0.5 tanh(x/2)       one MUFU.TANH (max rel. error 2^-11 on tanh)
//   2: as 0 with the reciprocal done by a bit-trick seed + 3 Newton steps on the FMA pipe (1 MUFU per gate)
template <int GATE>
DEVINL constexpr float df_prescale() { return GATE == 1 ? 0.5f : DF_NL2E; }
DEVINL float df_rcp_newton(float d) {  // d in [1, inf): 1/d to ~1e-7 relative
    float r = __uint_as_float(0x7EF311C3u - __float_as_uint(d));
    r = r * fmaf(-d, r, 2.f);
    r = r * fmaf(-d, r, 2.f);
    r = r * fmaf(-d, r, 2.f);
    return r;
}
template <int GATE>
DEVINL float df_gate(float t) {  // t = pre-scaled pre-activation
    if (GATE == 1) return fmaf(0.5f, df_tanh(t), 0.5f);
    const float d = 1.f + df_ex2(t);
    if (GATE == 2) return d < 1e30f ? df_rcp_newton(d) : 0.f;
    return df_rcp(d);
}
// one step of c_t = f_t c_{t-1} + (1 - f_t) u0_t, f_t = sigmoid(u1_t + v_f c_{t-1} + b_f); vf and u1 are pre-scaled
template <bool PRED, int GATE>
DEVINL float df_cstep(float c, float vf, float u1, float u0, float* dst, bool valid) {
    float cn;
    if (GATE == 1) {  // c_t = A + tanh(.) * B with A = (c + u0)/2, B = (c - u0)/2 formed beside the MUFU op
        const float th = df_tanh(fmaf(vf, c, u1));
        cn = fmaf(th, 0.5f * (c - u0), 0.5f * (c + u0));
    } else {
        const float f = df_gate<GATE>(fmaf(vf, c, u1));
        cn = fmaf(f, c - u0, u0);
    }
    if (PRED) {
        if (valid) *dst = cn;
        return valid ? cn : c;
    }
    *dst = cn;
    return cn;
}

// h_t = r_t c_t + (1 - r_t) x'_t, r_t = sigmoid(u2_t + v_r c_{t-1} + b_r) for the 16 steps of TMEM column block m.
// FULL: all 16 steps lie inside [p_lo, p_hi) (no per-step predicates); K4: x' = highway projection (acc1), else the
// previous layer's h read in place from the slab.  rev: scan order is descending (c_{t-1} is the row above).
template <bool FULL, bool K4, int GATE, int NP>
DEVINL void df_hchunk(uint32_t tl, int m, bool rev, int p_lo, int p_hi, int p_end, float vr, float br, const float* csj, unsigned char* hb) {
    uint32_t ua[16], ub[16];
    tmem_ld16(tl + 16 * m, ua);
    if (K4) tmem_ld16(tl + NP + 16 * m, ub);
    const int p0 = 16 * m;
    float cc[18], xp[16];  // cc[i + 1] = c at step p0 + i ; cc[0], cc[17] = neighbours (0 outside the sequence)
#pragma unroll
    for (int i = -1; i <= 16; ++i) {
        const int p = p0 + i;
        const bool in = FULL ? ((i >= 0 && i < 16) || (p >= p_lo && p < p_hi)) : (p >= p_lo && p < p_hi);
        cc[i + 1] = in ? csj[p * 64] : 0.f;
    }
    if (!K4) {
#pragma unroll
        for (int i = 0; i < 16; ++i) xp[i] = (FULL || (p0 + i >= p_lo && p0 + i < p_hi)) ? *reinterpret_cast<const float*>(hb + (p0 + i) * 16) : 0.f;
    }
    tmem_ld_wait();
    const float brs = br * df_prescale<GATE>();
    float hv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float x = K4 ? __uint_as_float(ub[i]) : xp[i];
        const float cprev = rev ? cc[i + 2] : cc[i];
        const float r = df_gate<GATE>(fmaf(vr, cprev, fmaf(__uint_as_float(ua[i]), df_prescale<GATE>(), brs)));
        hv[i] = tf32r_fast(fmaf(r, cc[i + 1] - x, x));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int p = p0 + i;
        if (FULL || (p >= p_lo && p < p_hi)) *reinterpret_cast<float*>(hb + p * 16) = hv[i];
        else if (p >= p_hi && p < p_end) *reinterpret_cast<float*>(hb + p * 16) = 0.f;
    }
}

// (weight-producer warps: one thread sustains one cp.async.bulk per ~300-500 cycles whatever its size, so DF_NPROD warps share the stream)
DEVINL void df_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// DUAL = false: one tile per CTA, grid = tiles (NP = 256: one CTA per SM; NP = 128: two).
// DUAL = true (NP = 128): a persistent CTA of 512 threads runs TWO independent half-tile pipelines (threads 0-255 / 256-511, each
// with its own slab, c buffer, weight ring, barriers and 256 TMEM columns) over tiles 2 (blockIdx + it gridDim) + half.  The halves
// take turns on the long layer-0 GEMM (l0_done hand-off), which keeps them half a tile apart for the whole kernel: one half's
// serial recurrences, loads and stores run under the other half's UMMAs, and TMEM allocation / barrier set-up happen once per CTA.
template <int GATE, int NP, bool DUAL>
__global__ void __launch_bounds__(DUAL ? 512 : DfC<NP>::NT, DUAL ? 1 : DfC<NP>::MINB) dprnn_fused_kernel(DfArgs a) {
    using C = DfC<NP>;
    static_assert(!DUAL || NP == 128, "the two-pipeline kernel runs 128-position tiles");
    constexpr int DF_NT = C::NT, DF_LBO = C::LBO, DF_HBUF = C::HBUF, DF_CS = C::CS;
    constexpr int NW = DF_NT / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int half = DUAL ? (int)(threadIdx.x >> 8) : 0;
    const int tid = DUAL ? (int)(threadIdx.x & 255) : (int)threadIdx.x;  // within the pipeline
    unsigned char* const sm = smem_raw + half * C::SMEM;
    unsigned char* hbuf = sm;
    float* cs = reinterpret_cast<float*>(sm + DF_HBUF);
    unsigned char* wring = sm + DF_HBUF + DF_CS;
    int* pos2off = reinterpret_cast<int*>(wring + C::NRING * DF_SLAB);  // [2][NP]: double-buffered over tiles
    uint64_t* full_w = reinterpret_cast<uint64_t*>(pos2off + 2 * NP);
    uint64_t* mma_done = full_w + 8;
    uint64_t* acc_ready = mma_done + 8;
    static_assert((2 * 8 + 1) * 8 + 32 <= 256, "barriers + seqstat fit their 256 bytes");
    float* seqstat = reinterpret_cast<float*>(acc_ready + 1);  // [4 sequences][mean, rstd] of the gLN applied when first
    // chunk_bar[(sequence, direction)][16-step block]: c-recurrence warp -> h warp hand-off, one completion per layer
    uint64_t* chunk_bar = reinterpret_cast<uint64_t*>(sm + DF_HBUF + DF_CS + C::NRING * DF_SLAB + 2 * NP * 4 + 256);
    // shared by both pipelines: the layer-0 hand-off barriers and the TMEM base
    uint64_t* l0_done = reinterpret_cast<uint64_t*>(smem_raw + (DUAL ? 2 : 1) * C::SMEM);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(l0_done + 2);

    const int lane = tid & 31, warp = tid >> 5;
    const int q = warp & 3, sw = warp >> 2;  // q is also the TMEM lane quarter of the hardware warp (8 warps per pipeline)
    const int S = a.S, L = a.L;
    const int ntiles = (a.nseq_total + a.nseq_tile - 1) / a.nseq_tile;
    const int niter = DUAL ? (ntiles + 2 * (int)gridDim.x - 1) / (2 * (int)gridDim.x) : 1;
    int dbg_i = 0;
    bool dbg_on = a.dbg != nullptr && tid == 0 && !DUAL;
#define DF_STAMP()                                                                           \
    do {                                                                                     \
        if (dbg_on && dbg_i < 16) a.dbg[(blockIdx.x * (DUAL ? 2 : 1) + half) * 16 + dbg_i++] = clock64(); \
    } while (0)
#define DF_SYNC()                        \
    do {                                 \
        if (DUAL) df_bar(1 + half, 256); \
        else __syncthreads();            \
    } while (0)
    DF_STAMP();

    // barrier initialisation before anything is in flight: its release fence would otherwise wait for the loads below
    if (tid == 32) {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            mbar_init(full_w + s, 1);
            mbar_init(mma_done + s, 1);
        }
        mbar_init(acc_ready, 1);
        if (half == 0) {
            mbar_init(l0_done, 1);
            mbar_init(l0_done + 1, 1);
        }
        fence_mbar_init();
    }
    if (tid >= 128 && tid < 256) {  // the 128 hand-off barriers, one per thread (a single thread took ~1.5 k cycles per tile)
        mbar_init(chunk_bar + (tid - 128), 1);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) tmem_alloc<DUAL ? 512 : 2 * NP>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot + (uint32_t)(half * 2 * NP);

    // ---- weight producers (warps 1..DF_NPROD of the pipeline): slab n (counted over all tiles) -> df_slab_slot(n % DF_NSLAB); every
    //      producer scans the slabs in order and issues the ones it owns.  Called by the whole warp, one lane issues.
    const int wtotal = niter * DF_NSLAB;
    const bool is_prod = warp >= 1 && warp <= DF_NPROD;
    int wnext = 0, wi = 0, wt = 0;  // next slab to look at: index over all tiles / inside the tile, its tile
    auto produce_until = [&](int limit) {
        if (limit > wtotal) limit = wtotal;
        const bool lead = elect_one();
        for (; wnext < limit; ++wnext) {
            if (lead) {
                int ss, owner;
                uint32_t par;
                df_slab_slot<NP>(wi, wt, ss, par, owner);
                if (owner == warp - 1) {
                    mbar_wait(mma_done + ss, par ^ 1u);  // the previous use has been multiplied (passes at once on a fresh barrier)
                    mbar_expect_tx(full_w + ss, DF_SLAB);
                    bulk_g2s(reinterpret_cast<unsigned char*>(cs) + ss * DF_SLAB, a.wimg + (size_t)wi * (DF_SLAB / 4), DF_SLAB, full_w + ss);
                }
            }
            if (++wi == DF_NSLAB) {
                wi = 0;
                ++wt;
            }
        }
        __syncwarp();
    };
    if (is_prod) produce_until(C::NRING);

    const int l16 = tid & 15, c = l16 * 4;
    // descriptors are advanced by integer adds on the (address >> 4) field: one 16-byte slab row = 1, one 4-channel piece = LBO / 16
    const uint64_t d_slab = umma_desc(smem_u32(hbuf), DF_LBO, 128);       // slab rows as the positions operand
    const uint64_t d_wsru = umma_desc(smem_u32(cs), 2048, 128);          // SRU weights (slot 0): 128 features x 4-channel pieces
    const uint64_t d_wct = umma_desc(smem_u32(cs), 1024, 128);           // transposed-conv weights: 64 outputs x 4-channel pieces
    constexpr uint32_t PIECE = DF_LBO / 16, SLAB16 = DF_SLAB / 16;
    constexpr uint32_t IDESC_SRU = umma_idesc_tf32(128, NP);
    constexpr uint32_t IDESC_CT = umma_idesc_tf32(128, 64);
    const int p_lo = sw * S, p_hi = p_lo + L;
    const bool seq_on = sw < a.nseq_tile;
    int nacc = 0;  // completions of acc_ready waited for so far (5 per tile)

#pragma unroll 1
    for (int it = 0; it < niter; ++it) {
        const int tile = DUAL ? (it * (int)gridDim.x + (int)blockIdx.x) * 2 + half : (int)blockIdx.x;  // tiles past the end run empty
        const int seq0 = tile * a.nseq_tile;
        int* p2o = pos2off + (it & 1) * NP;
        const int ibase = it * DF_NSLAB;  // slabs before this tile
        if (DUAL) {
            dbg_on = a.dbg != nullptr && tid == 0 && it == (niter > 1 ? 1 : 0);
            dbg_i = 0;
            DF_STAMP();
        }

        // ---- P0 loads first: every thread derives the global offsets of its 8 rows (positions p = (NT/16) i + tid/16, channel quad
        //      tid%16) incrementally and has them in flight while the tables below are set up
        float4 v[8], plv[8];
        int offs[8];
        {
            const float* src = a.first ? a.d1_pre : a.g_in;
            const float* src2 = a.first ? a.pool : a.g_in;  // second stream only read when first
            int p = tid >> 4;
            int s = p / S, l = p - s * S;
            int seqg = seq0 + s;
            int b = seqg / a.n_other, o = seqg - b * a.n_other;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int off = -1;
                if (s < a.nseq_tile && seqg < a.nseq_total) {
                    const int t = a.time_path ? l : o, f = a.time_path ? o : l;
                    off = ((b * a.Tc + t) * a.Fc + f) * 64;
                }
                offs[i] = off;
                v[i] = plv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (off >= 0) {
                    v[i] = ldg4(src + off + c);
                    if (a.first) plv[i] = ldg4(src2 + off + c);
                }
                if (l16 == 0) p2o[p] = off;  // the position -> offset table of the epilogue
                p += DF_NT / 16;
                l += DF_NT / 16;
                while (l >= S) {
                    l -= S;
                    ++s;
                    ++seqg;
                    if (++o == a.n_other) {
                        o = 0;
                        ++b;
                    }
                }
            }
        }
        if (a.first && tid >= 64 && tid < 68) {
            const int seqg = seq0 + (tid - 64);
            float mean = 0.f, rstd = 0.f;
            if (seqg < a.nseq_total) gln_mean_rstd(a.gln.sums, seqg / a.n_other, a.gln.inv_n, mean, rstd);
            seqstat[2 * (tid - 64)] = mean;
            seqstat[2 * (tid - 64) + 1] = rstd;
        }
        DF_SYNC();  // the previous tile's epilogue has left its staging rows (slab tail + c buffer); seqstat is visible
        // guard rows (7 before, 7 after) of every 4-channel piece (the epilogue staging overlaps the slab's tail, so every tile)
        for (int i = tid; i < 14 * 16; i += DF_NT) {
            const int kq = i / 14, rr = i - kq * 14;
            const int row = rr < 7 ? rr : NP + rr;
            *reinterpret_cast<float4*>(hbuf + kq * DF_LBO + row * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        DF_STAMP();

        // ---- P0: g -> LayerNorm over C -> n (tf32) into the slab; all 8 rows of a thread are in flight together
        {
            const float4 gm = ldg4(a.ln_gamma + c), be = ldg4(a.ln_beta + c);
            if (a.first) {  // g = gLN(d1_pre) + pool, written out as the residual / next stage input
                const float4 gg = ldg4(a.gln.gamma + c), gb = ldg4(a.gln.beta + c);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (offs[i] >= 0) {
                        const float4 pl = plv[i];
                        const int sq = (i * (DF_NT / 16) + (tid >> 4)) / S;
                        const float mean = seqstat[2 * sq], rstd = seqstat[2 * sq + 1];
                        float4 x = v[i];
                        x.x = (x.x - mean) * rstd * gg.x + gb.x + pl.x;
                        x.y = (x.y - mean) * rstd * gg.y + gb.y + pl.y;
                        x.z = (x.z - mean) * rstd * gg.z + gb.z + pl.z;
                        x.w = (x.w - mean) * rstd * gg.w + gb.w + pl.w;
                        *reinterpret_cast<float4*>(a.g_first + offs[i] + c) = x;
                        v[i] = x;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int p = i * (DF_NT / 16) + (tid >> 4);
                const float4 x = v[i];
                float s = x.x + x.y + x.z + x.w;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const float mu = s * (1.f / 64.f);
                const float dx = x.x - mu, dy = x.y - mu, dz = x.z - mu, dw = x.w - mu;
                float qq = dx * dx + dy * dy + dz * dz + dw * dw;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
                const float rs = rsqrtf(qq * (1.f / 64.f) + RTFS_EPS);
                float4 n;
                n.x = tf32r_fast(dx * rs * gm.x + be.x);
                n.y = tf32r_fast(dy * rs * gm.y + be.y);
                n.z = tf32r_fast(dz * rs * gm.z + be.z);
                n.w = tf32r_fast(dw * rs * gm.w + be.w);
                if (offs[i] < 0) n = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(hbuf + l16 * DF_LBO + (7 + p) * 16) = n;
            }
        }
        fence_proxy_async();
        tc_fence_before();
        DF_SYNC();
        tc_fence_after();
        DF_STAMP();

        int cbeg = 0;  // slab index inside the tile
#pragma unroll 1
        for (int ly = 0; ly < 4; ++ly) {
            const int nch = ly == 0 ? 32 : 4;
            // ---- GEMM: U^T[features][positions]
            if (warp == 0) {
                if (elect_one()) {
                    if (DUAL && ly == 0) {  // the pipelines alternate on the layer-0 GEMM: 0, 1, 0, 1, ...
                        if (half == 1) mbar_wait(l0_done, it & 1);
                        else if (it > 0) mbar_wait(l0_done + 1, (it - 1) & 1);
                    }
                    for (int gl = 0; gl < nch; ++gl) {
                        const uint32_t tap = ly == 0 ? (uint32_t)(gl >> 2) : 0u;
                        const uint32_t cb = (uint32_t)(ly == 0 ? (gl & 3) : gl) * 4u;
                        const uint64_t db = d_slab + (uint64_t)(cb * PIECE + 7u + tap);
                        const uint32_t acc_on = gl > 0 ? 1u : 0u;
                        int ss, owner;
                        uint32_t par;
                        df_slab_slot<NP>(cbeg + gl, it, ss, par, owner);
                        mbar_wait(full_w + ss, par);
                        tc_fence_after();
                        const uint64_t da = d_wsru + (uint64_t)((uint32_t)ss * SLAB16);  // [accumulator 0 | 1][16 K-channels][128 features]
                        umma_tf32(tmem, da, db, IDESC_SRU, acc_on);
                        umma_tf32(tmem + NP, da + 512u, db, IDESC_SRU, acc_on);
                        umma_tf32(tmem, da + 256u, db + 2u * PIECE, IDESC_SRU, 1u);
                        umma_tf32(tmem + NP, da + 768u, db + 2u * PIECE, IDESC_SRU, 1u);
                        umma_commit(mma_done + ss);
                    }
                    if (DUAL && ly == 0) umma_commit(l0_done + half);
                    umma_commit(acc_ready);
                }
                __syncwarp();
            } else if (is_prod) {
                produce_until(ibase + cbeg + nch + C::NRING);
            }
            cbeg += nch;
            mbar_wait(acc_ready, nacc & 1);
            ++nacc;
            tc_fence_after();
            DF_STAMP();

            // ---- c-recurrence: warps q = 0 (forward columns 0-31) and q = 1 (backward columns 32-63)
            if (q < 2 && seq_on) {
                const int j = q * 32 + lane;
                const float vf = __ldg(a.wc[ly] + j) * df_prescale<GATE>(), bf = __ldg(a.bias[ly] + j);
                const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
                float cc = 0.f;
                float* csj = cs + j;
                const int m_lo = p_lo >> 4, m_hi = (p_hi - 1) >> 4;
                for (int mm = m_lo; mm <= m_hi; ++mm) {
                    const int m = q == 0 ? mm : m_lo + m_hi - mm;
                    uint32_t ua[16], ub[16];
                    tmem_ld16(tl + 16 * m, ua);
                    tmem_ld16(tl + NP + 16 * m, ub);
                    tmem_ld_wait();
                    float u1[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) u1[i] = (__uint_as_float(ub[i]) + bf) * df_prescale<GATE>();  // off the serial chain
                    const bool full = 16 * m >= p_lo && 16 * m + 16 <= p_hi;
                    if (q == 0) {
                        if (full) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) cc = df_cstep<false, GATE>(cc, vf, u1[i], __uint_as_float(ua[i]), csj + (16 * m + i) * 64, true);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                cc = df_cstep<true, GATE>(cc, vf, u1[i], __uint_as_float(ua[i]), csj + (16 * m + i) * 64, 16 * m + i >= p_lo && 16 * m + i < p_hi);
                        }
                    } else {
                        if (full) {
#pragma unroll
                            for (int i = 15; i >= 0; --i) cc = df_cstep<false, GATE>(cc, vf, u1[i], __uint_as_float(ua[i]), csj + (16 * m + i) * 64, true);
                        } else {
#pragma unroll
                            for (int i = 15; i >= 0; --i)
                                cc = df_cstep<true, GATE>(cc, vf, u1[i], __uint_as_float(ua[i]), csj + (16 * m + i) * 64, 16 * m + i >= p_lo && 16 * m + i < p_hi);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(chunk_bar + (sw * 2 + q) * 16 + m);  // release: this block's c values are in shared memory
                }
            }
            DF_STAMP();
            // ---- h: warps q = 2 (columns 0-31) and q = 3 (columns 32-63); every step is independent, so each
            //      16-step batch is loaded, computed and stored as a block (no load waits behind a store)
            if (!DUAL && a.dbg != nullptr && tid == 64) a.dbg[(gridDim.x + blockIdx.x) * 16 + 2 * ly] = clock64();
            // When the tile holds half as many sequences as there are warp groups (time path: one or two long sequences), the groups
            // without a sequence of their own lend their h warps: they take every other 16-step block, so that h keeps up with the
            // c-recurrence (h costs ~1.6 k cycles per block and warp, c ~0.8 k: the h warps were the long pole of every layer).
            const bool h_help = 2 * a.nseq_tile == C::NG && sw >= a.nseq_tile;
            const int hs = h_help ? sw - a.nseq_tile : sw;
            if (q >= 2 && (seq_on || h_help)) {
                const int j = (q - 2) * 32 + lane;
                const float vr = __ldg(a.wc[ly] + 64 + j) * df_prescale<GATE>(), br = __ldg(a.bias[ly] + 64 + j);
                const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
                unsigned char* hb = hbuf + (j >> 2) * DF_LBO + 7 * 16 + (j & 3) * 4;
                const float* csj = cs + j;
                const int hp_lo = hs * S, hp_hi = hp_lo + L;
                const int p_end = ly == 3 ? hp_lo + S : hp_hi;  // the last layer also zeroes the 7 tail rows (conv-transpose padding)
                const int m_lo = hp_lo >> 4, m_hi = (p_end - 1) >> 4, m_chi = (hp_hi - 1) >> 4;
                const int par = 2 * a.nseq_tile == C::NG ? (h_help ? 1 : 0) : -1;  // which blocks (in scan order) this group takes; -1: all
                for (int mm = m_lo; mm <= m_hi; ++mm) {
                    if (par >= 0 && ((mm - m_lo) & 1) != par) continue;
                    const int m = q == 2 ? mm : m_lo + m_hi - mm;  // follow the c-recurrence's block order
                    if (m <= m_chi) mbar_wait(chunk_bar + (hs * 2 + (q - 2)) * 16 + m, ly & 1);
                    const bool full = 16 * m >= hp_lo && 16 * m + 16 <= hp_hi;
                    if (full) {
                        if (ly == 0) df_hchunk<true, true, GATE, NP>(tl, m, q == 3, hp_lo, hp_hi, p_end, vr, br, csj, hb);
                        else df_hchunk<true, false, GATE, NP>(tl, m, q == 3, hp_lo, hp_hi, p_end, vr, br, csj, hb);
                    } else {
                        if (ly == 0) df_hchunk<false, true, GATE, NP>(tl, m, q == 3, hp_lo, hp_hi, p_end, vr, br, csj, hb);
                        else df_hchunk<false, false, GATE, NP>(tl, m, q == 3, hp_lo, hp_hi, p_end, vr, br, csj, hb);
                    }
                }
            }
            if (!DUAL && a.dbg != nullptr && tid == 64) a.dbg[(gridDim.x + blockIdx.x) * 16 + 2 * ly + 1] = clock64();
            fence_proxy_async();
            tc_fence_before();
            DF_SYNC();
            tc_fence_after();
            DF_STAMP();
        }

        // ---- ConvTranspose1d: out[positions][64] = sum_kk slab[p + kk] . Wct_kk   (positions on the lanes)
        if (warp == 0) {
            if (elect_one()) {
                for (int gl = 0; gl < 8; ++gl) {  // one slab = the 16 K-pieces (4 channels each) of tap gl
                    int ss, owner;
                    uint32_t par;
                    df_slab_slot<NP>(cbeg + gl, it, ss, par, owner);
                    mbar_wait(full_w + ss, par);
                    tc_fence_after();
                    const uint64_t dw = d_wct + (uint64_t)((uint32_t)ss * SLAB16);
#pragma unroll
                    for (int k8 = 0; k8 < 8; ++k8) {
                        const uint64_t dh = d_slab + (uint64_t)(2u * (uint32_t)k8 * PIECE + (uint32_t)gl);
                        const uint32_t acc_on = (gl > 0 || k8 > 0) ? 1u : 0u;
                        umma_tf32(tmem, dh, dw + (uint64_t)(k8 * 128), IDESC_CT, acc_on);
                        if (NP == 256) umma_tf32(tmem + 64, dh + 128u, dw + (uint64_t)(k8 * 128), IDESC_CT, acc_on);
                    }
                    umma_commit(mma_done + ss);
                }
                umma_commit(acc_ready);
            }
            __syncwarp();
        } else if (is_prod) {
            produce_until(ibase + DF_NSLAB + C::NRING);  // runs ahead into the next tile's first layer-0 slabs (ring half only)
        }
        mbar_wait(acc_ready, nacc & 1);
        ++nacc;
        tc_fence_after();
        DF_STAMP();
        {
            const int mt = NP == 256 ? (warp >> 2) & 1 : 0, chalf = NP == 256 ? warp >> 3 : warp >> 2;
            // [32][36] floats per warp, ending at the end of the c buffer: the first rows overlap the tail of the slab, which the
            // transposed conv has finished reading (acc_ready) and the next tile rewrites after its first barrier
            float* stg = reinterpret_cast<float*>(sm + DF_HBUF + DF_CS) - (NW - warp) * (32 * TC_STG_LD);
            const int rsub = lane >> 3, c4 = (lane & 7) * 4;
            uint32_t vv[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * 64 + chalf * 32), vv);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4*>(stg + lane * TC_STG_LD + 4 * i) =
                    make_float4(__uint_as_float(vv[4 * i]), __uint_as_float(vv[4 * i + 1]), __uint_as_float(vv[4 * i + 2]), __uint_as_float(vv[4 * i + 3]));
            __syncwarp();
            const float* resid = a.first ? a.g_first : a.g_in;
            const float4 bi = ldg4(a.ct_bias + chalf * 32 + c4);
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
                const int r = pp * 4 + rsub;
                const int off = p2o[mt * 128 + q * 32 + r];
                if (off >= 0) {
                    const float4 x = *reinterpret_cast<const float4*>(stg + r * TC_STG_LD + c4);
                    const float4 rr = __ldcg(reinterpret_cast<const float4*>(resid + off + chalf * 32 + c4));
                    *reinterpret_cast<float4*>(a.g_out + off + chalf * 32 + c4) =
                        make_float4(x.x + bi.x + rr.x, x.y + bi.y + rr.y, x.z + bi.z + rr.z, x.w + bi.w + rr.w);
                }
            }
        }
        tc_fence_before();
        DF_STAMP();
    }
    __syncthreads();
#undef DF_STAMP
#undef DF_SYNC
    if (threadIdx.x < 32) tmem_dealloc<DUAL ? 512 : 2 * NP>(*tmem_slot);
}

template <int GATE, int NP, bool DUAL>
inline cudaError_t launch_dprnn_fused_g(DfArgs a, int tiles, cudaStream_t st) {
    static SmemCfg cfg;  // per instantiation, per device
    constexpr int smem = (DUAL ? 2 : 1) * DfC<NP>::SMEM + 64;
    if (cudaError_t e = ensure_smem(dprnn_fused_kernel<GATE, NP, DUAL>, smem, cfg); e != cudaSuccess) return e;
    const int grid = DUAL ? ((tiles + 1) / 2 < sm_count() ? (tiles + 1) / 2 : sm_count()) : tiles;
    dprnn_fused_kernel<GATE, NP, DUAL><<<grid, DUAL ? 512 : DfC<NP>::NT, smem, st>>>(a);
    return cudaGetLastError();
}
// Tile size: sequences of up to 128 steps run on 128-position tiles, two pipelines per persistent CTA; longer ones on 256-position
// tiles, one per CTA.  RTFS_DF_TILE=256 forces the latter (the round-1 kernel, kept as the A/B baseline), RTFS_DF_TILE=128 the
// 128-position tiles as independent CTAs (two resident per SM, no hand-off).
// The persistent kernel holds every SM for the whole launch (225 KB of shared memory per CTA), so a kernel forked onto a side stream
// would wait behind it: while one is in flight (yield_sms) the one-tile CTAs are used instead.
inline int dprnn_fused_mode(int S, bool yield_sms) {  // 0: 256-position tiles, 1: 128-position CTAs, 2: two 128-position pipelines per CTA
    static const int forced = [] {
        const char* v = getenv("RTFS_DF_TILE");
        return v ? atoi(v) : 0;
    }();
    if (forced == 256 || S > 128 || (yield_sms && forced == 0)) return 0;
    return forced == 128 ? 1 : 2;
}
inline int dprnn_fused_seq_per_tile(int S, bool yield_sms) {
    const int np = dprnn_fused_mode(S, yield_sms) == 0 ? 256 : 128, ng = np / 64;
    return np / S < ng ? np / S : ng;
}
inline int dprnn_fused_tiles(int S, int nseq_total, bool yield_sms) {
    const int per = dprnn_fused_seq_per_tile(S, yield_sms);
    return (nseq_total + per - 1) / per;
}
// rows of 16 clock64 stamps the kernel writes when DfArgs::dbg is set (x 2: the h-warp spans of the one-tile-per-CTA kernels)
inline int dprnn_fused_dbg_rows(int S, int nseq_total, bool yield_sms) {
    const int tiles = dprnn_fused_tiles(S, nseq_total, yield_sms);
    if (dprnn_fused_mode(S, yield_sms) != 2) return tiles;
    const int grid = (tiles + 1) / 2 < sm_count() ? (tiles + 1) / 2 : sm_count();
    return 2 * grid;
}
inline cudaError_t launch_dprnn_fused(DfArgs a, cudaStream_t st) {
    static const int gate = [] {
        const char* v = getenv("RTFS_DF_GATE");  // A/B of the gate arithmetic (see df_gate); tanh.approx measured parity-neutral
        return v ? atoi(v) : 1;
    }();
    const int mode = dprnn_fused_mode(a.S, a.yield_sms != 0);
    a.nseq_tile = dprnn_fused_seq_per_tile(a.S, a.yield_sms != 0);
    const int tiles = dprnn_fused_tiles(a.S, a.nseq_total, a.yield_sms != 0);
    if (mode == 2) {
        if (gate == 1) return launch_dprnn_fused_g<1, 128, true>(a, tiles, st);
        if (gate == 2) return launch_dprnn_fused_g<2, 128, true>(a, tiles, st);
        return launch_dprnn_fused_g<0, 128, true>(a, tiles, st);
    }
    if (mode == 1) return launch_dprnn_fused_g<1, 128, false>(a, tiles, st);
    if (gate == 1) return launch_dprnn_fused_g<1, 256, false>(a, tiles, st);
    if (gate == 2) return launch_dprnn_fused_g<2, 256, false>(a, tiles, st);
    return launch_dprnn_fused_g<0, 256, false>(a, tiles, st);
}

}  // namespace rtfs
