// tcgen05 (5th-gen tensor core) row-tile GEMM for the dense contractions of the RTFS-Net forward:
//     C[M x BN] = f(A)[M x KTOT] * W[BN x KTOT]^T        (fp32 in HBM, TF32 operands, fp32 accumulate in TMEM)
// One CTA = one 128-row tile.  K is streamed in chunks of 32 through an NS-stage shared-memory ring:
//   * A chunk: produced by the same loader functors as the legacy kernel (gemm.cuh: fused gLN-apply /
//     PReLU / gateway / TF-AR combine ...), rounded to TF32 (cvt.rna) and written by all 256 threads in
//     the UMMA canonical K-major no-swizzle layout (8-row x 16-byte core matrices):
//         byte(r, k) = (k/4) * LBO_A + r*16 + (k%4)*4 ,  LBO_A = 128*16 + 16 (pad: conflict-free STS.128), SBO = 128
//   * W chunk: a contiguous BN*128-byte slab of the host-prepared image Wimg[K/4][BN][4], fetched with one
//     cp.async.bulk (TMA engine, mbarrier complete_tx) -- LBO_B = BN*16, SBO = 128.
//   * thread 0 issues 4 x tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) per chunk and commits to the stage's
//     mbarrier; MMAs run asynchronously while the CTA stages the following chunks.
// Epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> per-warp shared-memory transpose -> the epilogue
// functor sees float4 pieces of rows, so every global access of the epilogue is a coalesced 128-byte row
// segment (bias / residual / gateway recompute / complex mask / gLN statistics fused as before).
#pragma once
#include <cstdlib>
#include "common.cuh"
#include "gemm.cuh"

namespace rtfs {

// ---------------------------------------------------------------------------------- PTX wrappers
DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Potentially blocking test of a phase.  The suspend-time hint (ns) lets the hardware park the thread until the phase completes
// or the time is up; without it the test returns almost at once and every waiting warp spins hot -- ncu on the mask + decoder GEMM
// counted 60 % of all executed warp instructions in such loops (16 producer warps polling next to the 8 epilogue warps that
// bound the kernel).  RTFS_MBAR_HINT_NS=0 at compile time restores the plain form for A/B runs.
#ifndef RTFS_MBAR_HINT_NS
#define RTFS_MBAR_HINT_NS 1000000
#endif
DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
#if RTFS_MBAR_HINT_NS > 0
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)RTFS_MBAR_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
#endif
    return ok != 0;
}
// Bounded wait: a protocol bug traps instead of hanging the device (2^14 tests of up to 1 ms each, or 2^24 plain ones).
DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (RTFS_MBAR_HINT_NS > 0 ? (1u << 14) : (1u << 24))) __trap();
    }
}
// TMA engine bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
DEVINL void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// True in exactly one lane of a CONVERGED warp (all 32 lanes must reach it).  tcgen05 / bulk-copy instructions take their operands
// from uniform registers: issued from an `if (tid == 0)` branch every UTCHMMA / UTCBAR is wrapped by the compiler in an ELECT +
// 5 x R2UR.BROADCAST + branch loop (measured 102 cycles of issue time per MMA and 120 per commit); behind elect.sync they are
// issued straight from uniform registers (62 cycles per MMA: tools/probe/dfissue_probe.cu, profiles/r02_df_probes.txt).
DEVINL bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// as elect_one, also returning the elected lane (the same value in every lane) so that state it kept can be broadcast afterwards
DEVINL bool elect_leader(uint32_t& leader) {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync %0|p, 0xffffffff;\n\tselp.u32 %1, 1, 0, p;\n\t}" : "=r"(leader), "=r"(pred));
    return pred != 0;
}
DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
DEVINL void tmem_alloc(uint32_t* slot) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"((uint32_t)NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
DEVINL void tmem_dealloc(uint32_t taddr) {  // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)NCOLS) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, TF32 inputs, fp32 accumulate; issued by ONE thread
DEVINL void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all tcgen05 ops issued so far by this thread have completed
DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (lane_base + i)
DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {  // no wait: pair with tcgen05.wait::ld
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor, version 1):
//   [0,14) start>>4 | [16,30) leading (K-direction core-matrix) byte offset>>4 | [32,46) stride (8-row group) byte offset>>4
DEVINL uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, TF32 x TF32, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// (sum, sumsq) of a thread group synchronised by a named barrier -> one fp64 atomic pair
DEVINL void group_stats_atomic(float s, float ss, double* dst, float* scratch, int gtid, int nthr, int barid) {
    s = warp_sum(s);
    ss = warp_sum(ss);
    const int lane = gtid & 31, w = gtid >> 5, nw = nthr >> 5;
    named_bar_sync(barid, nthr);
    if (lane == 0) {
        scratch[2 * w] = s;
        scratch[2 * w + 1] = ss;
    }
    named_bar_sync(barid, nthr);
    if (w == 0) {
        float a = lane < nw ? scratch[2 * lane] : 0.f;
        float b = lane < nw ? scratch[2 * lane + 1] : 0.f;
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0 && dst != nullptr) {
            atomicAdd(dst, (double)a);
            atomicAdd(dst + 1, (double)b);
        }
    }
}

// ---------------------------------------------------------------------------------- float4 epilogues
// contract: init(row0, M) ; pre = load(row, col) fetches what the row piece needs from global memory (all loads of a
//           batch of rows are issued before the first store, so they overlap instead of serialising behind the
//           stores) ; store4(row, col, v, pre) for columns col..col+3 of an in-range row ; finish(scratch)
struct NoPre {};
DEVINL float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

struct StoreEpi4 {
    float* C;
    long long ldc;
    const float* bias;  // may be null
    using Pre = NoPre;
    float4 bi_;
    DEVINL void init(int, int) {}
    DEVINL void prep(int col) { bi_ = bias ? ldg4(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f); }
    DEVINL Pre load(int, int) const { return Pre{}; }
    DEVINL void store4(int row, int col, float4 v, const Pre&) {
        v = add4(v, bi_);
        *reinterpret_cast<float4*>(C + (long long)row * ldc + col) = v;
    }
    DEVINL void finish(float*) {}
    DEVINL void finish_group(float*, int, int, int) {}
};

struct StatsEpi4 {
    float* C;
    long long ldc;
    const float* bias;
    double* sums;  // [B][2]
    int P, B;
    int bfirst_, split_;
    float s0_, q0_, s1_, q1_;
    DEVINL void init(int row0, int) {
        bfirst_ = row0 / P;
        split_ = (bfirst_ + 1) * P;
        s0_ = q0_ = s1_ = q1_ = 0.f;
    }
    using Pre = NoPre;
    float4 bi_;
    DEVINL void prep(int col) { bi_ = bias ? ldg4(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f); }
    DEVINL Pre load(int, int) const { return Pre{}; }
    DEVINL void store4(int row, int col, float4 v, const Pre&) {
        v = add4(v, bi_);
        *reinterpret_cast<float4*>(C + (long long)row * ldc + col) = v;
        const float s = (v.x + v.y) + (v.z + v.w);
        const float q = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        if (row < split_) {
            s0_ += s;
            q0_ += q;
        } else {
            s1_ += s;
            q1_ += q;
        }
    }
    DEVINL void finish(float* scratch) {
        block_stats_atomic(s0_, q0_, sums + 2 * bfirst_, scratch);
        block_stats_atomic(s1_, q1_, (bfirst_ + 1 < B) ? sums + 2 * (bfirst_ + 1) : nullptr, scratch);
    }
    DEVINL void finish_group(float* scratch, int gtid, int nthr, int barid) {
        group_stats_atomic(s0_, q0_, sums + 2 * bfirst_, scratch, gtid, nthr, barid);
        group_stats_atomic(s1_, q1_, (bfirst_ + 1 < B) ? sums + 2 * (bfirst_ + 1) : nullptr, scratch, gtid, nthr, barid);
    }
};

// residual_conv epilogue (tdanet.py:131): out = acc + bias + gateway(x) [+ a1 -> next block input]
struct ResidOutEpi4 {
    float* out;
    const float* bias;
    const float* x;
    const float* wg;
    const float* bg;
    const float* slope;
    const float* a1;  // may be null
    float a_;
    struct Pre {
        float4 x, a1;
    };
    float4 w_, b_, bi_;  // per-column constants of the current 32-column block (prep)
    DEVINL void init(int, int) { a_ = __ldg(slope); }
    DEVINL void prep(int col) {
        w_ = ldg4(wg + col);
        b_ = ldg4(bg + col);
        bi_ = ldg4(bias + col);
    }
    DEVINL Pre load(int row, int col) const {
        const long long o = (long long)row * 256 + col;
        Pre p;
        p.x = ldg4(x + o);
        p.a1 = a1 ? ldg4(a1 + o) : make_float4(0.f, 0.f, 0.f, 0.f);
        return p;
    }
    DEVINL void store4(int row, int col, float4 v, const Pre& p) {
        const long long o = (long long)row * 256 + col;
        const float4 w = w_, b = b_, bi = bi_;
        v.x += bi.x + prelu(fmaf(w.x, p.x.x, b.x), a_) + p.a1.x;
        v.y += bi.y + prelu(fmaf(w.y, p.x.y, b.y), a_) + p.a1.y;
        v.z += bi.z + prelu(fmaf(w.z, p.x.z, b.z), a_) + p.a1.z;
        v.w += bi.w + prelu(fmaf(w.w, p.x.w, b.w), a_) + p.a1.w;
        *reinterpret_cast<float4*>(out + o) = v;
    }
    // affine variant used by gemm_tc_kernel: the eight rows a thread touches in a 32x32 block are rowq0 + rsub + 4p, so
    // every access is one of three block pointers plus a compile-time offset (no per-access 64-bit address arithmetic)
    static constexpr bool kAffine = true;
    const float *xp_, *a1p_;
    float* op_;
    DEVINL void prep_block(int rowq0, int rsub, int col, int) {
        prep(col);
        const long long o = (long long)(rowq0 + rsub) * 256 + col;
        xp_ = x + o;
        a1p_ = a1 ? a1 + o : nullptr;
        op_ = out + o;
    }
    DEVINL Pre load_p(int p) const {
        Pre r;
        r.x = ldg4(xp_ + p * 1024);
        r.a1 = a1p_ ? ldg4(a1p_ + p * 1024) : make_float4(0.f, 0.f, 0.f, 0.f);
        return r;
    }
    DEVINL void store_p(int p, int, float4 v, const Pre& r) {
        const float4 w = w_, b = b_, bi = bi_;
        v.x += bi.x + prelu(fmaf(w.x, r.x.x, b.x), a_) + r.a1.x;
        v.y += bi.y + prelu(fmaf(w.y, r.x.y, b.y), a_) + r.a1.y;
        v.z += bi.z + prelu(fmaf(w.z, r.x.z, b.z), a_) + r.a1.z;
        v.w += bi.w + prelu(fmaf(w.w, r.x.w, b.w), a_) + r.a1.w;
        *reinterpret_cast<float4*>(op_ + p * 1024) = v;
    }
    DEVINL void finish(float*) {}
    DEVINL void finish_group(float*, int, int, int) {}
};

// residual_conv epilogue of the FIRST block pass with the CAF fusion (layers/fusion.py:252-274) applied to the block
// output in registers: y = ReLU(v*sk+tk) * vk[b][near(t)][c] + att[b][near(t)][c] * (v*sv+tv) [+ a1], v = block output.
// Saves the separate streaming pass over the (B,T,F,256) tensor (one read + one write).
struct ResidOutCafEpi4 {
    float* out;
    const float* bias;
    const float* x;
    const float* wg;
    const float* bg;
    const float* slope;
    const float* a1;  // may be null
    const float* vk;  // (B,Tv,256)
    const float* att;
    const float* sk;
    const float* tk;
    const float* sv;
    const float* tv;
    int T, F, Tv;
    float a_;
    struct Pre {
        float4 x, a1;
    };
    float4 w_, b_, bi_, s1_, t1_, s2_, t2_;  // per-column constants of the current 32-column block
    DEVINL void init(int, int) { a_ = __ldg(slope); }
    DEVINL void prep(int col) {
        w_ = ldg4(wg + col);
        b_ = ldg4(bg + col);
        bi_ = ldg4(bias + col);
        s1_ = ldg4(sk + col);
        t1_ = ldg4(tk + col);
        s2_ = ldg4(sv + col);
        t2_ = ldg4(tv + col);
    }
    DEVINL Pre load(int row, int col) const {
        const long long o = (long long)row * 256 + col;
        Pre p;
        p.x = ldg4(x + o);
        p.a1 = a1 ? ldg4(a1 + o) : make_float4(0.f, 0.f, 0.f, 0.f);
        return p;
    }
    DEVINL void store4(int row, int col, float4 v, const Pre& p) {
        const long long o = (long long)row * 256 + col;
        const float4 w = w_, b = b_, bi = bi_;
        v.x += bi.x + prelu(fmaf(w.x, p.x.x, b.x), a_);
        v.y += bi.y + prelu(fmaf(w.y, p.x.y, b.y), a_);
        v.z += bi.z + prelu(fmaf(w.z, p.x.z, b.z), a_);
        v.w += bi.w + prelu(fmaf(w.w, p.x.w, b.w), a_);
        const int bt = row / F, t = bt % T, bb = bt / T;
        const int tvi = (t * Tv) / T;  // nearest: floor(t * Tv / T) (< Tv)
        const long long vo = ((long long)bb * Tv + tvi) * 256 + col;
        const float4 k = ldg4(vk + vo), at = ldg4(att + vo);
        const float4 s1 = s1_, t1 = t1_, s2 = s2_, t2 = t2_;
        float4 y;
        y.x = fmaxf(fmaf(v.x, s1.x, t1.x), 0.f) * k.x + at.x * fmaf(v.x, s2.x, t2.x) + p.a1.x;
        y.y = fmaxf(fmaf(v.y, s1.y, t1.y), 0.f) * k.y + at.y * fmaf(v.y, s2.y, t2.y) + p.a1.y;
        y.z = fmaxf(fmaf(v.z, s1.z, t1.z), 0.f) * k.z + at.z * fmaf(v.z, s2.z, t2.z) + p.a1.z;
        y.w = fmaxf(fmaf(v.w, s1.w, t1.w), 0.f) * k.w + at.w * fmaf(v.w, s2.w, t2.w) + p.a1.w;
        *reinterpret_cast<float4*>(out + o) = y;
    }
    // affine variant (see ResidOutEpi4).  The 32 rows of a warp's block are consecutive positions, so with F >= 32 they
    // lie in at most two (b, t) frames: the video key / attention vectors of both frames are loaded once per block
    // instead of once per row behind the previous row's store (host side falls back to the unfused CAF when F < 32).
    static constexpr bool kAffine = true;
    // In the fused pass the addend, when present, IS the block input (a1 + block(a1)): one load serves both.
    struct PreA {
        float4 x;
    };
    const float* xp_;
    float* op_;
    bool addx_;
    float4 k0_, at0_, k1_, at1_;
    int rsplit_;  // first row of the second frame
    DEVINL long long vrow(int bt) const {
        const int t = bt % T, bb = bt / T;
        return ((long long)bb * Tv + (t * Tv) / T) * 256;  // nearest: floor(t * Tv / T) (< Tv)
    }
    DEVINL void prep_block(int rowq0, int rsub, int col, int M) {
        prep(col);
        const long long o = (long long)(rowq0 + rsub) * 256 + col;
        xp_ = x + o;
        addx_ = a1 != nullptr;
        op_ = out + o;
        const int bt0 = rowq0 / F;
        rsplit_ = (bt0 + 1) * F;
        const long long v0 = vrow(bt0) + col, v1 = (rsplit_ < M ? vrow(bt0 + 1) : vrow(bt0)) + col;
        k0_ = ldg4(vk + v0);
        at0_ = ldg4(att + v0);
        k1_ = ldg4(vk + v1);
        at1_ = ldg4(att + v1);
    }
    DEVINL PreA load_p(int p) const {
        PreA r;
        r.x = ldg4(xp_ + p * 1024);
        return r;
    }
    DEVINL void store_p(int p, int row, float4 v, const PreA& r) {
        const float4 w = w_, b = b_, bi = bi_;
        v.x += bi.x + prelu(fmaf(w.x, r.x.x, b.x), a_);
        v.y += bi.y + prelu(fmaf(w.y, r.x.y, b.y), a_);
        v.z += bi.z + prelu(fmaf(w.z, r.x.z, b.z), a_);
        v.w += bi.w + prelu(fmaf(w.w, r.x.w, b.w), a_);
        const bool second = row >= rsplit_;
        const float4 k = second ? k1_ : k0_, at = second ? at1_ : at0_;
        const float4 s1 = s1_, t1 = t1_, s2 = s2_, t2 = t2_;
        float4 y;
        y.x = fmaxf(fmaf(v.x, s1.x, t1.x), 0.f) * k.x + at.x * fmaf(v.x, s2.x, t2.x) + (addx_ ? r.x.x : 0.f);
        y.y = fmaxf(fmaf(v.y, s1.y, t1.y), 0.f) * k.y + at.y * fmaf(v.y, s2.y, t2.y) + (addx_ ? r.x.y : 0.f);
        y.z = fmaxf(fmaf(v.z, s1.z, t1.z), 0.f) * k.z + at.z * fmaf(v.z, s2.z, t2.z) + (addx_ ? r.x.z : 0.f);
        y.w = fmaxf(fmaf(v.w, s1.w, t1.w), 0.f) * k.w + at.w * fmaf(v.w, s2.w, t2.w) + (addx_ ? r.x.w : 0.f);
        *reinterpret_cast<float4*>(op_ + p * 1024) = y;
    }
    DEVINL void finish(float*) {}
    DEVINL void finish_group(float*, int, int, int) {}
};

// S^3 mask epilogue (mask_generator.py:67-99); GEMM columns interleaved on the host: col 2c = real-half
// channel c, col 2c+1 = imag-half channel c+128 -> a float4 holds (re c, im c, re c+1, im c+1).
template <bool SAVE_M>
struct MaskEpi4T {
    float* z;
    const float* bias;
    const float* a0;
    float* m_out;  // SAVE_M (training tape): the mask m = ReLU(conv) in natural channel order (0..127 real | 128..255 imag)
    static constexpr int kTcpEpiRegs = 104;  // persistent kernel: register re-allocation towards the epilogue warps (setmaxnreg)
    static constexpr bool kRollPre = false;  // rolling prefetch of the next block (gemm_tcp.cuh ROLL): measured slower (0.96 vs 0.80 ms; 1.10 vs 0.77 ms with the register split)
    struct Pre {
        float2 er, ei;
    };
    float4 bi_;
    DEVINL void init(int, int) {}
    DEVINL void prep(int col) { bi_ = ldg4(bias + col); }
    DEVINL Pre load(int row, int col) const {
        const long long o = (long long)row * 256 + (col >> 1);
        Pre p;
        p.er = ldg2(a0 + o);
        p.ei = ldg2(a0 + o + 128);
        return p;
    }
    DEVINL void store4(int row, int col, float4 v, const Pre& p) {
        const float4 bi = bi_;
        const float mr0 = fmaxf(v.x + bi.x, 0.f), mi0 = fmaxf(v.y + bi.y, 0.f);
        const float mr1 = fmaxf(v.z + bi.z, 0.f), mi1 = fmaxf(v.w + bi.w, 0.f);
        const long long o = (long long)row * 256 + (col >> 1);
        const float2 er = p.er, ei = p.ei;
        if (SAVE_M) {
            *reinterpret_cast<float2*>(m_out + o) = make_float2(mr0, mr1);
            *reinterpret_cast<float2*>(m_out + o + 128) = make_float2(mi0, mi1);
        }
        *reinterpret_cast<float2*>(z + o) = make_float2(er.x * mr0 - ei.x * mi0, er.y * mr1 - ei.y * mi1);
        *reinterpret_cast<float2*>(z + o + 128) = make_float2(er.x * mi0 + ei.x * mr0, er.y * mi1 + ei.y * mr1);
    }
    DEVINL void finish(float*) {}
    DEVINL void finish_group(float*, int, int, int) {}
};
using MaskEpi4 = MaskEpi4T<false>;

constexpr int TC_STG_LD = 36;  // floats per staged epilogue row (32 + 4 pad)

// S^3 mask + decoder contraction in one epilogue (mask_generator.py:67-99 + decoder.py:110-116): the masked embedding z is never
// written.  Every 32 x 32 block of z goes back into the warp's staging tile; then thread = row accumulates the 18 partial
// products of the transposed 3x3 conv (ConvTranspose2d 256 -> 2: o*9 + i*3 + j) for its row in fp32 against a shared-memory
// copy of the decoder filter ([interleaved column][20]: all lanes read the same entry -> broadcast); the two column halves of a
// row are combined through the staging tile and Q18 (7 % of a 256-channel tensor) is the only output.
struct MaskDecEpi4 {
    float* q18;          // [M][18]
    const float* bias;   // [256] interleaved like the GEMM columns
    const float* a0;     // encoder output [M][256]
    const float* wdec;   // [256][20] decoder filter, rows in interleaved column order, 18 taps + 2 zeros (weights.py: RTFS_P_DEC_WT)
    static constexpr int kTcpEpiRegs = 104;
    static constexpr bool kRollPre = false;
    static constexpr bool kFusedRows = true;
    static constexpr bool kPrefetchTile = true;
    static constexpr int kSmemBytes = 256 * 20 * 4;
    struct Pre {
        float2 er, ei;
    };
    float4 bi_;
    const float* wtab_;
    float2 q_[2][9];  // the 18 partial products of this lane's two rows, packed for f32x2 FMAs
    DEVINL void bind_smem(float* tab, int tid, int nthr) {
        for (int i = tid; i < 256 * 5; i += nthr) reinterpret_cast<float4*>(tab)[i] = ldg4(wdec + 4 * i);
        wtab_ = tab;
    }
    DEVINL void init(int, int) {
#pragma unroll
        for (int r = 0; r < 9; ++r) q_[0][r] = q_[1][r] = make_float2(0.f, 0.f);
    }
    DEVINL void prep(int col) { bi_ = ldg4(bias + col); }
    DEVINL Pre load(int row, int col) const {
        const long long o = (long long)row * 256 + (col >> 1);
        Pre p;
        p.er = ldg2(a0 + o);
        p.ei = ldg2(a0 + o + 128);
        return p;
    }
    // z piece of (row, interleaved columns col..col+3) = (re c, im c, re c+1, im c+1) -> the staging slot it came from
    DEVINL void store4s(int, int, float4 v, const Pre& p, float* slot) {
        const float4 bi = bi_;
        const float mr0 = fmaxf(v.x + bi.x, 0.f), mi0 = fmaxf(v.y + bi.y, 0.f);
        const float mr1 = fmaxf(v.z + bi.z, 0.f), mi1 = fmaxf(v.w + bi.w, 0.f);
        const float2 er = p.er, ei = p.ei;
        *reinterpret_cast<float4*>(slot) = make_float4(er.x * mr0 - ei.x * mi0, er.x * mi0 + ei.x * mr0, er.y * mr1 - ei.y * mi1, er.y * mi1 + ei.y * mr1);
    }
    // L2 prefetch of the a0 rows of a tile (128 rows x 1 KB = 1024 lines), issued by the 256 epilogue threads one tile ahead: the
    // epilogue's a0 loads are exposed once per 32-column block, and an L2 hit costs a third of a DRAM round trip under load
    DEVINL void prefetch_tile(int row0, int M, int tid) const {
        const char* base = reinterpret_cast<const char*>(a0 + (long long)row0 * 256);
        const long long lim = ((long long)M - row0) * 1024;  // bytes of a0 from row0 to the end
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long off = (long long)(tid + 256 * i) * 128;
            if (off < lim) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
        }
    }
    // The decoder taps of the 32 staged z columns of this block.  Lane = (row pair rp = lane & 15, column parity cp = lane >> 4):
    // rows 2 rp, 2 rp + 1 of the warp's 32 against the filter rows of columns col0 + cp, col0 + 2 + cp, ...: every filter row read
    // from shared memory feeds 18 packed FMAs instead of 9 (the kernel was bound by these loads: ncu `mio` stalls on LDS, 5 loads per
    // 9 FMAs with one row per lane), and the two parities read rows 20 floats apart, i.e. disjoint banks.
    DEVINL void block_reduce(const float* stg, int lane, int col0) {
        const int rp = lane & 15, cp = lane >> 4;
        const float* z0 = stg + (2 * rp) * TC_STG_LD;
        const float4* w = reinterpret_cast<const float4*>(wtab_ + (col0 + cp) * 20);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 a = *reinterpret_cast<const float4*>(z0 + 4 * j4);
            const float4 b = *reinterpret_cast<const float4*>(z0 + TC_STG_LD + 4 * j4);
            const float za[2] = {cp ? a.y : a.x, cp ? a.w : a.z};
            const float zb[2] = {cp ? b.y : b.x, cp ? b.w : b.z};
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const float4* wr = w + (4 * j4 + 2 * jj) * 5;
                const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
                const float2 w4 = *reinterpret_cast<const float2*>(wr + 4);
                const float2 wv[9] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w), make_float2(w2.x, w2.y),
                                      make_float2(w2.z, w2.w), make_float2(w3.x, w3.y), make_float2(w3.z, w3.w), w4};
                const float2 ya = make_float2(za[jj], za[jj]), yb = make_float2(zb[jj], zb[jj]);
#pragma unroll
                for (int r = 0; r < 9; ++r) {
                    q_[0][r] = __ffma2_rn(ya, wv[r], q_[0][r]);
                    q_[1][r] = __ffma2_rn(yb, wv[r], q_[1][r]);
                }
            }
        }
    }
    // fold the two column parities of a row pair (lanes l, l ^ 16), then the two column halves of the tile (warps q and q + 4, through
    // the upper half's staging tile) and store Q18; row0w = first row of the warp's 32
    DEVINL void tile_done(float* stg_all, int warp, int lane, int row0w, int M) {
        const int q = warp & 3, hlf = warp >> 2;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int r = 0; r < 9; ++r) {
                q_[h][r].x += __shfl_xor_sync(0xffffffffu, q_[h][r].x, 16);
                q_[h][r].y += __shfl_xor_sync(0xffffffffu, q_[h][r].y, 16);
            }
        if (hlf == 1 && lane < 16) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float* mine = stg_all + warp * (32 * TC_STG_LD) + (2 * lane + h) * TC_STG_LD;
#pragma unroll
                for (int r = 0; r < 9; ++r) *reinterpret_cast<float2*>(mine + 2 * r) = q_[h][r];
            }
        }
        named_bar_sync(4 + q, 64);
        if (hlf == 0 && lane < 16) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = row0w + 2 * lane + h;
                const float* other = stg_all + (warp + 4) * (32 * TC_STG_LD) + (2 * lane + h) * TC_STG_LD;
                if (row < M) {
                    float* dst = q18 + (long long)row * 18;
#pragma unroll
                    for (int r = 0; r < 9; ++r) {
                        const float2 o = *reinterpret_cast<const float2*>(other + 2 * r);
                        *reinterpret_cast<float2*>(dst + 2 * r) = make_float2(q_[h][r].x + o.x, q_[h][r].y + o.y);
                    }
                }
            }
        }
        named_bar_sync(4 + q, 64);  // the upper half may reuse its staging tile
    }
    DEVINL void finish(float*) {}
    DEVINL void finish_group(float*, int, int, int) {}
};

// ConvTranspose1d-as-GEMM epilogue of the dual-path RNN (rnn_layers.py:153-160)
struct ConvTEpi4 {
    float* out;
    const float* resid;
    const float* bias;
    int S, n_other, time_path, Tc, Fc;
    struct Pre {
        float4 r;
        long long off;  // < 0: padding row, nothing to store
    };
    DEVINL void init(int, int) {}
    DEVINL void prep(int) {}
    DEVINL Pre load(int row, int col) const {
        Pre p;
        const int seq = row / (S + 7), s = row - seq * (S + 7);
        p.off = -1;
        p.r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < S) {
            const int b = seq / n_other, o = seq - b * n_other;
            const int t = time_path ? s : o, f = time_path ? o : s;
            p.off = ((((long long)b * Tc + t) * Fc) + f) * 64 + col;
            p.r = ldg4(resid + p.off);
        }
        return p;
    }
    DEVINL void store4(int, int col, float4 v, const Pre& p) {
        if (p.off < 0) return;
        v = add4(add4(v, ldg4(bias + col)), p.r);
        *reinterpret_cast<float4*>(out + p.off) = v;
    }
    DEVINL void finish(float*) {}
    DEVINL void finish_group(float*, int, int, int) {}
};

// ---------------------------------------------------------------------------------- kernel
constexpr int TC_BM = 128, TC_KC = 32, TC_THREADS = 256;
constexpr int TC_LBO_A = TC_BM * 16 + 16;              // 2064 B between K-direction core matrices of A
constexpr int TC_A_STAGE = 16640;                      // 8 * 2064 = 16512, rounded up to 128
constexpr int TC_STG_BYTES = 8 * 32 * TC_STG_LD * 4;   // 8 warps x [32][36] floats

template <int BN>
__host__ __device__ constexpr int tc_tmem_cols() {
    return BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
}
template <int BN, int NS>
__host__ __device__ constexpr int tc_stage_bytes() {
    return NS * (TC_A_STAGE + BN * 128);
}
template <int BN, int NS, int NT = 256>
__host__ __device__ constexpr int tc_smem_bytes(int extra_floats) {
    return (tc_stage_bytes<BN, NS>() > (NT / 32) * 32 * TC_STG_LD * 4 ? tc_stage_bytes<BN, NS>() : (NT / 32) * 32 * TC_STG_LD * 4) + extra_floats * 4 + 256 + 128;
}

// NT = 256 or 512 threads: HBM bandwidth on this part scales with the number of warps that issue loads
// (tools/probe/inflight_probe2.cu: 8 warps/SM ~4 TB/s, 16 ~5.9, 32 ~6.1), so the stream-heavy instances run 2 x 512.
template <class EP, class = void>
struct ep_affine {
    static constexpr bool value = false;
};
template <class EP>
struct ep_affine<EP, decltype((void)EP::kAffine)> {
    static constexpr bool value = EP::kAffine;
};

template <class EP, class = void>
struct ep_pre {
    using type = typename EP::Pre;
};
template <class EP>
struct ep_pre<EP, decltype((void)sizeof(typename EP::PreA))> {
    using type = typename EP::PreA;  // leaner per-row state of the affine interface
};

template <class EP, class = void>
struct ep_epi_regs {
    static constexpr int value = 0;
};
template <class EP>
struct ep_epi_regs<EP, decltype((void)EP::kTcpEpiRegs)> {
    static constexpr int value = EP::kTcpEpiRegs;
};

template <class EP, class = void>
struct ep_roll {
    static constexpr bool value = false;
};
template <class EP>
struct ep_roll<EP, decltype((void)EP::kRollPre)> {
    static constexpr bool value = EP::kRollPre;
};

// epilogues that consume whole rows through the staging tile (fused S^3 mask + decoder) and own a shared-memory table
template <class EP, class = void>
struct ep_prefetch {
    static constexpr bool value = false;
};
template <class EP>
struct ep_prefetch<EP, decltype((void)EP::kPrefetchTile)> {
    static constexpr bool value = EP::kPrefetchTile;
};
template <class EP, class = void>
struct ep_fused_rows {
    static constexpr bool value = false;
};
template <class EP>
struct ep_fused_rows<EP, decltype((void)EP::kFusedRows)> {
    static constexpr bool value = EP::kFusedRows;
};
template <class EP, class = void>
struct ep_smem_bytes {
    static constexpr int value = 0;
};
template <class EP>
struct ep_smem_bytes<EP, decltype((void)EP::kSmemBytes)> {
    static constexpr int value = EP::kSmemBytes;
};

template <class AL, class = void>
struct loader_batched {
    static constexpr bool value = false;
};
template <class AL>
struct loader_batched<AL, decltype((void)AL::kBatched)> {
    static constexpr bool value = AL::kBatched;
};

template <int BN, int KTOT, int NS, int MINB, int PF, int NT, class AL, class EP>
__global__ void __launch_bounds__(NT, MINB) gemm_tc_kernel(AL al, const float* __restrict__ Wimg, EP ep, int M) {
    constexpr int NK = KTOT / TC_KC;
    static_assert(PF >= 1 && PF <= 4, "A-operand register prefetch depth (chunks in flight per thread)");
    constexpr int WBYTES = BN * 128;
    constexpr int TCOLS = tc_tmem_cols<BN>();
    constexpr int PD = (NK <= NS) ? NK : NS - 2;  // W chunks in flight ahead of the A producer
    static_assert(KTOT % TC_KC == 0, "K must be a multiple of 32");
    static_assert(BN % 64 == 0 && BN <= 256, "N must be 64, 128, 192 or 256");
    static_assert(NK <= NS || NS >= 3, "a ring shorter than K needs >= 3 stages");
    constexpr int RPT = 1024 / NT;          // A rows per thread per chunk (4 or 2)
    constexpr int RS = NT / 8;              // row stride between them
    constexpr int STG_BYTES = (NT / 32) * 32 * TC_STG_LD * 4;
    static_assert(NT == 256 || NT == 512, "256 or 512 threads");
    constexpr int STAGE_AREA = tc_stage_bytes<BN, NS>() > STG_BYTES ? tc_stage_bytes<BN, NS>() : STG_BYTES;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* a_stage = smem_raw;
    unsigned char* w_stage = smem_raw + NS * TC_A_STAGE;
    float* extra = reinterpret_cast<float*>(smem_raw + STAGE_AREA);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + STAGE_AREA + AL::kExtra * 4);  // full_w[NS] | mma_done[NS]
    uint64_t* full_w = bars;
    uint64_t* mma_done = bars + NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS);
    float* scratch = reinterpret_cast<float*>(tmem_slot + 4);  // 16 floats for block_stats_atomic

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.x * TC_BM;

    if (warp == 0) tmem_alloc<TCOLS>(tmem_slot);
    if (tid == 32) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(full_w + s, 1);
            mbar_init(mma_done + s, 1);
        }
        fence_mbar_init();
    }
    al.init(row0, M, extra);
    ep.init(row0, M);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    auto issue_w = [&](int c) {  // thread 0 only
        const int s = c % NS;
        mbar_expect_tx(full_w + s, WBYTES);
        bulk_g2s(w_stage + (size_t)s * WBYTES, Wimg + (size_t)c * (BN * TC_KC), WBYTES, full_w + s);
    };
    if (warp == 0 && elect_one()) {
#pragma unroll
        for (int c = 0; c < PD && c < NK; ++c) issue_w(c);
    }

    // thread (tid>>3)+32i owns row r_i of the tile, 16-byte K piece (tid&7) of every chunk
    const int kq = tid & 7;
    unsigned char* a_dst0 = a_stage + kq * TC_LBO_A + (tid >> 3) * 16;
    constexpr uint32_t IDESC = umma_idesc_tf32(TC_BM, BN);

    auto issue_mma = [&](int kc, int s) {  // thread 0, after the CTA barrier that follows the chunk's stores
        mbar_wait(full_w + s, (kc / NS) & 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(a_stage + (size_t)s * TC_A_STAGE);
        const uint32_t w_base = smem_u32(w_stage + (size_t)s * WBYTES);
#pragma unroll
        for (int q = 0; q < TC_KC / 8; ++q) {
            const uint64_t da = umma_desc(a_base + 2 * q * TC_LBO_A, TC_LBO_A, 128);
            const uint64_t db = umma_desc(w_base + 2 * q * (BN * 16), BN * 16, 128);
            umma_tf32(tmem, da, db, IDESC, (kc > 0 || q > 0) ? 1u : 0u);
        }
        umma_commit(mma_done + s);
    };

    if constexpr (loader_batched<AL>::value) {
        // loaders with several operand loads per element (TF-AR combine: 4): the raw loads of a whole chunk are issued
        // as ONE batch and transformed afterwards -- otherwise the compiler serialises load -> transform per row and the
        // A phase costs RPT * NK memory round trips instead of NK (measured: 31 % of the residual conv's warp time)
        typename AL::Raw raw[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) raw[i] = al.raw_load(i, kq * 4);
#pragma unroll
        for (int kc = 0; kc < NK; ++kc) {
            const int s = kc % NS;
            if (kc >= NS) mbar_wait(mma_done + s, ((kc / NS) - 1) & 1);
            if (NK > NS && kc + PD < NK && warp == 0 && elect_one()) {
                const int c = kc + PD;
                if (c >= NS) mbar_wait(mma_done + (c % NS), ((c / NS) - 1) & 1);
                issue_w(c);
            }
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                float4 v = al.xform_raw(raw[i], i, kc * TC_KC + kq * 4);
                v.x = tf32r(v.x);
                v.y = tf32r(v.y);
                v.z = tf32r(v.z);
                v.w = tf32r(v.w);
                *reinterpret_cast<float4*>(a_dst0 + (size_t)s * TC_A_STAGE + i * (RS * 16)) = v;
            }
            fence_proxy_async();
            if (kc + 1 < NK) {
#pragma unroll
                for (int i = 0; i < RPT; ++i) raw[i] = al.raw_load(i, (kc + 1) * TC_KC + kq * 4);
            }
            __syncthreads();
            if (warp == 0 && elect_one()) issue_mma(kc, s);
        }
    } else {
    float4 areg[PF][RPT];
#pragma unroll
    for (int c = 0; c < PF && c < NK; ++c) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) areg[c][i] = al.load(i, c * TC_KC + kq * 4);
    }

#pragma unroll
    for (int kc = 0; kc < NK; ++kc) {
        const int s = kc % NS;
        if (kc >= NS) mbar_wait(mma_done + s, ((kc / NS) - 1) & 1);  // MMAs of chunk kc-NS have read this stage
        if (NK > NS && kc + PD < NK && warp == 0 && elect_one()) {
            const int c = kc + PD;  // its stage was last read by chunk c-NS = kc-2
            if (c >= NS) mbar_wait(mma_done + (c % NS), ((c / NS) - 1) & 1);
            issue_w(c);
        }
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            float4 v = areg[kc % PF][i];
            v.x = tf32r(v.x);
            v.y = tf32r(v.y);
            v.z = tf32r(v.z);
            v.w = tf32r(v.w);
            *reinterpret_cast<float4*>(a_dst0 + (size_t)s * TC_A_STAGE + i * (RS * 16)) = v;
        }
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy (issued BEFORE the next
                              // chunk's loads so that it does not wait behind them)
        if (kc + PF < NK) {
#pragma unroll
            for (int i = 0; i < RPT; ++i) areg[kc % PF][i] = al.load(i, (kc + PF) * TC_KC + kq * 4);
        }
        __syncthreads();
        if (warp == 0 && elect_one()) issue_mma(kc, s);
    }
    }
    // accumulator complete once the last chunk's commit has arrived (commits complete in order)
    mbar_wait(mma_done + ((NK - 1) % NS), ((NK - 1) / NS) & 1);
    tc_fence_after();

    // ---- epilogue: warp w reads TMEM lanes 32*(w&3).. (rows), column half (w>>2)
    {
        constexpr int NCG = (NT / 128) < (BN / 32) ? (NT / 128) : (BN / 32);  // column groups sharing a lane quarter
        const int q = warp & 3, hlf = warp >> 2;
        float* stg = reinterpret_cast<float*>(smem_raw) + warp * (32 * TC_STG_LD);
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll 1
        for (int cb = 0; cb < (hlf < NCG ? BN / (32 * NCG) : 0); ++cb) {
            const int col0 = hlf * (BN / NCG) + cb * 32;
            // all global loads of this 32x32 block are issued first and stay in flight while the
            // accumulator block is read from TMEM and transposed through shared memory
            constexpr bool AFF = ep_affine<EP>::value;
            const int rowq0 = row0 + q * 32;
            if constexpr (AFF) ep.prep_block(rowq0, rsub, col0 + c4, M);
            else ep.prep(col0 + c4);
            typename ep_pre<EP>::type pre[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int row = rowq0 + rsub + p * 4;
                if constexpr (AFF) {
                    if (row < M) pre[p] = ep.load_p(p);
                } else {
                    pre[p] = ep.load(row < M ? row : M - 1, col0 + c4);
                }
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t v[16];
                tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col0 + 16 * hh, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4*>(stg + lane * TC_STG_LD + 16 * hh + 4 * i) =
                        make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
            __syncwarp();
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int r = p * 4 + rsub;
                const float4 x = *reinterpret_cast<const float4*>(stg + r * TC_STG_LD + c4);
                const int row = rowq0 + r;
                if (row < M) {
                    if constexpr (AFF) ep.store_p(p, row, x, pre[p]);
                    else ep.store4(row, col0 + c4, x, pre[p]);
                }
            }
            __syncwarp();
        }
    }
    ep.finish(scratch);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TCOLS>(tmem);
}

// ---------------------------------------------------------------------------------- overlapping-view variant
// C[M x BN] = A_view[M x 8*64] * W^T where A_view[r][tap*64+c] = X[(r+tap)*64 + c] (nn.Unfold(8) o Linear of
// the dual-path RNN, and ConvTranspose1d(k=8) over the zero-padded h).  The (128+7) x 64 slab of X is staged
// ONCE in the K-major no-swizzle layout; tap t is then just the same slab with the descriptor start address
// advanced by t rows (t*16 bytes) -- no im2col copy, 8x less shared-memory fill than the generic kernel.
constexpr int TCU_ROWS = TC_BM + 8;                 // 136 rows staged (135 needed)
constexpr int TCU_LBO = TCU_ROWS * 16 + 16;         // 2192
constexpr int TCU_A_BYTES = 16 * TCU_LBO;           // 16 K-pieces of 4 channels = 35072

template <int BN, int NS>
__host__ __device__ constexpr int tcu_smem_bytes() {
    return (TCU_A_BYTES + NS * BN * 128 > TC_STG_BYTES ? TCU_A_BYTES + NS * BN * 128 : TC_STG_BYTES) + 256 + 128;
}

template <int BN, int NS, int MINB, class EP>
__global__ void __launch_bounds__(TC_THREADS, MINB) gemm_tc_unfold_kernel(const float* __restrict__ X, const float* __restrict__ Wimg, EP ep, int M) {
    constexpr int NK = 16;  // 8 taps x 2 halves of 32 channels
    constexpr int WBYTES = BN * 128;
    constexpr int TCOLS = tc_tmem_cols<BN>();
    constexpr int AREA = TCU_A_BYTES + NS * WBYTES > TC_STG_BYTES ? TCU_A_BYTES + NS * WBYTES : TC_STG_BYTES;
    static_assert(NS >= 2, "ring");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* a_slab = smem_raw;
    unsigned char* w_stage = smem_raw + TCU_A_BYTES;
    uint64_t* full_w = reinterpret_cast<uint64_t*>(smem_raw + AREA);
    uint64_t* mma_done = full_w + NS;
    uint64_t* acc_ready = mma_done + NS;  // single-phase: only thread 0 tracks the ring phases
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);
    float* scratch = reinterpret_cast<float*>(tmem_slot + 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.x * TC_BM;
    if (warp == 0) tmem_alloc<TCOLS>(tmem_slot);
    if (tid == 32) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(full_w + s, 1);
            mbar_init(mma_done + s, 1);
        }
        mbar_init(acc_ready, 1);
        fence_mbar_init();
    }
    ep.init(row0, M);
    // stage rows row0 .. row0+135 (rows past M+7 are never multiplied into a stored output: zero them)
    for (int i = tid; i < TCU_ROWS * 16; i += TC_THREADS) {
        const int r = i >> 4, kq = i & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < M + 7) v = ldg4(X + (long long)(row0 + r) * 64 + kq * 4);
        v.x = tf32r(v.x);
        v.y = tf32r(v.y);
        v.z = tf32r(v.z);
        v.w = tf32r(v.w);
        *reinterpret_cast<float4*>(a_slab + kq * TCU_LBO + r * 16) = v;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t IDESC = umma_idesc_tf32(TC_BM, BN);
    if (warp == 0 && elect_one()) {
        auto issue_w = [&](int c) {
            const int s = c % NS;
            mbar_expect_tx(full_w + s, WBYTES);
            bulk_g2s(w_stage + (size_t)s * WBYTES, Wimg + (size_t)c * (BN * TC_KC), WBYTES, full_w + s);
        };
#pragma unroll
        for (int c = 0; c < NS; ++c) issue_w(c);
        const uint32_t a_base = smem_u32(a_slab);
#pragma unroll 1
        for (int kc = 0; kc < NK; ++kc) {
            const int s = kc % NS;
            mbar_wait(full_w + s, (kc / NS) & 1);
            tc_fence_after();
            const uint32_t w_base = smem_u32(w_stage + (size_t)s * WBYTES);
            const int tap = kc >> 1, half = kc & 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint64_t da = umma_desc(a_base + (half * 8 + 2 * q) * TCU_LBO + tap * 16, TCU_LBO, 128);
                const uint64_t db = umma_desc(w_base + 2 * q * (BN * 16), BN * 16, 128);
                umma_tf32(tmem, da, db, IDESC, (kc > 0 || q > 0) ? 1u : 0u);
            }
            umma_commit(mma_done + s);
            if (kc + NS < NK) {  // refill this stage once its MMAs have drained
                mbar_wait(mma_done + s, (kc / NS) & 1);
                issue_w(kc + NS);
            }
        }
        umma_commit(acc_ready);
    }
    mbar_wait(acc_ready, 0);
    tc_fence_after();
    {
        const int q = warp & 3, hlf = warp >> 2;
        float* stg = reinterpret_cast<float*>(smem_raw) + warp * (32 * TC_STG_LD);
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll 1
        for (int cb = 0; cb < BN / 64; ++cb) {
            const int col0 = hlf * (BN / 2) + cb * 32;
            // all global loads of this 32x32 block are issued first and stay in flight while the
            // accumulator block is read from TMEM and transposed through shared memory
            ep.prep(col0 + c4);
            typename EP::Pre pre[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int row = row0 + q * 32 + p * 4 + rsub;
                pre[p] = ep.load(row < M ? row : M - 1, col0 + c4);
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t v[16];
                tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col0 + 16 * hh, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4*>(stg + lane * TC_STG_LD + 16 * hh + 4 * i) =
                        make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
            __syncwarp();
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int r = p * 4 + rsub;
                const float4 x = *reinterpret_cast<const float4*>(stg + r * TC_STG_LD + c4);
                const int row = row0 + q * 32 + r;
                if (row < M) ep.store4(row, col0 + c4, x, pre[p]);
            }
            __syncwarp();
        }
    }
    ep.finish(scratch);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TCOLS>(tmem);
}

template <int BN, int NS, int MINB, class EP>
inline cudaError_t launch_gemm_tc_unfold(const float* X, const float* Wimg, const EP& ep, int M, cudaStream_t st) {
    auto kern = gemm_tc_unfold_kernel<BN, NS, MINB, EP>;
    const int smem = tcu_smem_bytes<BN, NS>();
    static SmemCfg cfg;  // per instantiation, per device
    if (cudaError_t e = ensure_smem(kern, smem, cfg); e != cudaSuccess) return e;
    kern<<<(M + TC_BM - 1) / TC_BM, TC_THREADS, smem, st>>>(X, Wimg, ep, M);
    return cudaGetLastError();
}

template <int BN, int KTOT, int NS, int MINB, int PF, int NT, class AL, class EP>
inline cudaError_t launch_gemm_tc(const AL& al, const float* Wimg, const EP& ep, int M, cudaStream_t st) {
    auto kern = gemm_tc_kernel<BN, KTOT, NS, MINB, PF, NT, AL, EP>;
    const int smem = tc_smem_bytes<BN, NS, NT>(AL::kExtra);
    static SmemCfg cfg;  // per instantiation, per device
    if (cudaError_t e = ensure_smem(kern, smem, cfg); e != cudaSuccess) return e;
    kern<<<(M + TC_BM - 1) / TC_BM, NT, smem, st>>>(al, Wimg, ep, M);
    return cudaGetLastError();
}

}  // namespace rtfs
