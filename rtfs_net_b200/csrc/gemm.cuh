// Row-tile TF32 GEMM used by every dense contraction on the path:
//     C[M x N] = f(A)[M x K] * W[N x K]^T      (A rows = positions / time steps, K = channels)
// A is produced by a *loader functor* (fused prologue: gLN-apply, PReLU, gateway, TF-AR combine,
// im2col, overlapping unfold view ...), C is consumed by an *epilogue functor* (bias, residual,
// gLN statistics, complex mask ...).  CTA tile 128 x BN, K streamed in chunks of 32 through a
// two-stage shared-memory ring (cp.async for W, register-staged + transformed A).
// Tensor-core path: mma.sync.m16n8k8 TF32, fp32 accumulate (cvt.rna on A; W pre-rounded on host).
// PREC3 = error-compensated 3xTF32 (hi/lo split) for the small-K encoder/decoder convs.
#pragma once
#include "common.cuh"

namespace rtfs {

constexpr int GEMM_BM = 128;
constexpr int GEMM_KC = 32;
constexpr int GEMM_LDS = GEMM_KC + 4;
constexpr int GEMM_THREADS = 256;

template <int BN>
constexpr int gemm_smem_floats(int extra) {
    return 2 * GEMM_BM * GEMM_LDS + 2 * BN * GEMM_LDS + extra;
}

// Per-CTA table of gLN scale/shift for the (at most two) samples a 128-row tile touches.
// tab layout: [2 samples][C][2]  (sc, sh):  y = x*sc + sh
DEVINL void fill_gln_table(float* tab, const GlnRef& r, int b_first, int B, int C, int tid = threadIdx.x, int nthr = blockDim.x) {
    for (int i = tid; i < 2 * C; i += nthr) {
        const int s = i / C, c = i - s * C;
        const int b = b_first + s;
        float sc = 0.f, sh = 0.f;
        if (b < B) {
            float mean, rstd;
            gln_mean_rstd(r.sums, b, r.inv_n, mean, rstd);
            const float gm = __ldg(r.gamma + c);
            sc = rstd * gm;
            sh = __ldg(r.beta + c) - mean * sc;
        }
        tab[2 * i] = sc;
        tab[2 * i + 1] = sh;
    }
}

// ====================================================================== loaders
// contract: init(row0, M, extra_smem) by all threads (a __syncthreads follows);
//           load(i, k): float4 of transformed A[row_i][k..k+3], row_i = row0 + (tid>>3) + 32*i
struct PlainLoader {
    const float* A;
    long long lda;
    int K;  // real K (columns >= K read as zero)
    static constexpr int kExtra = 0;
    int row0_, M_, rs_;
    DEVINL void init(int row0, int M, float*) {
        row0_ = row0 + (threadIdx.x >> 3);
        rs_ = blockDim.x >> 3;
        M_ = M;
    }
    DEVINL float4 load(int i, int k) const {
        const int row = row0_ + rs_ * i;
        if (row < M_ && k < K) return ldg4(A + (long long)row * lda + k);
        return make_float4(0.f, 0.f, 0.f, 0.f);
    }
};

// y = act(gLN(x)) ; ACT: 0 none, 1 ReLU   (audio bottleneck pre_norm + pre_act)
template <int C, int ACT>
struct GlnActLoader {
    const float* A;  // [M][C]
    GlnRef gln;
    int P, B;  // rows per sample, samples
    static constexpr int kExtra = 4 * C;
    int row0_, M_, rs_, bfirst_, split_;
    const float* tab_;
    DEVINL void init(int row0, int M, float* extra) {
        bfirst_ = row0 / P;
        split_ = (bfirst_ + 1) * P;
        fill_gln_table(extra, gln, bfirst_, B, C);
        tab_ = extra;
        row0_ = row0 + (threadIdx.x >> 3);
        rs_ = blockDim.x >> 3;
        M_ = M;
    }
    // producer-group variant (persistent kernel): ptid in [0, nthr)
    DEVINL void init_p(int row0, int M, float* extra, int ptid, int nthr) {
        bfirst_ = row0 / P;
        split_ = (bfirst_ + 1) * P;
        fill_gln_table(extra, gln, bfirst_, B, C, ptid, nthr);
        tab_ = extra;
        row0_ = row0 + (ptid >> 3);
        rs_ = nthr >> 3;
        M_ = M;
    }
    DEVINL float4 load(int i, int k) const {
        const int row = row0_ + rs_ * i;
        if (row >= M_) return make_float4(0.f, 0.f, 0.f, 0.f);
        return xform(ldg4(A + (long long)row * C + k), row, k);
    }
    DEVINL const float* raw(long long row, int k) const { return A + row * C + k; }
    // transform of a raw float4 at (row, k..k+3); valid after init/init_p of the row's tile
    DEVINL float4 xform(float4 x, int row, int k) const {
        const int s = row >= split_ ? 1 : 0;
        const float* tb = tab_ + 2 * (s * C + k);
        const float4 t0 = *reinterpret_cast<const float4*>(tb);
        const float4 t1 = *reinterpret_cast<const float4*>(tb + 4);
        float4 y;
        y.x = fmaf(x.x, t0.x, t0.y);
        y.y = fmaf(x.y, t0.z, t0.w);
        y.z = fmaf(x.z, t1.x, t1.y);
        y.w = fmaf(x.w, t1.z, t1.w);
        if (ACT == 1) {
            y.x = fmaxf(y.x, 0.f);
            y.y = fmaxf(y.y, 0.f);
            y.z = fmaxf(y.z, 0.f);
            y.w = fmaxf(y.w, 0.f);
        }
        return y;
    }
};

// gateway: r = PReLU(wg[c]*x + bg[c])   (depthwise 1x1 conv + PReLU, tdanet.py:34-41)
struct GateLoader {
    const float* A;  // [M][C]
    const float* wg;
    const float* bg;
    const float* slope;  // 1 float
    int C;
    // the 2 x 256 gateway coefficients live in shared memory: with ~200 KB of the SM's L1/shared array carved out for
    // shared memory the streaming A loads evict them from L1, and a global re-load would sit in every chunk's dependent chain
    static constexpr int kExtra = 512;
    static constexpr bool kTileInvariant = true;  // the table does not depend on the tile: a persistent kernel fills it once
    int row0_, M_, rs_;
    float a_;
    const float* tab_;
    DEVINL void init(int row0, int M, float* extra) { init_p(row0, M, extra, threadIdx.x, blockDim.x); }
    DEVINL void init_p(int row0, int M, float* extra, int ptid, int nthr) {
        row0_ = row0 + (ptid >> 3);
        rs_ = nthr >> 3;
        M_ = M;
        a_ = __ldg(slope);
        tab_ = extra;
        for (int i = ptid; i < 128; i += nthr) {
            const float4 w = ldg4(wg + 4 * (i & 63));
            const float4 bb = ldg4(bg + 4 * (i & 63));
            if (i < 64) *reinterpret_cast<float4*>(extra + 4 * i) = w;
            else *reinterpret_cast<float4*>(extra + 4 * i) = bb;
        }
    }
    DEVINL float4 load(int i, int k) const {
        const int row = row0_ + rs_ * i;
        if (row >= M_) return make_float4(0.f, 0.f, 0.f, 0.f);
        return xform(ldg4(A + (long long)row * C + k), row, k);
    }
    DEVINL const float* raw(long long row, int k) const { return A + row * C + k; }
    DEVINL float4 xform(float4 x, int, int k) const {
        const float4 w = *reinterpret_cast<const float4*>(tab_ + k), b = *reinterpret_cast<const float4*>(tab_ + 256 + k);
        float4 y;
        y.x = prelu(fmaf(w.x, x.x, b.x), a_);
        y.y = prelu(fmaf(w.y, x.y, b.y), a_);
        y.z = prelu(fmaf(w.z, x.z, b.z), a_);
        y.w = prelu(fmaf(w.w, x.w, b.w), a_);
        return y;
    }
};

// y = PReLU(x)   (mask generator front, mask_generator.py:45-60)
struct PreluLoader {
    const float* A;
    const float* slope;
    int C;
    static constexpr int kExtra = 0;
    int row0_, M_, rs_;
    float a_;
    DEVINL void init(int row0, int M, float*) {
        row0_ = row0 + (threadIdx.x >> 3);
        rs_ = blockDim.x >> 3;
        M_ = M;
        a_ = __ldg(slope);
    }
    DEVINL void init_p(int row0, int M, float*, int ptid, int nthr) {
        row0_ = row0 + (ptid >> 3);
        rs_ = nthr >> 3;
        M_ = M;
        a_ = __ldg(slope);
    }
    DEVINL float4 load(int i, int k) const {
        const int row = row0_ + rs_ * i;
        if (row >= M_) return make_float4(0.f, 0.f, 0.f, 0.f);
        return xform(ldg4(A + (long long)row * C + k), row, k);
    }
    DEVINL const float* raw(long long row, int k) const { return A + row * C + k; }
    DEVINL float4 xform(float4 x, int, int) const {
        x.x = prelu(x.x, a_);
        x.y = prelu(x.y, a_);
        x.z = prelu(x.z, a_);
        x.w = prelu(x.w, a_);
        return x;
    }
};

// TF-AR combine of the last fusion (tdanet.py:127) feeding residual_conv:
//   e = gLN_l(lec_pre) * up(sigmoid(gLN_g(ggc_pre))) + up(gLN_e(gec_pre)) + gLN_d(d0_pre)
// lec_pre, d0_pre at full resolution (T,F); ggc_pre, gec_pre at (Tc,Fc) nearest-upsampled.
struct TfarLoader {
    const float* lec;   // [B][T][F][64]
    const float* d0;    // [B][T][F][64]
    const float* ggc;   // [B][Tc][Fc][64]
    const float* gec;   // [B][Tc][Fc][64]
    GlnRef n_l, n_d, n_g, n_e;
    int T, F, Tc, Fc, B;
    static constexpr int kExtra = 4 * 4 * 64;  // 4 norms x [2][64][2]
    static constexpr bool kBatched = true;     // raw_load / xform_raw: the tcgen05 kernels issue a chunk's loads as one batch
    int row0_, M_, rs_, bfirst_, P_;
    const float* tab_;
    long long offc_[4];
    int s_[4];
    struct Raw {
        float4 l, d, g, e;
    };
    DEVINL void init(int row0, int M, float* extra) { init_p(row0, M, extra, threadIdx.x, blockDim.x); }
    DEVINL void init_p(int row0, int M, float* extra, int ptid, int nthr) {
        P_ = T * F;
        bfirst_ = row0 / P_;
        // the four scale/shift tables in one pass: entry i = (norm i>>7, sample (i>>6)&1, channel i&63); the loads of
        // all of a thread's entries are independent, so the fill costs one L2 round trip instead of one per table
#pragma unroll 2
        for (int i = ptid; i < 512; i += nthr) {
            const int nb = i >> 7, s = (i >> 6) & 1, c = i & 63;
            // selects on the fields (a reference select would spill the four structs to local memory)
            const double* sums = nb == 0 ? n_l.sums : nb == 1 ? n_d.sums : nb == 2 ? n_g.sums : n_e.sums;
            const float* gamma = nb == 0 ? n_l.gamma : nb == 1 ? n_d.gamma : nb == 2 ? n_g.gamma : n_e.gamma;
            const float* beta = nb == 0 ? n_l.beta : nb == 1 ? n_d.beta : nb == 2 ? n_g.beta : n_e.beta;
            const double inv_n = nb == 0 ? n_l.inv_n : nb == 1 ? n_d.inv_n : nb == 2 ? n_g.inv_n : n_e.inv_n;
            const int b = bfirst_ + s;
            float sc = 0.f, sh = 0.f;
            if (b < B) {
                float mean, rstd;
                gln_mean_rstd(sums, b, inv_n, mean, rstd);
                sc = rstd * __ldg(gamma + c);
                sh = __ldg(beta + c) - mean * sc;
            }
            *reinterpret_cast<float2*>(extra + 2 * i) = make_float2(sc, sh);
        }
        tab_ = extra;
        row0_ = row0 + (ptid >> 3);
        rs_ = nthr >> 3;
        M_ = M;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int row = row0_ + rs_ * i;
            if (row >= M) row = M - 1;
            const int b = row / P_;
            const int p = row - b * P_;
            const int t = p / F, f = p - t * F;
            const int tc = nearest_src32(t, Tc, T), fc = nearest_src32(f, Fc, F);
            offc_[i] = (((long long)b * Tc + tc) * Fc + fc) * 64;
            s_[i] = b - bfirst_;
        }
    }
    DEVINL float nrm(const float* tab, int s, int k, float x) const {
        const float* tb = tab + 2 * (s * 64 + k);
        return fmaf(x, tb[0], tb[1]);
    }
    // the four operand loads of row i (clamped to the last valid row: xform_raw zeroes rows past M)
    DEVINL Raw raw_load(int i, int k) const {
        int row = row0_ + rs_ * i;
        row = row < M_ ? row : M_ - 1;
        Raw r;
        r.l = ldg4(lec + (long long)row * 64 + k);
        r.d = ldg4(d0 + (long long)row * 64 + k);
        r.g = ldg4(ggc + offc_[i] + k);
        r.e = ldg4(gec + offc_[i] + k);
        return r;
    }
    DEVINL float4 xform_raw(const Raw& r, int i, int k) const {
        if (row0_ + rs_ * i >= M_) return make_float4(0.f, 0.f, 0.f, 0.f);
        const int s = s_[i];
        float4 y;
        y.x = nrm(tab_, s, k, r.l.x) * sigmoidf_fast(nrm(tab_ + 512, s, k, r.g.x)) + nrm(tab_ + 768, s, k, r.e.x) + nrm(tab_ + 256, s, k, r.d.x);
        y.y = nrm(tab_, s, k + 1, r.l.y) * sigmoidf_fast(nrm(tab_ + 512, s, k + 1, r.g.y)) + nrm(tab_ + 768, s, k + 1, r.e.y) + nrm(tab_ + 256, s, k + 1, r.d.y);
        y.z = nrm(tab_, s, k + 2, r.l.z) * sigmoidf_fast(nrm(tab_ + 512, s, k + 2, r.g.z)) + nrm(tab_ + 768, s, k + 2, r.e.z) + nrm(tab_ + 256, s, k + 2, r.d.z);
        y.w = nrm(tab_, s, k + 3, r.l.w) * sigmoidf_fast(nrm(tab_ + 512, s, k + 3, r.g.w)) + nrm(tab_ + 768, s, k + 3, r.e.w) + nrm(tab_ + 256, s, k + 3, r.d.w);
        return y;
    }
    DEVINL float4 load(int i, int k) const { return xform_raw(raw_load(i, k), i, k); }
};

// im2col of the 2-channel spectrogram for the 3x3 encoder conv (encoder.py:147-173):
//   k = (i*3 + j)*2 + ci  ->  spec[b][t+i-1][f+j-1][ci],  K = 18 (padded to 32)
struct Im2colLoader {
    const float* spec;  // [B][T][F][2]
    int T, F;
    static constexpr int kExtra = 0;
    int row0_, M_, rs_;
    DEVINL void init(int row0, int M, float*) {
        row0_ = row0 + (threadIdx.x >> 3);
        rs_ = blockDim.x >> 3;
        M_ = M;
    }
    DEVINL float2 tap(int b, int t, int f, int tapi) const {
        const int i = tapi / 3, j = tapi - 3 * i;
        const int tt = t + i - 1, ff = f + j - 1;
        if (tapi >= 9 || tt < 0 || tt >= T || ff < 0 || ff >= F) return make_float2(0.f, 0.f);
        return ldg2(spec + (((long long)b * T + tt) * F + ff) * 2);
    }
    DEVINL float4 load(int i, int k) const {
        const int row = row0_ + rs_ * i;
        if (row >= M_ || k >= 18) return make_float4(0.f, 0.f, 0.f, 0.f);
        const int P = T * F;
        const int b = row / P, p = row - b * P;
        const int t = p / F, f = p - t * F;
        const float2 u = tap(b, t, f, k >> 1), v = tap(b, t, f, (k >> 1) + 1);
        return make_float4(u.x, u.y, v.x, v.y);
    }
};

// Error-compensated (3xTF32) variant for the tcgen05 kernels: K' = 96 = [hi | lo | hi] of the 32 im2col columns, to be
// multiplied with the weight image [W_hi | W_hi | W_lo]: hi.W_hi + lo.W_hi + hi.W_lo accumulates in fp32.
struct Im2colLoader3x {
    Im2colLoader base;
    static constexpr int kExtra = 0;
    // per-tile state of the persistent producers: the (b, t, f) decomposition of a thread's 4 rows is done once per tile
    // (one division, then increments) and its two taps are fixed by its K piece, so a load is two predicated 8-byte reads
    const float* rp_[4];  // &spec[b][t][f][0] of row i
    int t_[4], f_[4];
    int oa_, ob_, dia_, dja_, dib_, djb_;  // element offsets and (dt, df) of taps 2*kq and 2*kq+1 (dia_ = 9: no such tap)
    DEVINL void init(int row0, int M, float* e) { init_p(row0, M, e, threadIdx.x, blockDim.x); }
    DEVINL void init_p(int row0, int M, float*, int ptid, int nthr) {
        base.row0_ = row0 + (ptid >> 3);
        base.rs_ = nthr >> 3;
        base.M_ = M;
        const int T = base.T, F = base.F, P = T * F;
        const int ta = 2 * (ptid & 7), tb = ta + 1;
        dia_ = ta < 9 ? ta / 3 - 1 : 9;
        dja_ = ta < 9 ? ta % 3 - 1 : 0;
        dib_ = tb < 9 ? tb / 3 - 1 : 9;
        djb_ = tb < 9 ? tb % 3 - 1 : 0;
        oa_ = (dia_ * F + dja_) * 2;
        ob_ = (dib_ * F + djb_) * 2;
        int row = base.row0_ < M ? base.row0_ : M - 1;
        int b = row / P, p = row - b * P;
        int t = p / F, f = p - t * F;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            rp_[i] = base.spec + (((long long)b * T + t) * F + f) * 2;
            t_[i] = t;
            f_[i] = f;
            f += base.rs_;                 // next row of this thread: rs_ (= 32 or 64) positions further
            while (f >= F) {
                f -= F;
                if (++t == T) {
                    t = 0;
                    ++b;
                }
            }
        }
    }
    DEVINL float4 load(int i, int k) const {
        if (base.row0_ + base.rs_ * i >= base.M_) return make_float4(0.f, 0.f, 0.f, 0.f);
        const int T = base.T, F = base.F;
        const int ta = t_[i] + dia_, fa = f_[i] + dja_, tb = t_[i] + dib_, fb = f_[i] + djb_;
        float2 u = make_float2(0.f, 0.f), w = make_float2(0.f, 0.f);
        if (dia_ != 9 && ta >= 0 && ta < T && fa >= 0 && fa < F) u = ldg2(rp_[i] + oa_);
        if (dib_ != 9 && tb >= 0 && tb < T && fb >= 0 && fb < F) w = ldg2(rp_[i] + ob_);
        const float4 v = make_float4(u.x, u.y, w.x, w.y);
        const float4 hi = make_float4(tf32r(v.x), tf32r(v.y), tf32r(v.z), tf32r(v.w));
        if ((k >> 5) != 1) return hi;
        return make_float4(tf32r(v.x - hi.x), tf32r(v.y - hi.y), tf32r(v.z - hi.z), tf32r(v.w - hi.w));
    }
};

// ====================================================================== epilogues
// contract: init(row0, M); store(row, col, v0, v1) for (row,col),(row,col+1); finish(scratch) by all threads
struct StoreEpi {
    float* C;
    long long ldc;
    const float* bias;  // may be null
    DEVINL void init(int, int) {}
    DEVINL void store(int row, int col, float v0, float v1) {
        if (bias) {
            v0 += __ldg(bias + col);
            v1 += __ldg(bias + col + 1);
        }
        *reinterpret_cast<float2*>(C + (long long)row * ldc + col) = make_float2(v0, v1);
    }
    DEVINL void finish(float*) {}
};

// store (+bias) and accumulate per-sample (sum, sumsq) of what was stored (gLN statistics)
struct StatsEpi {
    float* C;
    long long ldc;
    const float* bias;  // may be null
    double* sums;       // [B][2]
    int P, B;
    int bfirst_, split_;
    float s0_, q0_, s1_, q1_;
    DEVINL void init(int row0, int) {
        bfirst_ = row0 / P;
        split_ = (bfirst_ + 1) * P;
        s0_ = q0_ = s1_ = q1_ = 0.f;
    }
    DEVINL void store(int row, int col, float v0, float v1) {
        if (bias) {
            v0 += __ldg(bias + col);
            v1 += __ldg(bias + col + 1);
        }
        *reinterpret_cast<float2*>(C + (long long)row * ldc + col) = make_float2(v0, v1);
        if (row < split_) {
            s0_ += v0 + v1;
            q0_ += v0 * v0 + v1 * v1;
        } else {
            s1_ += v0 + v1;
            q1_ += v0 * v0 + v1 * v1;
        }
    }
    DEVINL void finish(float* scratch) {
        block_stats_atomic(s0_, q0_, sums + 2 * bfirst_, scratch);
        block_stats_atomic(s1_, q1_, (bfirst_ + 1 < B) ? sums + 2 * (bfirst_ + 1) : nullptr, scratch);
    }
};

// residual_conv epilogue (tdanet.py:131): out = acc + bias + gateway(x) [+ a1 -> next block input]
struct ResidOutEpi {
    float* out;        // [M][256]
    const float* bias; // [256]
    const float* x;    // block input [M][256]
    const float* wg;
    const float* bg;
    const float* slope;
    const float* a1;   // may be null
    float a_;
    DEVINL void init(int, int) { a_ = __ldg(slope); }
    DEVINL void store(int row, int col, float v0, float v1) {
        const long long o = (long long)row * 256 + col;
        const float2 xx = ldg2(x + o);
        const float2 w = ldg2(wg + col), b = ldg2(bg + col), bi = ldg2(bias + col);
        v0 += bi.x + prelu(fmaf(w.x, xx.x, b.x), a_);
        v1 += bi.y + prelu(fmaf(w.y, xx.y, b.y), a_);
        if (a1) {
            const float2 aa = ldg2(a1 + o);
            v0 += aa.x;
            v1 += aa.y;
        }
        *reinterpret_cast<float2*>(out + o) = make_float2(v0, v1);
    }
    DEVINL void finish(float*) {}
};

// S^3 mask epilogue (mask_generator.py:67-99).  GEMM columns are interleaved on the host:
// col 2c = real-half channel c, col 2c+1 = imag-half channel c+128, so one thread holds the pair.
struct MaskEpi {
    float* z;           // [M][256]  (0..127 real, 128..255 imag)
    const float* bias;  // interleaved like the columns
    const float* a0;    // encoder output [M][256]
    DEVINL void init(int, int) {}
    DEVINL void store(int row, int col, float v0, float v1) {
        const float2 bi = ldg2(bias + col);
        const float mr = fmaxf(v0 + bi.x, 0.f), mi = fmaxf(v1 + bi.y, 0.f);
        const int c = col >> 1;
        const long long o = (long long)row * 256 + c;
        const float er = __ldg(a0 + o), ei = __ldg(a0 + o + 128);
        z[o] = er * mr - ei * mi;
        z[o + 128] = er * mi + ei * mr;
    }
    DEVINL void finish(float*) {}
};

// ConvTranspose1d-as-GEMM epilogue of the dual-path RNN (rnn_layers.py:153-160):
// GEMM row = seq*(S+7) + s ; out[pos(seq,s)] = acc + bias + residual[pos(seq,s)]
struct ConvTEpi {
    float* out;          // (B,Tc,Fc,64)
    const float* resid;  // same layout
    const float* bias;   // [64]
    int S;               // sequence length
    int n_other;         // number of sequences per batch item (freq path: Tc ; time path: Fc)
    int time_path;       // 0: seq=(b,t), s=f ; 1: seq=(b,f), s=t
    int Tc, Fc;
    DEVINL void init(int, int) {}
    DEVINL void store(int row, int col, float v0, float v1) {
        const int seq = row / (S + 7), s = row - seq * (S + 7);
        if (s >= S) return;
        const int b = seq / n_other, o = seq - b * n_other;
        const int t = time_path ? s : o, f = time_path ? o : s;
        const long long off = ((((long long)b * Tc + t) * Fc) + f) * 64 + col;
        const float2 r = ldg2(resid + off), bi = ldg2(bias + col);
        *reinterpret_cast<float2*>(out + off) = make_float2(v0 + bi.x + r.x, v1 + bi.y + r.y);
    }
    DEVINL void finish(float*) {}
};

// ====================================================================== kernel
template <int BN, int KTOT, bool PREC3, class AL, class EP>
__global__ void __launch_bounds__(GEMM_THREADS, (BN <= 128 ? 2 : 1))
gemm_tf32_kernel(AL al, const float* __restrict__ W, EP ep, int M, int N) {
    constexpr int BM = GEMM_BM, KC = GEMM_KC, LDS = GEMM_LDS;
    constexpr int NI = BN / 16;
    constexpr int NK = KTOT / KC;
    static_assert(KTOT % KC == 0, "pad K to a multiple of 32");
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* Ws = smem + 2 * BM * LDS;
    float* extra = Ws + 2 * BN * LDS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const int nt = (N + BN - 1) / BN;
    const int ntile = blockIdx.x % nt, mtile = blockIdx.x / nt;
    const int row0 = mtile * BM, col0 = ntile * BN;

    al.init(row0, M, extra);
    ep.init(row0, M);
    __syncthreads();

    float acc[2][NI][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < NI; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

    auto issue_w = [&](int kc, int buf) {
#pragma unroll
        for (int i = 0; i < BN / 32; ++i) {
            const int idx = tid + i * GEMM_THREADS;
            const int n = idx >> 3, c4 = idx & 7;
            const bool valid = (col0 + n) < N;
            const int nn = valid ? (col0 + n) : (N - 1);
            cp_async16(Ws + ((size_t)buf * BN + n) * LDS + c4 * 4, W + (size_t)nn * KTOT + kc * KC + c4 * 4, valid);
        }
        cp_async_commit();
    };
    auto store_a = [&](const float4 (&r)[4], int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = (tid >> 3) + 32 * i, c4 = tid & 7;
            float4 v = r[i];
            if (!PREC3) {
                v.x = tf32r(v.x);
                v.y = tf32r(v.y);
                v.z = tf32r(v.z);
                v.w = tf32r(v.w);
            }
            *reinterpret_cast<float4*>(As + ((size_t)buf * BM + row) * LDS + c4 * 4) = v;
        }
    };

    float4 areg[4];
    issue_w(0, 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) areg[i] = al.load(i, (tid & 7) * 4);
    store_a(areg, 0);
    cp_async_wait<0>();
    __syncthreads();

    for (int kc = 0; kc < NK; ++kc) {
        const int cur = kc & 1, nxt = cur ^ 1;
        const bool more = (kc + 1) < NK;
        if (more) {
            issue_w(kc + 1, nxt);
#pragma unroll
            for (int i = 0; i < 4; ++i) areg[i] = al.load(i, (kc + 1) * KC + (tid & 7) * 4);
        }
        const float* Ab = As + (size_t)cur * BM * LDS + (wm * 32) * LDS;
        const float* Wb = Ws + (size_t)cur * BN * LDS + (wn * (BN / 2)) * LDS;
#pragma unroll
        for (int ks = 0; ks < KC / 8; ++ks) {
            uint32_t af[2][4], al_[2][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const float* p = Ab + (mi * 16 + g) * LDS + ks * 8 + t;
                const float x0 = p[0], x1 = p[8 * LDS], x2 = p[4], x3 = p[8 * LDS + 4];
                if (PREC3) {
                    af[mi][0] = f2tf32(x0);
                    af[mi][1] = f2tf32(x1);
                    af[mi][2] = f2tf32(x2);
                    af[mi][3] = f2tf32(x3);
                    al_[mi][0] = f2tf32(x0 - __uint_as_float(af[mi][0]));
                    al_[mi][1] = f2tf32(x1 - __uint_as_float(af[mi][1]));
                    al_[mi][2] = f2tf32(x2 - __uint_as_float(af[mi][2]));
                    al_[mi][3] = f2tf32(x3 - __uint_as_float(af[mi][3]));
                } else {
                    af[mi][0] = __float_as_uint(x0);
                    af[mi][1] = __float_as_uint(x1);
                    af[mi][2] = __float_as_uint(x2);
                    af[mi][3] = __float_as_uint(x3);
                }
            }
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                const float* q = Wb + (ni * 8 + g) * LDS + ks * 8 + t;
                const float w0 = q[0], w1 = q[4];
                uint32_t bf[2], bl[2];
                if (PREC3) {
                    bf[0] = f2tf32(w0);
                    bf[1] = f2tf32(w1);
                    bl[0] = f2tf32(w0 - __uint_as_float(bf[0]));
                    bl[1] = f2tf32(w1 - __uint_as_float(bf[1]));
                } else {
                    bf[0] = __float_as_uint(w0);
                    bf[1] = __float_as_uint(w1);
                }
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    if (PREC3) {
                        mma_tf32(acc[mi][ni], al_[mi], bf);
                        mma_tf32(acc[mi][ni], af[mi], bl);
                    }
                    mma_tf32(acc[mi][ni], af[mi], bf);
                }
            }
        }
        if (more) {
            store_a(areg, nxt);
            cp_async_wait<0>();
        }
        __syncthreads();
    }

#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            const int row = row0 + wm * 32 + mi * 16 + g;
            const int col = col0 + wn * (BN / 2) + ni * 8 + 2 * t;
            if (col < N) {
                if (row < M) ep.store(row, col, acc[mi][ni][0], acc[mi][ni][1]);
                if (row + 8 < M) ep.store(row + 8, col, acc[mi][ni][2], acc[mi][ni][3]);
            }
        }
    }
    ep.finish(smem);
}

template <int BN, int KTOT, bool PREC3, class AL, class EP>
inline cudaError_t launch_gemm(const AL& al, const float* W, const EP& ep, int M, int N, cudaStream_t st) {
    auto kern = gemm_tf32_kernel<BN, KTOT, PREC3, AL, EP>;
    const size_t smem = sizeof(float) * gemm_smem_floats<BN>(AL::kExtra);
    static SmemCfg cfg;  // per instantiation, per device
    if (cudaError_t e = ensure_smem(kern, (int)smem, cfg); e != cudaSuccess) return e;
    const int nt = (N + BN - 1) / BN;
    const int mt = (M + GEMM_BM - 1) / GEMM_BM;
    kern<<<mt * nt, GEMM_THREADS, smem, st>>>(al, W, ep, M, N);
    return cudaGetLastError();
}

}  // namespace rtfs
