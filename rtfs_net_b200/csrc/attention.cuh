// TF self-attention of the RTFS block (reference: MultiHeadSelfAttention2D.forward,
// layers/attention.py:149-189; ConvActNorm = 1x1 conv -> PReLU -> LayerNormalization4D over (C,F),
// layers/conv_layers.py:201-205, layers/normalizations.py:20-37).
//   rowblock_ln_kernel<96,0> : all 12 Q/K/V head convs of one (b,t) frame as one 64x96x64 GEMM,
//                              PReLU per conv, LN over (E,F) per conv, written as per-head token rows
//   attn_core_kernel         : softmax(Q K^T / sqrt(E*F)) V per (b, head, 32-query tile)
//   rowblock_ln_kernel<64,1> : concat-projection 64->64 + PReLU + LN over (C,F) + residual
// Tokens are time frames; a token's feature vector is the (f,e)-ordered slab of its frame, which
// in the channels-last layout is contiguous.
#pragma once
#include "common.cuh"

namespace rtfs {

constexpr int AT_NF = 64;  // n_freqs of the attention layer (config: n_freqs: 64)

struct RowblockArgs {
    const float* x;      // GEMM input  (B*Tc, 64 f, 64 c)
    const float* resid;  // MODE 1: residual, same layout as out
    const float* W;      // [N][64] tf32-rounded
    const float* bias;   // [N]
    const float* slope;  // [ngroups]
    const float* gamma;  // per group, [f*E+e] order, groups concatenated in column order
    const float* beta;
    float* q;  // MODE 0 outputs: (B,H,Tc,256) (B,H,Tc,256) (B,H,Tc,1024)
    float* k;
    float* v;
    float* out;  // MODE 1 output (B*Tc, 64, 64)
    int B, Tc, H;
};

// MODE 0: N = 96 = [Q: H x 4][K: H x 4][V: H x 16] (H = 4) ; MODE 1: N = 64, one group
template <int N, int MODE>
__global__ void __launch_bounds__(128) rowblock_ln_kernel(RowblockArgs a) {
    constexpr int LD = 68, LDY = N + 1, NI = N / 8;
    extern __shared__ __align__(16) float sm[];
    float* Xs = sm;                  // [64][LD]
    float* Wsm = Xs + 64 * LD;       // [N][LD]
    float* Ys = sm;                  // [64][LDY]  aliases Xs/Wsm once the MMAs have consumed them (more CTAs per SM)
    float* stat = sm + ((64 + N) * LD > 64 * LDY ? (64 + N) * LD : 64 * LDY);  // [12][2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const long long bt = blockIdx.x;
    const float* xin = a.x + bt * (64 * 64);
    for (int i = tid; i < 64 * 16; i += 128) {
        const int r = i >> 4, c4 = i & 15;
        float4 v = ldg4(xin + r * 64 + c4 * 4);
        v.x = tf32r(v.x);
        v.y = tf32r(v.y);
        v.z = tf32r(v.z);
        v.w = tf32r(v.w);
        *reinterpret_cast<float4*>(Xs + r * LD + c4 * 4) = v;
    }
    for (int i = tid; i < N * 16; i += 128) {
        const int r = i >> 4, c4 = i & 15;
        *reinterpret_cast<float4*>(Wsm + r * LD + c4 * 4) = ldg4(a.W + r * 64 + c4 * 4);
    }
    __syncthreads();
    float acc[NI][4];
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        uint32_t af[4];
        const float* p = Xs + (warp * 16 + g) * LD + ks * 8 + t;
        af[0] = __float_as_uint(p[0]);
        af[1] = __float_as_uint(p[8 * LD]);
        af[2] = __float_as_uint(p[4]);
        af[3] = __float_as_uint(p[8 * LD + 4]);
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            const float* q = Wsm + (ni * 8 + g) * LD + ks * 8 + t;
            uint32_t bf[2] = {__float_as_uint(q[0]), __float_as_uint(q[4])};
            mma_tf32(acc[ni], af, bf);
        }
    }
    __syncthreads();  // every warp is done reading Xs / Wsm: Ys may overwrite them
    // bias + PReLU -> Ys
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) {
        const int col = ni * 8 + 2 * t;
        int grp;
        if (MODE == 0) grp = col < 16 ? (col >> 2) : (col < 32 ? 4 + ((col - 16) >> 2) : 8 + ((col - 32) >> 4));
        else grp = 0;
        const float sl = __ldg(a.slope + grp);
        const float b0 = __ldg(a.bias + col), b1 = __ldg(a.bias + col + 1);
        const int r = warp * 16 + g;
        Ys[r * LDY + col] = prelu(acc[ni][0] + b0, sl);
        Ys[r * LDY + col + 1] = prelu(acc[ni][1] + b1, sl);
        Ys[(r + 8) * LDY + col] = prelu(acc[ni][2] + b0, sl);
        Ys[(r + 8) * LDY + col + 1] = prelu(acc[ni][3] + b1, sl);
    }
    __syncthreads();
    constexpr int NG = MODE == 0 ? 12 : 1;
    // LayerNormalization4D statistics per group over (E columns x 64 rows), two-pass
    for (int gi = warp; gi < NG; gi += 4) {
        int col0, E;
        if (MODE == 0) {
            if (gi < 4) { col0 = gi * 4; E = 4; }
            else if (gi < 8) { col0 = 16 + (gi - 4) * 4; E = 4; }
            else { col0 = 32 + (gi - 8) * 16; E = 16; }
        } else { col0 = 0; E = 64; }
        const int n = 64 * E;
        float s = 0.f;
        for (int i = lane; i < n; i += 32) s += Ys[(i / E) * LDY + col0 + (i % E)];
        s = warp_sum(s);
        const float mu = s / (float)n;
        float q = 0.f;
        for (int i = lane; i < n; i += 32) {
            const float d = Ys[(i / E) * LDY + col0 + (i % E)] - mu;
            q += d * d;
        }
        q = warp_sum(q);
        if (lane == 0) {
            stat[2 * gi] = mu;
            stat[2 * gi + 1] = 1.f / sqrtf(q / (float)n + RTFS_EPS);
        }
    }
    __syncthreads();
    if (MODE == 0) {
        const int b = (int)(bt / a.Tc), tt = (int)(bt % a.Tc);
        int goff = 0;
        for (int gi = 0; gi < 12; ++gi) {
            int col0, E, h;
            float* dst;
            if (gi < 4) { h = gi; col0 = h * 4; E = 4; dst = a.q + (((long long)b * a.H + h) * a.Tc + tt) * 256; }
            else if (gi < 8) { h = gi - 4; col0 = 16 + h * 4; E = 4; dst = a.k + (((long long)b * a.H + h) * a.Tc + tt) * 256; }
            else { h = gi - 8; col0 = 32 + h * 16; E = 16; dst = a.v + (((long long)b * a.H + h) * a.Tc + tt) * 1024; }
            const int n = 64 * E;
            const float mu = stat[2 * gi], rs = stat[2 * gi + 1];
            for (int i = tid; i < n; i += 128) {
                const float y = Ys[(i / E) * LDY + col0 + (i % E)];
                dst[i] = tf32r((y - mu) * rs * __ldg(a.gamma + goff + i) + __ldg(a.beta + goff + i));
            }
            goff += n;
        }
    } else {
        const float mu = stat[0], rs = stat[1];
        const float* res = a.resid + bt * 4096;
        float* dst = a.out + bt * 4096;
        for (int i = tid; i < 4096; i += 128) {
            const float y = Ys[(i >> 6) * LDY + (i & 63)];
            dst[i] = (y - mu) * rs * __ldg(a.gamma + i) + __ldg(a.beta + i) + __ldg(res + i);
        }
    }
}

template <int N>
constexpr int rowblock_smem_floats() {
    return ((64 + N) * 68 > 64 * (N + 1) ? (64 + N) * 68 : 64 * (N + 1)) + 32;
}

// ---------------------------------------------------------------- attention core
struct AttnArgs {
    const float* q;  // (B*H, Tc, 256) tf32-rounded
    const float* k;
    const float* v;  // (B*H, Tc, 1024), inner index f*16 + j
    float* o;        // (B, Tc, 64, 64), channel = h*16 + j
    int Tc, H, tk_pad;  // tk_pad = ceil(Tc/64)*64
    float scale;        // 1/sqrt(E*F)
};

constexpr int AT_KR = 32, AT_LDQ = 260, AT_LDV = 136;
// Q tile + score tile + one K/V chunk of AT_KR keys: 83 KB at 125 frames -> two CTAs per SM overlap each other's
// load -> wait -> MMA rounds
inline int attn_smem_floats(int tk_pad, int qt) {  // Q tile + score tile + max(one K chunk, two V chunks)
    const int nb = qt == 64 ? 4 : 2;  // V chunk buffers (64-query tiles: one CTA per SM, room for a deeper ring and two K buffers)
    const int kv = AT_KR * AT_LDQ > nb * AT_KR * AT_LDV ? AT_KR * AT_LDQ : nb * AT_KR * AT_LDV;
    return qt * AT_LDQ + qt * (tk_pad + 4) + kv;
}

// AT_QT queries per CTA (8 threads per query): K and V are re-read from L2 once per query tile, which is what bounds the
// kernel (4 tiles of 32 queries at 125 frames = ~6 TB/s of L2 traffic), so 64-query tiles (512 threads, one CTA per SM)
// halve it; the 32-query variant (two CTAs per SM) remains for short sequences.
template <int AT_QT>
__global__ void __launch_bounds__(AT_QT * 8, AT_QT == 32 ? 2 : 1) attn_core_kernel(AttnArgs a) {
    constexpr int NT = AT_QT * 8, WM = AT_QT / 16;
    constexpr int NB = AT_QT == 64 ? 4 : 2;      // V chunks in the ring (NB - 1 in flight)
    constexpr bool KDB = NB * AT_KR * AT_LDV >= 2 * AT_KR * AT_LDQ;  // room for two K chunks: double-buffered QK^T phase
    extern __shared__ __align__(16) float sm[];
    const int SLD = a.tk_pad + 4;
    float* Qs = sm;                    // [32][260]
    float* Ss = Qs + AT_QT * AT_LDQ;   // [32][SLD]
    float* KV = Ss + AT_QT * SLD;      // K chunk [32][260]  /  V chunk [32][136]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int wm = warp % WM, wn = warp / WM;
    const int bh = blockIdx.y, q0 = blockIdx.x * AT_QT;
    const int Tc = a.Tc;
    const float* Qg = a.q + (long long)bh * Tc * 256;
    const float* Kg = a.k + (long long)bh * Tc * 256;
    const float* Vg = a.v + (long long)bh * Tc * 1024;

    for (int i = tid; i < AT_QT * 64; i += NT) {
        const int r = i >> 6, c4 = i & 63;
        const bool valid = (q0 + r) < Tc;
        cp_async16(Qs + r * AT_LDQ + c4 * 4, Qg + (long long)(valid ? q0 + r : 0) * 256 + c4 * 4, valid);
    }
    cp_async_commit();
    const int nkc = a.tk_pad / AT_KR;
    // ---- S = scale * Q K^T : warp (wm, wn) -> 16 queries x 8 keys of the chunk
    auto issue_k = [&](int kc_) {
        float* dstb = KV + (KDB ? (kc_ & 1) * (AT_KR * AT_LDQ) : 0);
        for (int i = tid; i < AT_KR * 64; i += NT) {
            const int r = i >> 6, c4 = i & 63;
            const int key = kc_ * AT_KR + r;
            const bool valid = key < Tc;
            cp_async16(dstb + r * AT_LDQ + c4 * 4, Kg + (long long)(valid ? key : 0) * 256 + c4 * 4, valid);
        }
        cp_async_commit();
    };
    if (KDB) issue_k(0);
    for (int kc = 0; kc < nkc; ++kc) {
        if (KDB) {
            if (kc + 1 < nkc) {
                issue_k(kc + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
        } else {
            issue_k(kc);
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* Kb = KV + (KDB ? (kc & 1) * (AT_KR * AT_LDQ) : 0);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (int ks = 0; ks < 32; ++ks) {
            uint32_t af[4];
            const float* p = Qs + (wm * 16 + g) * AT_LDQ + ks * 8 + t;
            af[0] = __float_as_uint(p[0]);
            af[1] = __float_as_uint(p[8 * AT_LDQ]);
            af[2] = __float_as_uint(p[4]);
            af[3] = __float_as_uint(p[8 * AT_LDQ + 4]);
            const float* q = Kb + (wn * 8 + g) * AT_LDQ + ks * 8 + t;
            uint32_t bf[2] = {__float_as_uint(q[0]), __float_as_uint(q[4])};
            mma_tf32(acc, af, bf);
        }
        {
            const int r = wm * 16 + g, col = kc * AT_KR + wn * 8 + 2 * t;
            Ss[r * SLD + col] = acc[0] * a.scale;
            Ss[r * SLD + col + 1] = acc[1] * a.scale;
            Ss[(r + 8) * SLD + col] = acc[2] * a.scale;
            Ss[(r + 8) * SLD + col + 1] = acc[3] * a.scale;
        }
        __syncthreads();
    }
    // ---- O = P V, 128 value columns at a time: warp (wm, wn) -> 16 queries x 32 columns.  The 8 x nkc V chunks
    //      (32 keys x 128 columns) form one stream through TWO buffers: chunk s+1 is in flight (cp.async) while the
    //      MMAs of chunk s run, and chunk 0 is fetched behind the softmax (every round used to expose an L2 round trip).
    const int b = bh / a.H, h = bh - b * a.H;
    const int nst = 8 * nkc;
    auto issue_v = [&](int s_) {
        const int nc_ = s_ / nkc, kc_ = s_ - nc_ * nkc;
        float* dstb = KV + (s_ % NB) * (AT_KR * AT_LDV);
        for (int i = tid; i < AT_KR * 32; i += NT) {
            const int r = i >> 5, c4 = i & 31;
            const int key = kc_ * AT_KR + r;
            const bool valid = key < Tc;
            cp_async16(dstb + r * AT_LDV + c4 * 4, Vg + (long long)(valid ? key : 0) * 1024 + nc_ * 128 + c4 * 4, valid);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s_ = 0; s_ < NB - 1; ++s_) {
        if (s_ < nst) issue_v(s_);
        else cp_async_commit();
    }
    // ---- softmax over keys (rows of Ss); padded keys get probability 0
    for (int r = warp * 4; r < warp * 4 + 4; ++r) {
        float* row = Ss + r * SLD;
        float m = -INFINITY;
        for (int j = lane; j < Tc; j += 32) m = fmaxf(m, row[j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < Tc; j += 32) {
            const float e = __expf(row[j] - m);
            row[j] = e;
            s += e;
        }
        s = warp_sum(s);
        const float inv = 1.f / s;
        for (int j = lane; j < a.tk_pad; j += 32) row[j] = j < Tc ? tf32r(row[j] * inv) : 0.f;
    }
    float acc[4][4];
    int st = 0;
    for (int nc = 0; nc < 8; ++nc) {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        for (int kc = 0; kc < nkc; ++kc, ++st) {
            if (st + NB - 1 < nst) issue_v(st + NB - 1);  // into the buffer chunk st-1 was read from (all warps passed the barrier below)
            else cp_async_commit();                       // (empty group: keeps the wait count uniform)
            cp_async_wait<NB - 1>();                      // chunk st has landed (this thread's pieces)
            __syncthreads();          // everyone's pieces of chunk st (and, the first time, the softmax rows)
            const float* Vb = KV + (st % NB) * (AT_KR * AT_LDV);
#pragma unroll
            for (int ks = 0; ks < AT_KR / 8; ++ks) {
                uint32_t af[4];
                const float* p = Ss + (wm * 16 + g) * SLD + kc * AT_KR + ks * 8 + t;
                af[0] = __float_as_uint(p[0]);
                af[1] = __float_as_uint(p[8 * SLD]);
                af[2] = __float_as_uint(p[4]);
                af[3] = __float_as_uint(p[8 * SLD + 4]);
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const float* q = Vb + (ks * 8 + t) * AT_LDV + wn * 32 + ni * 8 + g;
                    uint32_t bf[2] = {__float_as_uint(q[0]), __float_as_uint(q[4 * AT_LDV])};
                    mma_tf32(acc[ni], af, bf);
                }
            }
            __syncthreads();          // chunk st's buffer may be refilled by the next iteration's issue
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int n = nc * 128 + wn * 32 + ni * 8 + 2 * t;  // f*16 + j
            const int f = n >> 4, j = n & 15;
            const int r = q0 + wm * 16 + g;
            if (r < Tc)
                *reinterpret_cast<float2*>(a.o + (((long long)b * Tc + r) * 64 + f) * 64 + h * 16 + j) = make_float2(acc[ni][0], acc[ni][1]);
            if (r + 8 < Tc)
                *reinterpret_cast<float2*>(a.o + (((long long)b * Tc + r + 8) * 64 + f) * 64 + h * 16 + j) = make_float2(acc[ni][2], acc[ni][3]);
        }
    }
}

}  // namespace rtfs
