// VP (video) block of RTFS-Net as ONE kernel: TDANetBlock.forward with is2d = False (separators/tdanet.py:106-133) on the
// (B, 512, Tv) lip embedding -- gateway, projection 512 -> 64, four depthwise k=3 down-samplers with eval BatchNorm1d, pooled
// global feature, GlobalAttention (layers/attention.py:28-73,192-220: LayerNorm, sinusoidal PE, 8-head self-attention over the
// <= 16 coarsest frames, LayerNorm; FeedForwardNetwork layers/conv_layers.py:218-259), seven 1-D TF-AR units
// (layers/fusion.py:54-69) and the residual 1x1 conv 64 -> 512.
//
// The tensors are tiny (100 KB in, 100 KB out, 3.4 MMAC per utterance), so the round-1 path -- ~150 eager torch launches replayed
// from a CUDA graph -- was pure launch / dependency latency.  Here one CTA owns one utterance and every intermediate stays in
// shared memory ([channel][frame] rows); all arithmetic is fp32 FMA (no tensor cores: K <= 512 on 50 columns).  Inference only
// (running-statistics BatchNorm folded on the host, no dropout); training keeps the torch modules for autograd.
#pragma once
#include "common.cuh"

namespace rtfs {

// packed parameter buffer (weights.py: pack_video), float offsets in this order
enum vp_field {
    VP_GW_W = 0, VP_GW_B, VP_GW_A,
    VP_PJ_WT, VP_PJ_S, VP_PJ_T, VP_PJ_A,
    VP_DS_W, VP_DS_S, VP_DS_T,
    VP_LN1_G, VP_LN1_B, VP_PE, VP_IN_WT, VP_IN_B, VP_OUT_WT, VP_OUT_B, VP_LN2_G, VP_LN2_B,
    VP_F1_WT, VP_F1_G, VP_F1_B, VP_FR_W, VP_FR_B, VP_F2_WT, VP_F2_G, VP_F2_B,
    VP_TF, VP_RC_WT, VP_RC_B,
    VP_COUNT
};
constexpr int VP_C = 512, VP_N = 64, VP_DEPTH = 4, VP_HEADS = 8, VP_HID = 128, VP_MAXTOK = 16, VP_TF_UNIT = 960;
constexpr int vp_sizes[VP_COUNT] = {
    VP_C, VP_C, 4,
    VP_C * VP_N, VP_N, VP_N, 4,
    VP_DEPTH * 3 * VP_N, VP_DEPTH * VP_N, VP_DEPTH * VP_N,
    VP_N, VP_N, VP_MAXTOK * VP_N, VP_N * 3 * VP_N, 3 * VP_N, VP_N * VP_N, VP_N, VP_N, VP_N,
    VP_N * VP_HID, VP_HID, VP_HID, 3 * VP_HID, VP_HID, VP_HID * VP_N, VP_N, VP_N,
    7 * VP_TF_UNIT, VP_N * VP_C, VP_C};

struct VpOffsets {
    int o[VP_COUNT];
    int total;
};
inline VpOffsets vp_offsets() {
    VpOffsets r;
    int acc = 0;
    for (int i = 0; i < VP_COUNT; ++i) {
        r.o[i] = acc;
        acc += (vp_sizes[i] + 3) / 4 * 4;
    }
    r.total = acc;
    return r;
}

struct VideoArgs {
    const float* x;   // (B, 512, Tv)
    const float* w;   // packed parameters
    float* out;       // (B, 512, Tv)
    VpOffsets off;
    int Tv;
    int len[VP_DEPTH];   // frames per scale
    int start[VP_DEPTH]; // column offset of scale i inside the [64][sum len] arrays
    int sumlen;
};

constexpr int VP_GS = 8 * VP_HID * VP_MAXTOK;  // global-attention scratch: 8 arrays of [128][16] floats (later the output staging tile)
inline int vp_smem_floats(int Tv, int sumlen) {
    // ds [64][sumlen], fu [64][sumlen], y [64][Tv], e [64][Tv + 2], stage [32][Tv] + 64 (x chunk / scale-1 expansion), scratch
    return 2 * VP_N * sumlen + VP_N * Tv + VP_N * (Tv + 2) + 32 * Tv + 64 + VP_GS + 64;
}

DEVINL float vp_block_sum(float v, float* red) {  // sum over the 256 threads, result to all
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    return s;
}

// depthwise k=3 conv (+ folded BatchNorm) of a [64][Li] array into [64][Lo]: stride 1 ('same': pad 1) or 2 (pad 1)
DEVINL void vp_dwconv(const float* in, int Li, int ldi, float* outp, int Lo, int ldo, int stride, const float* w3, const float* s, const float* t, int mode) {
    // mode 0: out = bn(conv) ; 1: out = sigmoid(bn(conv))
    for (int idx = threadIdx.x; idx < VP_N * Lo; idx += blockDim.x) {
        const int c = idx / Lo, to = idx - c * Lo;
        const int t0 = to * stride - 1;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int ti = t0 + k;
            if (ti >= 0 && ti < Li) acc = fmaf(in[c * ldi + ti], __ldg(w3 + k * VP_N + c), acc);
        }
        float v = fmaf(acc, __ldg(s + c), __ldg(t + c));
        if (mode == 1) v = 1.f / (1.f + expf(-v));
        outp[c * ldo + to] = v;
    }
}

// Gateway (dw 1x1 + PReLU) + projection 512 -> 64 (+ folded BN + PReLU): thread = (output channel co, frame phase tq), frames
// tq, tq + 4, ... , NI of them.  The FMA block is unrolled over the 32 staged input channels x NI frames with every shared-memory
// address = running row pointer + immediate and no per-element predicate (a frame index past Tv reads the next row / the 64-float
// pad behind the staging area and lands in an accumulator that is never stored).  It used to be one 32 x 32 block with a predicate
// and a multiply per address: 13.5 k instructions executed 16 times per warp -- 60 % of the kernel's samples, 35-55 % of them
// instruction-fetch stalls (ncu: `no_inst`), the body did not fit the instruction cache.
template <int NI>
DEVINL void vp_project(const VideoArgs& a, const float* xb, float* stage, float* ybuf) {
    const int tid = threadIdx.x, Tv = a.Tv;
    const float* W = a.w;
    const int* O = a.off.o;
    const int co = tid & 63, tq = tid >> 6;
    float acc[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[i] = 0.f;
    const float ga = __ldg(W + O[VP_GW_A]);
    for (int c0 = 0; c0 < VP_C; c0 += 32) {
        // the chunk's 32 weights of this thread's output channel are requested before the staging barrier: 32 independent loads
        // in flight instead of one L2 round trip per input channel
        float wv[32];
#pragma unroll
        for (int cc = 0; cc < 32; ++cc) wv[cc] = __ldg(W + O[VP_PJ_WT] + (c0 + cc) * VP_N + co);
        __syncthreads();
        for (int idx = tid; idx < 32 * Tv; idx += 256) {
            const int cc = idx / Tv, t = idx - cc * Tv;
            const float v = fmaf(__ldg(W + O[VP_GW_W] + c0 + cc), __ldg(xb + (c0 + cc) * Tv + t), __ldg(W + O[VP_GW_B] + c0 + cc));
            stage[cc * Tv + t] = prelu(v, ga);
        }
        __syncthreads();
        const float* r = stage + tq;
#pragma unroll
        for (int cc = 0; cc < 32; ++cc) {
            const float w = wv[cc];
#pragma unroll
            for (int i = 0; i < NI; ++i) acc[i] = fmaf(w, r[4 * i], acc[i]);
            r += Tv;
        }
    }
    const float s = __ldg(W + O[VP_PJ_S] + co), t0 = __ldg(W + O[VP_PJ_T] + co), pa = __ldg(W + O[VP_PJ_A]);
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int t = tq + 4 * i;
        if (t < Tv) ybuf[co * Tv + t] = prelu(fmaf(acc[i], s, t0), pa);
    }
}

__global__ void __launch_bounds__(256) video_block_kernel(VideoArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, b = blockIdx.x, Tv = a.Tv, SL = a.sumlen;
    const float* W = a.w;
    const int* O = a.off.o;
    float* ds = sm;                      // [64][SL] down-sampled scales
    float* fu = ds + VP_N * SL;          // [64][SL] TF-AR outputs per scale
    float* ybuf = fu + VP_N * SL;        // [64][Tv] projection output, later the expanded feature
    float* ebuf = ybuf + VP_N * Tv;      // [64][Tv] scratch (TF-AR global branches, expansion)
    float* stage = ebuf + VP_N * (Tv + 2);  // [32][Tv] + 64: x chunk of the projection, later the scale-1 expansion [64][ceil(Tv/2)]
    float* gs = stage + 32 * Tv + 64;    // global-attention scratch: 8 arrays of [128][16]; at the end the [512][17] output staging tile
    float* red = gs + VP_GS;
    const float* xb = a.x + (long long)b * VP_C * Tv;
    const int Tg = a.len[VP_DEPTH - 1];

    // ---- gateway (dw 1x1 + PReLU) fused into the projection 512 -> 64 (+ BN + PReLU)        tdanet.py:34-49,107-113
    {
        const int ni = (Tv + 3) >> 2;  // frames per thread; the unrolled FMA block is instantiated for a few bounds (vp_project)
        if (ni <= 4) vp_project<4>(a, xb, stage, ybuf);
        else if (ni <= 7) vp_project<7>(a, xb, stage, ybuf);
        else if (ni <= 10) vp_project<10>(a, xb, stage, ybuf);
        else if (ni <= 13) vp_project<13>(a, xb, stage, ybuf);
        else if (ni <= 19) vp_project<19>(a, xb, stage, ybuf);
        else vp_project<25>(a, xb, stage, ybuf);
    }
    __syncthreads();
    // ---- down-samplers: ds[0] = bn(dw(y)), ds[i] = bn(dw_s2(ds[i-1]))                         tdanet.py:61-76,113-115
    vp_dwconv(ybuf, Tv, Tv, ds + a.start[0], a.len[0], SL, 1, W + O[VP_DS_W], W + O[VP_DS_S], W + O[VP_DS_T], 0);
    __syncthreads();
    for (int i = 1; i < VP_DEPTH; ++i) {
        vp_dwconv(ds + a.start[i - 1], a.len[i - 1], SL, ds + a.start[i], a.len[i], SL, 2, W + O[VP_DS_W] + i * 3 * VP_N, W + O[VP_DS_S] + i * VP_N,
                  W + O[VP_DS_T] + i * VP_N, 0);
        __syncthreads();
    }
    // ---- g = sum_i adaptive_avg_pool1d(ds[i], Tg)   -> gs[0] as [64][Tg]                      tdanet.py:117-118
    float* g = gs;                        // [64][16]
    for (int idx = tid; idx < VP_N * Tg; idx += 256) {
        const int c = idx / Tg, j = idx - c * Tg;
        float acc = 0.f;
        for (int i = 0; i < VP_DEPTH; ++i) {
            const int Li = a.len[i];
            const int s0 = (j * Li) / Tg, s1 = ((j + 1) * Li + Tg - 1) / Tg;
            float m = 0.f;
            for (int t = s0; t < s1; ++t) m += ds[c * SL + a.start[i] + t];
            acc += m / (float)(s1 - s0);
        }
        g[c * VP_MAXTOK + j] = acc;
    }
    __syncthreads();
    // ---- GlobalAttention.MHSA                                                               attention.py:57-73
    float* yt = gs + 1 * VP_HID * VP_MAXTOK;   // [Tg][64] tokens after LN1 + PE (also the residual)
    float* qkv = gs + 2 * VP_HID * VP_MAXTOK;  // [Tg][192]
    float* ob = gs + 4 * VP_HID * VP_MAXTOK;   // [Tg][64] attention output / projections
    float* att = gs + 5 * VP_HID * VP_MAXTOK;  // [8][Tg][Tg]
    {  // LayerNorm over the 64 channels of a token: 16 lanes per token (every thread runs the shuffles; tokens >= Tg are not stored)
        const int tok = tid >> 4, l = tid & 15;
        const bool live = tok < Tg;
        float v[4], s = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = live ? g[(l * 4 + k) * VP_MAXTOK + tok] : 0.f;
            s += v[k];
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
        const float mu = s * (1.f / 64.f);
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) q += (v[k] - mu) * (v[k] - mu);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o, 16);
        const float rs = 1.f / sqrtf(q * (1.f / 64.f) + 1e-5f);
        if (live) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = l * 4 + k;
                yt[tok * 64 + c] = (v[k] - mu) * rs * __ldg(W + O[VP_LN1_G] + c) + __ldg(W + O[VP_LN1_B] + c) + __ldg(W + O[VP_PE] + tok * 64 + c);
            }
        }
    }
    __syncthreads();
    for (int idx = tid; idx < Tg * 192; idx += 256) {  // in_proj
        const int tok = idx / 192, o = idx - tok * 192;
        float acc = __ldg(W + O[VP_IN_B] + o);
#pragma unroll 16
        for (int c = 0; c < 64; ++c) acc = fmaf(yt[tok * 64 + c], __ldg(W + O[VP_IN_WT] + c * 192 + o), acc);
        qkv[tok * 192 + o] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < VP_HEADS * Tg * Tg; idx += 256) {  // scores
        const int h = idx / (Tg * Tg), r = idx - h * Tg * Tg, tq = r / Tg, tk = r - tq * Tg;
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) acc = fmaf(qkv[tq * 192 + h * 8 + d], qkv[tk * 192 + 64 + h * 8 + d], acc);
        att[idx] = acc * 0.35355339059327373f;  // 1/sqrt(8)
    }
    __syncthreads();
    for (int idx = tid; idx < VP_HEADS * Tg; idx += 256) {  // softmax rows
        float* row = att + idx * Tg;
        float mx = -INFINITY;
        for (int k = 0; k < Tg; ++k) mx = fmaxf(mx, row[k]);
        float den = 0.f;
        for (int k = 0; k < Tg; ++k) {
            row[k] = expf(row[k] - mx);
            den += row[k];
        }
        const float inv = 1.f / den;
        for (int k = 0; k < Tg; ++k) row[k] *= inv;
    }
    __syncthreads();
    for (int idx = tid; idx < Tg * 64; idx += 256) {  // context
        const int tok = idx >> 6, c = idx & 63, h = c >> 3;
        float acc = 0.f;
        for (int k = 0; k < Tg; ++k) acc = fmaf(att[(h * Tg + tok) * Tg + k], qkv[k * 192 + 128 + c], acc);
        ob[idx] = acc;
    }
    __syncthreads();
    float* o2 = gs + 6 * VP_HID * VP_MAXTOK;  // [Tg][64] out_proj + residual
    for (int idx = tid; idx < Tg * 64; idx += 256) {
        const int tok = idx >> 6, o = idx & 63;
        float acc = __ldg(W + O[VP_OUT_B] + o);
#pragma unroll 16
        for (int c = 0; c < 64; ++c) acc = fmaf(ob[tok * 64 + c], __ldg(W + O[VP_OUT_WT] + c * 64 + o), acc);
        o2[idx] = acc + yt[idx];
    }
    __syncthreads();
    {  // LayerNorm 2, transpose back, + g
        const int tok = tid >> 4, l = tid & 15;
        const bool live = tok < Tg;
        float v[4], s = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = live ? o2[tok * 64 + l * 4 + k] : 0.f;
            s += v[k];
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
        const float mu = s * (1.f / 64.f);
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) q += (v[k] - mu) * (v[k] - mu);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o, 16);
        const float rs = 1.f / sqrtf(q * (1.f / 64.f) + 1e-5f);
        if (live) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = l * 4 + k;
                g[c * VP_MAXTOK + tok] += (v[k] - mu) * rs * __ldg(W + O[VP_LN2_G] + c) + __ldg(W + O[VP_LN2_B] + c);
            }
        }
    }
    __syncthreads();
    // ---- GlobalAttention.FFN: 1x1 64 -> 128 + gLN ; dw k=3 + bias + ReLU ; 1x1 128 -> 64 + gLN ; + g     conv_layers.py:252-259
    float* h1 = gs + 1 * VP_HID * VP_MAXTOK;  // [128][16]
    float* h2 = gs + 2 * VP_HID * VP_MAXTOK;
    float* h3 = gs + 3 * VP_HID * VP_MAXTOK;  // [64][16]
    float part = 0.f, part2 = 0.f;
    for (int idx = tid; idx < VP_HID * Tg; idx += 256) {
        const int o = idx / Tg, t = idx - o * Tg;
        float acc = 0.f;
#pragma unroll 16
        for (int c = 0; c < 64; ++c) acc = fmaf(g[c * VP_MAXTOK + t], __ldg(W + O[VP_F1_WT] + c * VP_HID + o), acc);
        h1[o * VP_MAXTOK + t] = acc;
        part += acc;
    }
    {
        const float n = (float)(VP_HID * Tg);
        const float mu = vp_block_sum(part, red) / n;
        for (int idx = tid; idx < VP_HID * Tg; idx += 256) {
            const float d = h1[(idx / Tg) * VP_MAXTOK + idx % Tg] - mu;
            part2 += d * d;
        }
        const float rs = 1.f / sqrtf(vp_block_sum(part2, red) / n + 1e-5f);
        for (int idx = tid; idx < VP_HID * Tg; idx += 256) {
            const int o = idx / Tg, t = idx - o * Tg;
            h1[o * VP_MAXTOK + t] = (h1[o * VP_MAXTOK + t] - mu) * rs * __ldg(W + O[VP_F1_G] + o) + __ldg(W + O[VP_F1_B] + o);
        }
    }
    __syncthreads();
    for (int idx = tid; idx < VP_HID * Tg; idx += 256) {
        const int o = idx / Tg, t = idx - o * Tg;
        float acc = __ldg(W + O[VP_FR_B] + o);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int ti = t - 1 + k;
            if (ti >= 0 && ti < Tg) acc = fmaf(h1[o * VP_MAXTOK + ti], __ldg(W + O[VP_FR_W] + k * VP_HID + o), acc);
        }
        h2[o * VP_MAXTOK + t] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    part = 0.f;
    part2 = 0.f;
    for (int idx = tid; idx < VP_N * Tg; idx += 256) {
        const int o = idx / Tg, t = idx - o * Tg;
        float acc = 0.f;
#pragma unroll 16
        for (int c = 0; c < VP_HID; ++c) acc = fmaf(h2[c * VP_MAXTOK + t], __ldg(W + O[VP_F2_WT] + c * VP_N + o), acc);
        h3[o * VP_MAXTOK + t] = acc;
        part += acc;
    }
    {
        const float n = (float)(VP_N * Tg);
        const float mu = vp_block_sum(part, red) / n;
        for (int idx = tid; idx < VP_N * Tg; idx += 256) {
            const float d = h3[(idx / Tg) * VP_MAXTOK + idx % Tg] - mu;
            part2 += d * d;
        }
        const float rs = 1.f / sqrtf(vp_block_sum(part2, red) / n + 1e-5f);
        for (int idx = tid; idx < VP_N * Tg; idx += 256) {
            const int o = idx / Tg, t = idx - o * Tg;
            g[o * VP_MAXTOK + t] += (h3[o * VP_MAXTOK + t] - mu) * rs * __ldg(W + O[VP_F2_G] + o) + __ldg(W + O[VP_F2_B] + o);
        }
    }
    __syncthreads();
    // ---- TF-AR units                                                                        tdanet.py:124-129, fusion.py:54-69
    // unit u: out[c][t] = bn(dw(local))[t] * sigmoid(bn(dw(glob)))[near(t)] + bn(dw(glob))[near(t)]
    auto tfar = [&](int u, const float* local, int Ll, int ldl, const float* glob, int Lg, int ldg_, float* outp, int ldo, const float* addend, int lda) {
        const float* P = W + O[VP_TF] + u * VP_TF_UNIT;
        float* ge = ebuf;                 // [64][Lg]
        float* gg = ebuf + VP_N * Lg;     // [64][Lg]   (2 * Lg <= Tv for every unit: Lg <= Tv / 2 except the equal-size case Lg = Tg)
        vp_dwconv(glob, Lg, ldg_, ge, Lg, Lg, 1, P + 320, P + 512, P + 576, 0);
        vp_dwconv(glob, Lg, ldg_, gg, Lg, Lg, 1, P + 640, P + 832, P + 896, 1);
        __syncthreads();
        for (int idx = tid; idx < VP_N * Ll; idx += 256) {
            const int c = idx / Ll, t = idx - c * Ll;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int ti = t - 1 + k;
                if (ti >= 0 && ti < Ll) acc = fmaf(local[c * ldl + ti], __ldg(P + k * VP_N + c), acc);
            }
            const float le = fmaf(acc, __ldg(P + 192 + c), __ldg(P + 256 + c));
            const int tg = nearest_src32(t, Lg, Ll);
            float v = le * gg[c * Lg + tg] + ge[c * Lg + tg];
            if (addend != nullptr) v += addend[c * lda + t];
            outp[c * ldo + t] = v;
        }
        __syncthreads();
    };
    for (int i = 0; i < VP_DEPTH; ++i) tfar(i, ds + a.start[i], a.len[i], SL, g, Tg, VP_MAXTOK, fu + a.start[i], SL, nullptr, 0);
    // expanded = TFAR_cat[depth-2](fused[depth-2], fused[depth-1]) + ds[depth-2]; then down to scale 0
    {
        const float* cur = fu + a.start[VP_DEPTH - 1];
        int curL = a.len[VP_DEPTH - 1], curld = SL;
        for (int i = VP_DEPTH - 2; i >= 0; --i) {
            float* dst = (i & 1) ? stage : ybuf;  // scale 1 (<= Tv/2 frames, 64 rows) fits the [32][Tv] staging area; scales 2, 0 use ybuf
            tfar(VP_DEPTH + i, fu + a.start[i], a.len[i], SL, cur, curL, curld, dst, a.len[i], ds + a.start[i], SL);
            cur = dst;
            curL = a.len[i];
            curld = a.len[i];
        }
    }
    // ---- residual conv 64 -> 512 + bias + gateway(x)                                          tdanet.py:131
    {
        const float ga = __ldg(W + O[VP_GW_A]);
        float* tile = gs;  // [512][17] output staging (the attention scratch is dead): global reads / writes then run along t
        for (int t0 = 0; t0 < Tv; t0 += 16) {
            float acc0[16], acc1[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc0[i] = acc1[i] = 0.f;
            for (int k0 = 0; k0 < VP_N; k0 += 16) {
                float w0[16], w1[16];  // 32 independent weight loads per 512 FMAs
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    w0[k] = __ldg(W + O[VP_RC_WT] + (k0 + k) * VP_C + tid);
                    w1[k] = __ldg(W + O[VP_RC_WT] + (k0 + k) * VP_C + tid + 256);
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const float* e = ybuf + (k0 + k) * Tv + t0;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float ev = (t0 + i < Tv) ? e[i] : 0.f;
                        acc0[i] = fmaf(w0[k], ev, acc0[i]);
                        acc1[i] = fmaf(w1[k], ev, acc1[i]);
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                tile[tid * 17 + i] = acc0[i];
                tile[(tid + 256) * 17 + i] = acc1[i];
            }
            __syncthreads();
            for (int idx = tid; idx < VP_C * 16; idx += 256) {
                const int co = idx >> 4, i = idx & 15, t = t0 + i;
                if (t < Tv) {
                    const float xv = __ldg(xb + co * Tv + t);
                    const float r = prelu(fmaf(__ldg(W + O[VP_GW_W] + co), xv, __ldg(W + O[VP_GW_B] + co)), ga);
                    a.out[((long long)b * VP_C + co) * Tv + t] = tile[co * 17 + i] + __ldg(W + O[VP_RC_B] + co) + r;
                }
            }
        }
    }
}

}  // namespace rtfs
