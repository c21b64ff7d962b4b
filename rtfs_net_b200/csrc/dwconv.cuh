// Depthwise 4x4 convolutions of the RTFS block on channels-last (B,T,F,64) tensors
// (reference: ConvNormAct with groups=C, layers/conv_layers.py:65-129; used by the down-samplers
// tdanet.py:61-76 and by the three TF-AR units layers/fusion.py:25-52).
//   out[to][fo][c] = bias[c] + sum_{i,j<4} w[c][i][j] * X[S*to-1+i][S*fo-1+j][c]     (zero outside)
// stride 1 'same' (pad 1 before / 2 after) and stride 2 pad 1 share this index form.
// X is produced on the fly by an input functor (fused gLN-apply / PReLU / TF-AR combine), so the
// normalised tensors are never materialised.  One warp owns a strip of FS output columns x TSEG
// output rows over all 64 channels (lane = channel pair), with a rolling 4-row register window:
// every input element is fetched once per strip (L1 absorbs the 3-column halo between strips).
// Epilogue: per-sample (sum, sumsq) of each output for the following gLN, fp64 atomics.
#pragma once
#include "common.cuh"

namespace rtfs {

// ---------------------------------------------------------------- input functors
// contract: init(b) ; operator()(ti, fi) -> channels (2*lane, 2*lane+1) at an in-range coordinate
struct XfPlain {
    const float* x;  // [B][Ti][Fi][64]
    int Ti, Fi;
    const float* base_;
    DEVINL void init(int b) { base_ = x + (long long)b * Ti * Fi * 64 + 2 * (threadIdx.x & 31); }
    DEVINL float2 operator()(int ti, int fi) const { return ldg2(base_ + ((long long)ti * Fi + fi) * 64); }
};

// X = act(gLN(x)) ; ACT 0 none / 2 PReLU(slope)
template <int ACT>
struct XfGln {
    const float* x;
    int Ti, Fi;
    GlnRef gln;
    const float* slope;
    const float* base_;
    float2 sc_, sh_;
    float a_;
    DEVINL void init(int b) {
        const int c = 2 * (threadIdx.x & 31);
        base_ = x + (long long)b * Ti * Fi * 64 + c;
        float mean, rstd;
        gln_mean_rstd(gln.sums, b, gln.inv_n, mean, rstd);
        const float2 g = ldg2(gln.gamma + c), be = ldg2(gln.beta + c);
        sc_ = make_float2(rstd * g.x, rstd * g.y);
        sh_ = make_float2(be.x - mean * sc_.x, be.y - mean * sc_.y);
        a_ = (ACT == 2) ? __ldg(slope) : 0.f;
    }
    DEVINL float2 operator()(int ti, int fi) const {
        const float2 v = ldg2(base_ + ((long long)ti * Fi + fi) * 64);
        float2 y = make_float2(fmaf(v.x, sc_.x, sh_.x), fmaf(v.y, sc_.y, sh_.y));
        if (ACT == 2) {
            y.x = prelu(y.x, a_);
            y.y = prelu(y.y, a_);
        }
        return y;
    }
};

// X = TF-AR output (layers/fusion.py:54-69) computed on the fly:
//   gLN_l(l)[ti][fi] * sigmoid(gLN_g(g))[near(ti)][near(fi)] + gLN_e(e)[near(ti)][near(fi)]
struct XfTfar {
    const float* l;  // [B][Ti][Fi][64]
    const float* g;  // [B][Tg][Fg][64]  gate pre-norm
    const float* e;  // [B][Tg][Fg][64]  embedding pre-norm
    int Ti, Fi, Tg, Fg;
    GlnRef nl, ng, ne;
    const float *bl_, *bg_, *be_;
    float2 scl_, shl_, scg_, shg_, sce_, she_;
    DEVINL void mk(const GlnRef& r, int b, int c, float2& sc, float2& sh) {
        float mean, rstd;
        gln_mean_rstd(r.sums, b, r.inv_n, mean, rstd);
        const float2 gm = ldg2(r.gamma + c), be = ldg2(r.beta + c);
        sc = make_float2(rstd * gm.x, rstd * gm.y);
        sh = make_float2(be.x - mean * sc.x, be.y - mean * sc.y);
    }
    DEVINL void init(int b) {
        const int c = 2 * (threadIdx.x & 31);
        bl_ = l + (long long)b * Ti * Fi * 64 + c;
        bg_ = g + (long long)b * Tg * Fg * 64 + c;
        be_ = e + (long long)b * Tg * Fg * 64 + c;
        mk(nl, b, c, scl_, shl_);
        mk(ng, b, c, scg_, shg_);
        mk(ne, b, c, sce_, she_);
    }
    DEVINL float2 operator()(int ti, int fi) const {
        const float2 vl = ldg2(bl_ + ((long long)ti * Fi + fi) * 64);
        const int tg = nearest_src(ti, Tg, Ti), fg = nearest_src(fi, Fg, Fi);
        const long long og = ((long long)tg * Fg + fg) * 64;
        const float2 vg = ldg2(bg_ + og), ve = ldg2(be_ + og);
        float2 y;
        y.x = fmaf(vl.x, scl_.x, shl_.x) * sigmoidf_fast(fmaf(vg.x, scg_.x, shg_.x)) + fmaf(ve.x, sce_.x, she_.x);
        y.y = fmaf(vl.y, scl_.y, shl_.y) * sigmoidf_fast(fmaf(vg.y, scg_.y, shg_.y)) + fmaf(ve.y, sce_.y, she_.y);
        return y;
    }
};

template <int NW>
struct DwArgs {
    int Ti, Fi, To, Fo, tseg;
    const float* w[NW];     // [16][64] tap-major
    const float* bias[NW];  // [64] or null
    float* out[NW];         // [B][To][Fo][64]
    double* sums[NW];       // [B][2] or null
    float* pool;            // [B][To][Fo][64] (POOL only)
};

template <int S, int FS, int NW, bool POOL, class XF>
__global__ void __launch_bounds__(256) dw4x4_kernel(XF xf, DwArgs<NW> a) {
    constexpr int WC = S * (FS - 1) + 4;
    __shared__ float scratch[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int nstrips = (a.Fo + FS - 1) / FS;
    const int nseg = (a.To + a.tseg - 1) / a.tseg;
    const int wid = blockIdx.x * 8 + warp;
    const bool active = wid < nstrips * nseg;
    const int seg = wid / nstrips, strip = wid - seg * nstrips;
    const int fo0 = strip * FS;
    const int to0 = seg * a.tseg;
    const int to1 = active ? min(to0 + a.tseg, a.To) : to0;

    float st_s[NW], st_q[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) st_s[w] = st_q[w] = 0.f;

    if (active) {
        xf.init(b);
        float2 wr[NW][16], bs[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) {
#pragma unroll
            for (int k = 0; k < 16; ++k) wr[w][k] = ldg2(a.w[w] + k * 64 + 2 * lane);
            bs[w] = a.bias[w] ? ldg2(a.bias[w] + 2 * lane) : make_float2(0.f, 0.f);
        }
        const int wT = 2 + (a.Ti & 1), wF = 2 + (a.Fi & 1);
        const float pool_scale = 1.f / (float)(wT * wF);

        float2 win[4][WC];
        auto load_row = [&](int ti, float2 (&dst)[WC]) {
#pragma unroll
            for (int q = 0; q < WC; ++q) {
                const int fi = S * fo0 - 1 + q;
                dst[q] = (ti >= 0 && ti < a.Ti && fi >= 0 && fi < a.Fi) ? xf(ti, fi) : make_float2(0.f, 0.f);
            }
        };
#pragma unroll
        for (int r = 0; r < 4; ++r) load_row(S * to0 - 1 + r, win[r]);

        for (int to = to0; to < to1; ++to) {
#pragma unroll
            for (int o = 0; o < FS; ++o) {
                const int fo = fo0 + o;
                if (fo < a.Fo) {
                    const long long off = (((long long)b * a.To + to) * a.Fo + fo) * 64 + 2 * lane;
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        float2 acc = bs[w];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc.x = fmaf(wr[w][i * 4 + j].x, win[i][S * o + j].x, acc.x);
                                acc.y = fmaf(wr[w][i * 4 + j].y, win[i][S * o + j].y, acc.y);
                            }
                        *reinterpret_cast<float2*>(a.out[w] + off) = acc;
                        st_s[w] += acc.x + acc.y;
                        st_q[w] += acc.x * acc.x + acc.y * acc.y;
                    }
                    if (POOL) {
                        float2 p = make_float2(0.f, 0.f);
#pragma unroll
                        for (int i = 1; i < 4; ++i)
#pragma unroll
                            for (int j = 1; j < 4; ++j)
                                if (i <= wT && j <= wF) {
                                    p.x += win[i][S * o + j].x;
                                    p.y += win[i][S * o + j].y;
                                }
                        *reinterpret_cast<float2*>(a.pool + off) = make_float2(p.x * pool_scale, p.y * pool_scale);
                    }
                }
            }
            if (to + 1 < to1) {
                if (S == 1) {
#pragma unroll
                    for (int q = 0; q < WC; ++q) {
                        win[0][q] = win[1][q];
                        win[1][q] = win[2][q];
                        win[2][q] = win[3][q];
                    }
                    load_row(to + 3, win[3]);
                } else {
#pragma unroll
                    for (int q = 0; q < WC; ++q) {
                        win[0][q] = win[2][q];
                        win[1][q] = win[3][q];
                    }
                    load_row(2 * to + 3, win[2]);
                    load_row(2 * to + 4, win[3]);
                }
            }
        }
    }
#pragma unroll
    for (int w = 0; w < NW; ++w)
        block_stats_atomic(st_s[w], st_q[w], a.sums[w] ? a.sums[w] + 2 * b : nullptr, scratch);
}

template <int S, int FS, int NW, bool POOL, class XF>
inline cudaError_t launch_dw(const XF& xf, const DwArgs<NW>& a, int B, cudaStream_t st) {
    const int nstrips = (a.Fo + FS - 1) / FS;
    const int nseg = (a.To + a.tseg - 1) / a.tseg;
    dim3 grid((nstrips * nseg + 7) / 8, 1, B);
    dw4x4_kernel<S, FS, NW, POOL, XF><<<grid, 256, 0, st>>>(xf, a);
    return cudaGetLastError();
}

}  // namespace rtfs
