// CAF cross-dimensional attention fusion (reference: ATTNFusionCell.forward, layers/fusion.py:252-274;
// ATTNFusion.forward, TDAVNet/fusion.py:204-212).  Eval-mode BatchNorm2d is folded on the host
// into per-channel scale/shift.
//   caf_video_kernel : video (B,Cv,Tv) -> vk[b][tv][c] = gLN(Conv1d_{groups=Ca}(v))           (resize)
//                                        att[b][tv][c] = softmax_tv(mean_4(gLN(Conv1d_{groups=Ca}(v))))
//   caf_apply_kernel : out = ReLU(a*sk+tk) * vk[near(t)] + att[near(t)] * (a*sv+tv)  [+ a1]
//                      streaming over the channels-last (B,T,F,Ca) audio tensor
#pragma once
#include "common.cuh"

namespace rtfs {

struct CafVideoArgs {
    const float* v;  // (B, Cv=2*Ca, Tv)
    const float* wr;  // resize conv weight [Ca][2], bias [Ca], gLN gamma/beta [Ca]
    const float* br;
    const float* gr;
    const float* ber;
    const float* wa;  // attention conv weight [4*Ca][2], bias [4*Ca], gLN gamma/beta [4*Ca]
    const float* ba;
    const float* ga;
    const float* bea;
    float* vk;   // (B,Tv,Ca)
    float* att;  // (B,Tv,Ca)
    int Ca, Tv;
};

// one CTA per utterance, Ca (=256) threads: thread c owns audio channel c (video group c)
__global__ void __launch_bounds__(256) caf_video_kernel(CafVideoArgs a) {
    extern __shared__ float m_s[];  // [Tv][Ca] head-mean of the normalised attention logits
    __shared__ float red[4][8];
    __shared__ double st[4];
    const int c = threadIdx.x, b = blockIdx.x, Ca = a.Ca, Tv = a.Tv;
    const float* v0 = a.v + ((long long)b * 2 * Ca + 2 * c) * Tv;
    const float* v1 = v0 + Tv;
    const float wr0 = a.wr[2 * c], wr1 = a.wr[2 * c + 1], brc = a.br[c];
    float wa0[4], wa1[4], bac[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        wa0[i] = a.wa[2 * (4 * c + i)];
        wa1[i] = a.wa[2 * (4 * c + i) + 1];
        bac[i] = a.ba[4 * c + i];
    }
    // pass 1: gLN statistics of both grouped convs (over channels x Tv)
    float sr = 0.f, qr = 0.f, sa = 0.f, qa = 0.f;
    for (int t = 0; t < Tv; ++t) {
        const float x0 = __ldg(v0 + t), x1 = __ldg(v1 + t);
        const float r = fmaf(wr0, x0, fmaf(wr1, x1, brc));
        sr += r;
        qr += r * r;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float y = fmaf(wa0[i], x0, fmaf(wa1[i], x1, bac[i]));
            sa += y;
            qa += y * y;
        }
    }
    float vals[4] = {sr, qr, sa, qa};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float w = warp_sum(vals[i]);
        if ((c & 31) == 0) red[i][c >> 5] = w;
    }
    __syncthreads();
    if (c < 4) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += (double)red[c][w];
        st[c] = s;  // finalised below in double by every thread
    }
    __syncthreads();
    const double nr = (double)Ca * Tv, na = 4.0 * nr;
    const double mr = st[0] / nr, ma = st[2] / na;
    double vr = st[1] / nr - mr * mr, va = st[3] / na - ma * ma;
    vr = vr < 0 ? 0 : vr;
    va = va < 0 ? 0 : va;
    const float mean_r = (float)mr, rstd_r = (float)(1.0 / sqrt(vr + 1e-5));
    const float mean_a = (float)ma, rstd_a = (float)(1.0 / sqrt(va + 1e-5));
    const float grc = a.gr[c], berc = a.ber[c];
    float gac[4], beac[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        gac[i] = a.ga[4 * c + i];
        beac[i] = a.bea[4 * c + i];
    }
    // pass 2: normalise; resize -> vk ; attention -> mean over the 4 sub-channels -> softmax over Tv
    float mx = -INFINITY;
    for (int t = 0; t < Tv; ++t) {
        const float x0 = __ldg(v0 + t), x1 = __ldg(v1 + t);
        const float r = fmaf(wr0, x0, fmaf(wr1, x1, brc));
        a.vk[((long long)b * Tv + t) * Ca + c] = (r - mean_r) * rstd_r * grc + berc;
        float m = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float y = fmaf(wa0[i], x0, fmaf(wa1[i], x1, bac[i]));
            m += (y - mean_a) * rstd_a * gac[i] + beac[i];
        }
        m *= 0.25f;
        m_s[t * Ca + c] = m;
        mx = fmaxf(mx, m);
    }
    float den = 0.f;
    for (int t = 0; t < Tv; ++t) {
        const float e = __expf(m_s[t * Ca + c] - mx);
        m_s[t * Ca + c] = e;
        den += e;
    }
    const float inv = 1.f / den;
    for (int t = 0; t < Tv; ++t) a.att[((long long)b * Tv + t) * Ca + c] = m_s[t * Ca + c] * inv;
}

struct CafApplyArgs {
    const float* a;   // (B,T,F,C) audio
    const float* a1;  // optional addend (next block input = CAF + a1), may be null
    const float* vk;  // (B,Tv,C)
    const float* att;
    const float* sk;  // folded BN of key_embed: ReLU(a*sk + tk)
    const float* tk;
    const float* sv;  // folded BN of value_embed
    const float* tv;
    float* out;
    int T, F, C, Tv;
    long long total4;  // B*T*F*C/4
};

__global__ void __launch_bounds__(256) caf_apply_kernel(CafApplyArgs p) {
    const int c4n = p.C >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4;
        const long long pos = i / c4n;
        const long long bt = pos / p.F;
        const int t = (int)(bt % p.T);
        const int b = (int)(bt / p.T);
        const int tv = nearest_src(t, p.Tv, p.T);
        const long long vo = ((long long)b * p.Tv + tv) * p.C + c;
        const float4 x = ldg4(p.a + i * 4);
        const float4 k = ldg4(p.vk + vo), at = ldg4(p.att + vo);
        const float4 sk = ldg4(p.sk + c), tk = ldg4(p.tk + c), sv = ldg4(p.sv + c), tvv = ldg4(p.tv + c);
        float4 y;
        y.x = fmaxf(fmaf(x.x, sk.x, tk.x), 0.f) * k.x + at.x * fmaf(x.x, sv.x, tvv.x);
        y.y = fmaxf(fmaf(x.y, sk.y, tk.y), 0.f) * k.y + at.y * fmaf(x.y, sv.y, tvv.y);
        y.z = fmaxf(fmaf(x.z, sk.z, tk.z), 0.f) * k.z + at.z * fmaf(x.z, sv.z, tvv.z);
        y.w = fmaxf(fmaf(x.w, sk.w, tk.w), 0.f) * k.w + at.w * fmaf(x.w, sv.w, tvv.w);
        if (p.a1) {
            const float4 r = ldg4(p.a1 + i * 4);
            y.x += r.x;
            y.y += r.y;
            y.z += r.z;
            y.w += r.w;
        }
        *reinterpret_cast<float4*>(p.out + i * 4) = y;
    }
}

}  // namespace rtfs
