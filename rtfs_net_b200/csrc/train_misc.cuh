// Backward / training-step kernels, part 5: S^3 mask head, CAF fusion with batch-statistics BatchNorm, iSTFT adjoint,
// on-device SNR loss, gradient-norm clipping + AdamW.
//   mask:    MaskGenerator.__apply_masks (TDAVNet/mask_generator.py:67-82)
//   CAF:     ATTNFusionCell.forward (layers/fusion.py:252-274); key_embed / value_embed are depthwise 1x1 conv (no bias) +
//            BatchNorm2d, which in train() normalises with the statistics of the batch (and of all ranks under
//            sync_batchnorm, train.py:145): per channel y = w*a, so mean_y = w*mean_a, var_y = w^2*var_a and the whole
//            layer is a per-channel scale/shift (sk, tk) that the host forms from the channel sums computed here
//   iSTFT:   STFTDecoder.forward (TDAVNet/decoder.py:122-128)
//   loss:    PairwiseNegSDR("snr") for n_src = 1 under PITLossWrapper (src/losses/matrix.py:22-53, train.py:99)
//   AdamW:   torch.optim.AdamW semantics (src/system/optimizers.py:58-75) + clip_grad_norm_(5.0) (train.py:143)
#pragma once
#include "common.cuh"

namespace rtfs {

// ------------------------------------------------------------------------------------------------------------- mask
// z_r = e_r m_r - e_i m_i ; z_i = e_r m_i + e_i m_r   (e = encoder output a0, m = ReLU(conv) mask; halves 0..127 | 128..255)
//   dm_r = dz_r e_r + dz_i e_i ; dm_i = -dz_r e_i + dz_i e_r   (then * [m > 0] for the ReLU)
//   de_r = dz_r m_r + dz_i m_i ; de_i = -dz_r m_i + dz_i m_r
__global__ void __launch_bounds__(256) mask_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ a0, const float* __restrict__ m,
                                                       float* __restrict__ dm, float* __restrict__ da0, long long rows) {
    const long long total = rows * 32;  // 32 float4 per half row
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx >> 5;
        const int c = (int)(idx & 31) * 4;
        const long long o = row * 256 + c;
        const float4 zr = ldg4(dz + o), zi = ldg4(dz + o + 128), er = ldg4(a0 + o), ei = ldg4(a0 + o + 128), mr = ldg4(m + o), mi = ldg4(m + o + 128);
        float4 dmr, dmi, der, dei;
#define RTFS_MB(k)                                              \
    dmr.k = mr.k > 0.f ? zr.k * er.k + zi.k * ei.k : 0.f;       \
    dmi.k = mi.k > 0.f ? -zr.k * ei.k + zi.k * er.k : 0.f;      \
    der.k = zr.k * mr.k + zi.k * mi.k;                          \
    dei.k = -zr.k * mi.k + zi.k * mr.k;
        RTFS_MB(x) RTFS_MB(y) RTFS_MB(z) RTFS_MB(w)
#undef RTFS_MB
        *reinterpret_cast<float4*>(dm + o) = dmr;
        *reinterpret_cast<float4*>(dm + o + 128) = dmi;
        *reinterpret_cast<float4*>(da0 + o) = der;
        *reinterpret_cast<float4*>(da0 + o + 128) = dei;
    }
}

// -------------------------------------------------------------------------------------------------------------- CAF
// per-channel (sum, sum of squares) of a (rows, 256) tensor -> fp64 sums[256][2] (accumulated)
__global__ void __launch_bounds__(256) chan_stats_kernel(const float* __restrict__ x, long long rows, double* sums) {
    __shared__ float sh[512];
    sh[threadIdx.x] = 0.f;
    sh[threadIdx.x + 256] = 0.f;
    __syncthreads();
    const int q = threadIdx.x & 63, rl = threadIdx.x >> 6;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = s;
    for (long long r = (long long)blockIdx.x * 4 + rl; r < rows; r += (long long)gridDim.x * 4) {
        const float4 v = ldg4(x + r * 256 + q * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        ss.x = fmaf(v.x, v.x, ss.x); ss.y = fmaf(v.y, v.y, ss.y); ss.z = fmaf(v.z, v.z, ss.z); ss.w = fmaf(v.w, v.w, ss.w);
    }
    atomicAdd(sh + q * 4, s.x); atomicAdd(sh + q * 4 + 1, s.y); atomicAdd(sh + q * 4 + 2, s.z); atomicAdd(sh + q * 4 + 3, s.w);
    atomicAdd(sh + 256 + q * 4, ss.x); atomicAdd(sh + 256 + q * 4 + 1, ss.y); atomicAdd(sh + 256 + q * 4 + 2, ss.z); atomicAdd(sh + 256 + q * 4 + 3, ss.w);
    __syncthreads();
    atomicAdd(sums + 2 * threadIdx.x, (double)sh[threadIdx.x]);
    atomicAdd(sums + 2 * threadIdx.x + 1, (double)sh[256 + threadIdx.x]);
}

// out = ReLU(a*sk+tk) * vk[up] + att[up] * (a*sv+tv)  -- backward, pass 1 (reductions).
//   g_k = dout * vk[up] * [a*sk+tk > 0] ; g_v = dout * att[up]
//   csum[c] += (sum g_k, sum g_k*a, sum g_v, sum g_v*a)          (BatchNorm backward needs the batch means of g and g*xhat)
//   dvk[b][tv][c] += sum_{t -> tv, f} dout * ReLU(a*sk+tk) ; datt[b][tv][c] += sum dout * (a*sv+tv)
struct CafBwdArgs {
    const float* dout;  // (B,T,F,256)
    const float* a;     // (B,T,F,256) CAF audio input
    const float* vk;    // (B,Tv,256)
    const float* att;
    const float* sk;
    const float* tk;
    const float* sv;
    const float* tv;
    double* csum;       // [256][4]
    float* dvk;         // (B,Tv,256) zeroed by the caller
    float* datt;
    const float* mu;    // pass 2: batch mean of a per channel
    const float* c0;    // pass 2: sk*m1k + sv*m1v
    const float* c1;    // pass 2: sk*m2k*wk/sigk + sv*m2v*wv/sigv
    float* da;          // pass 2 output (B,T,F,256)
    int T, F, Tv, t_per_cta;
};

__global__ void __launch_bounds__(256) caf_bwd_reduce_kernel(CafBwdArgs p) {
    __shared__ float sh[4 * 256];
    for (int i = threadIdx.x; i < 1024; i += 256) sh[i] = 0.f;
    __syncthreads();
    const int q = threadIdx.x & 63, fl = threadIdx.x >> 6, c = q * 4;
    const int b = blockIdx.y;
    const int t_begin = blockIdx.x * p.t_per_cta, t_end = min(p.T, t_begin + p.t_per_cta);
    const float4 sk = ldg4(p.sk + c), tk = ldg4(p.tk + c), sv = ldg4(p.sv + c), tvv = ldg4(p.tv + c);
    float s1k[4] = {0.f, 0.f, 0.f, 0.f}, s2k[4] = {0.f, 0.f, 0.f, 0.f}, s1v[4] = {0.f, 0.f, 0.f, 0.f}, s2v[4] = {0.f, 0.f, 0.f, 0.f};
    float avk[4] = {0.f, 0.f, 0.f, 0.f}, aat[4] = {0.f, 0.f, 0.f, 0.f};
    int cur_tv = -1;
    auto flush = [&]() {
        if (cur_tv < 0) return;
        float* d1 = p.dvk + ((long long)b * p.Tv + cur_tv) * 256 + c;
        float* d2 = p.datt + ((long long)b * p.Tv + cur_tv) * 256 + c;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            atomicAdd(d1 + k, avk[k]);
            atomicAdd(d2 + k, aat[k]);
            avk[k] = aat[k] = 0.f;
        }
    };
    for (int t = t_begin; t < t_end; ++t) {
        const int tv = nearest_src(t, p.Tv, p.T);
        if (tv != cur_tv) {
            flush();
            cur_tv = tv;
        }
        const long long vo = ((long long)b * p.Tv + tv) * 256 + c;
        const float4 k4 = ldg4(p.vk + vo), at4 = ldg4(p.att + vo);
        const float kk[4] = {k4.x, k4.y, k4.z, k4.w}, at[4] = {at4.x, at4.y, at4.z, at4.w};
        const float skk[4] = {sk.x, sk.y, sk.z, sk.w}, tkk[4] = {tk.x, tk.y, tk.z, tk.w}, svv[4] = {sv.x, sv.y, sv.z, sv.w}, tv4[4] = {tvv.x, tvv.y, tvv.z, tvv.w};
        for (int f = fl; f < p.F; f += 4) {
            const long long o = ((((long long)b * p.T + t) * p.F) + f) * 256 + c;
            const float4 d4 = ldg4(p.dout + o), a4 = ldg4(p.a + o);
            const float d[4] = {d4.x, d4.y, d4.z, d4.w}, x[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float key = fmaf(x[k], skk[k], tkk[k]);
                const float val = fmaf(x[k], svv[k], tv4[k]);
                const float gk = key > 0.f ? d[k] * kk[k] : 0.f;
                const float gv = d[k] * at[k];
                s1k[k] += gk;
                s2k[k] = fmaf(gk, x[k], s2k[k]);
                s1v[k] += gv;
                s2v[k] = fmaf(gv, x[k], s2v[k]);
                avk[k] = fmaf(d[k], fmaxf(key, 0.f), avk[k]);
                aat[k] = fmaf(d[k], val, aat[k]);
            }
        }
    }
    flush();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        atomicAdd(sh + (c + k) * 4 + 0, s1k[k]);
        atomicAdd(sh + (c + k) * 4 + 1, s2k[k]);
        atomicAdd(sh + (c + k) * 4 + 2, s1v[k]);
        atomicAdd(sh + (c + k) * 4 + 3, s2v[k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += 256) atomicAdd(p.csum + i, (double)sh[i]);
}

// pass 2: da = sk*g_k + sv*g_v - c0 - (a - mu)*c1   (BatchNorm backward through the batch statistics, both branches)
__global__ void __launch_bounds__(256) caf_bwd_apply_kernel(CafBwdArgs p, long long total4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i & 63) * 4;
        const long long pos = i >> 6;
        const long long bt = pos / p.F;
        const int t = (int)(bt % p.T);
        const long long b = bt / p.T;
        const int tv = nearest_src(t, p.Tv, p.T);
        const long long vo = (b * p.Tv + tv) * 256 + c;
        const float4 d4 = ldg4(p.dout + i * 4), a4 = ldg4(p.a + i * 4), k4 = ldg4(p.vk + vo), at4 = ldg4(p.att + vo);
        const float4 sk = ldg4(p.sk + c), tk = ldg4(p.tk + c), sv = ldg4(p.sv + c), mu = ldg4(p.mu + c), c0 = ldg4(p.c0 + c), c1 = ldg4(p.c1 + c);
        float4 o;
#define RTFS_CB(k) o.k = sk.k * (fmaf(a4.k, sk.k, tk.k) > 0.f ? d4.k * k4.k : 0.f) + sv.k * d4.k * at4.k - c0.k - (a4.k - mu.k) * c1.k;
        RTFS_CB(x) RTFS_CB(y) RTFS_CB(z) RTFS_CB(w)
#undef RTFS_CB
        *reinterpret_cast<float4*>(p.da + i * 4) = o;
    }
}

// video side of the CAF backward: one CTA per utterance, thread c owns audio channel c (video group c) as caf_video_kernel.
struct CafVideoBwdArgs {
    const float* v;  // (B,512,Tv)
    const float* wr; const float* br; const float* gr; const float* ber;
    const float* wa; const float* ba; const float* ga; const float* bea;
    const float* dvk;   // (B,Tv,256)
    const float* datt;  // (B,Tv,256)
    float* dv;          // (B,512,Tv)
    float* dwr; float* dbr; float* dgr; float* dber;   // [256][2],[256],[256],[256]   accumulated
    float* dwa; float* dba; float* dga; float* dbea;   // [1024][2],[1024] x3
    int Ca, Tv;
};

__global__ void __launch_bounds__(256) caf_video_bwd_kernel(CafVideoBwdArgs a) {
    extern __shared__ float att_s[];  // [Tv][Ca] softmax output
    __shared__ float red[4][8];
    __shared__ double st[4];
    const int c = threadIdx.x, b = blockIdx.x, Ca = a.Ca, Tv = a.Tv;
    const float* v0 = a.v + ((long long)b * 2 * Ca + 2 * c) * Tv;
    const float* v1 = v0 + Tv;
    const float wr0 = a.wr[2 * c], wr1 = a.wr[2 * c + 1], brc = a.br[c];
    float wa0[4], wa1[4], bac[4], gac[4], beac[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        wa0[i] = a.wa[2 * (4 * c + i)];
        wa1[i] = a.wa[2 * (4 * c + i) + 1];
        bac[i] = a.ba[4 * c + i];
        gac[i] = a.ga[4 * c + i];
        beac[i] = a.bea[4 * c + i];
    }
    const float grc = a.gr[c];
    auto block4 = [&](float (&vals)[4]) {  // block-wide sums of 4 values -> st[0..3] (double), visible to all threads
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float w = warp_sum(vals[i]);
            if ((c & 31) == 0) red[i][c >> 5] = w;
        }
        __syncthreads();
        if (c < 4) {
            double s = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += (double)red[c][w];
            st[c] = s;
        }
        __syncthreads();
    };
    // forward statistics of both gLNs
    float vals[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < Tv; ++t) {
        const float x0 = __ldg(v0 + t), x1 = __ldg(v1 + t);
        const float r = fmaf(wr0, x0, fmaf(wr1, x1, brc));
        vals[0] += r;
        vals[1] += r * r;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float y = fmaf(wa0[i], x0, fmaf(wa1[i], x1, bac[i]));
            vals[2] += y;
            vals[3] += y * y;
        }
    }
    block4(vals);
    const double nr = (double)Ca * Tv, na = 4.0 * nr;
    const double mr = st[0] / nr, ma = st[2] / na;
    double vr = st[1] / nr - mr * mr, va = st[3] / na - ma * ma;
    vr = vr < 0 ? 0 : vr;
    va = va < 0 ? 0 : va;
    const float mean_r = (float)mr, rstd_r = (float)(1.0 / sqrt(vr + 1e-5));
    const float mean_a = (float)ma, rstd_a = (float)(1.0 / sqrt(va + 1e-5));
    // forward softmax of the head-mean (per channel over Tv)
    float mx = -INFINITY;
    for (int t = 0; t < Tv; ++t) {
        const float x0 = __ldg(v0 + t), x1 = __ldg(v1 + t);
        float m = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) m += (fmaf(wa0[i], x0, fmaf(wa1[i], x1, bac[i])) - mean_a) * rstd_a * gac[i] + beac[i];
        m *= 0.25f;
        att_s[t * Ca + c] = m;
        mx = fmaxf(mx, m);
    }
    float den = 0.f;
    for (int t = 0; t < Tv; ++t) {
        const float e = __expf(att_s[t * Ca + c] - mx);
        att_s[t * Ca + c] = e;
        den += e;
    }
    const float inv = 1.f / den;
    float dot = 0.f;  // sum_t datt*att (softmax backward, per channel)
    for (int t = 0; t < Tv; ++t) {
        const float p = att_s[t * Ca + c] * inv;
        att_s[t * Ca + c] = p;
        dot += p * __ldg(a.datt + ((long long)b * Tv + t) * Ca + c);
    }
    // gLN backward sums over (channels, Tv): R1 = sum g*dn, R2 = sum g*dn*xhat for both norms; affine gradients per channel
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    float dgr = 0.f, dber = 0.f, dga[4] = {0.f, 0.f, 0.f, 0.f}, dbea[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < Tv; ++t) {
        const float x0 = __ldg(v0 + t), x1 = __ldg(v1 + t);
        const float xh = (fmaf(wr0, x0, fmaf(wr1, x1, brc)) - mean_r) * rstd_r;
        const float dk = __ldg(a.dvk + ((long long)b * Tv + t) * Ca + c);
        dgr += dk * xh;
        dber += dk;
        s4[0] += grc * dk;
        s4[1] += grc * dk * xh;
        const float p = att_s[t * Ca + c];
        const float dn = 0.25f * p * (__ldg(a.datt + ((long long)b * Tv + t) * Ca + c) - dot);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float xa = (fmaf(wa0[i], x0, fmaf(wa1[i], x1, bac[i])) - mean_a) * rstd_a;
            dga[i] += dn * xa;
            dbea[i] += dn;
            s4[2] += gac[i] * dn;
            s4[3] += gac[i] * dn * xa;
        }
    }
    block4(s4);
    const float r1 = (float)(st[0] / nr), r2 = (float)(st[1] / nr), a1 = (float)(st[2] / na), a2 = (float)(st[3] / na);
    float dwr0 = 0.f, dwr1 = 0.f, dbr = 0.f, dwa0[4] = {0.f, 0.f, 0.f, 0.f}, dwa1[4] = {0.f, 0.f, 0.f, 0.f}, dba[4] = {0.f, 0.f, 0.f, 0.f};
    float* dv0 = a.dv + ((long long)b * 2 * Ca + 2 * c) * Tv;
    float* dv1 = dv0 + Tv;
    for (int t = 0; t < Tv; ++t) {
        const float x0 = __ldg(v0 + t), x1 = __ldg(v1 + t);
        const float xh = (fmaf(wr0, x0, fmaf(wr1, x1, brc)) - mean_r) * rstd_r;
        const float dk = __ldg(a.dvk + ((long long)b * Tv + t) * Ca + c);
        const float dr = (grc * dk - r1 - xh * r2) * rstd_r;
        dwr0 += dr * x0;
        dwr1 += dr * x1;
        dbr += dr;
        float g0 = dr * wr0, g1 = dr * wr1;
        const float p = att_s[t * Ca + c];
        const float dn = 0.25f * p * (__ldg(a.datt + ((long long)b * Tv + t) * Ca + c) - dot);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float xa = (fmaf(wa0[i], x0, fmaf(wa1[i], x1, bac[i])) - mean_a) * rstd_a;
            const float dy = (gac[i] * dn - a1 - xa * a2) * rstd_a;
            dwa0[i] += dy * x0;
            dwa1[i] += dy * x1;
            dba[i] += dy;
            g0 = fmaf(dy, wa0[i], g0);
            g1 = fmaf(dy, wa1[i], g1);
        }
        dv0[t] = g0;
        dv1[t] = g1;
    }
    atomicAdd(a.dwr + 2 * c, dwr0);
    atomicAdd(a.dwr + 2 * c + 1, dwr1);
    atomicAdd(a.dbr + c, dbr);
    atomicAdd(a.dgr + c, dgr);
    atomicAdd(a.dber + c, dber);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        atomicAdd(a.dwa + 2 * (4 * c + i), dwa0[i]);
        atomicAdd(a.dwa + 2 * (4 * c + i) + 1, dwa1[i]);
        atomicAdd(a.dba + 4 * c + i, dba[i]);
        atomicAdd(a.dga + 4 * c + i, dga[i]);
        atomicAdd(a.dbea + 4 * c + i, dbea[i]);
    }
}

// ------------------------------------------------------------------------------------------------------------ iSTFT
// adjoint of dec_istft*_kernel's irDFT / window / overlap-add / envelope: dwav (B,L) -> dspec (B,T,129,2) = gradient w.r.t.
// the (Re, Im) planes the transposed 3x3 conv produces.  Frame t covers output samples n = 128*(t-1) + mm, mm in [0,256).
struct IstftBwdArgs {
    const float* dwav;
    const float* window;
    const float* costab;
    const float* sintab;
    float* dspec;
    int L, T;
};

__global__ void __launch_bounds__(288) istft_bwd_kernel(IstftBwdArgs a) {
    __shared__ float fr[256], ct[256], st[256];
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    if (tid < 256) {
        const int n = 128 * (t - 1) + tid;
        float v = 0.f;
        if (n >= 0 && n < a.L) {
            const int h = n >> 7, m = n & 127;
            const float w1 = __ldg(a.window + m), w0 = __ldg(a.window + m + 128);
            float env = w0 * w0;
            if (h + 1 < a.T) env += w1 * w1;
            v = __ldg(a.dwav + (long long)b * a.L + n) / env * __ldg(a.window + tid) * (1.f / 256.f);
        }
        fr[tid] = v;
        ct[tid] = __ldg(a.costab + tid);
        st[tid] = __ldg(a.sintab + tid);
    }
    __syncthreads();
    if (tid < 258) {
        const int f = tid >> 1, part = tid & 1;
        const float* tab = part ? st : ct;
        float acc = 0.f;
#pragma unroll 8
        for (int n = 0; n < 256; ++n) acc = fmaf(fr[n], tab[(n * f) & 255], acc);
        const bool edge = f == 0 || f == 128;
        const float r = part ? (edge ? 0.f : -2.f * acc) : (edge ? acc : 2.f * acc);
        a.dspec[(((long long)b * a.T + t) * 129 + f) * 2 + part] = r;
    }
}

// ------------------------------------------------------------------------------------------------------------- loss
// pairwise_neg_snr, n_src = 1: both signals zero-meaned, noise = est - target, loss_b = -10 log10(sum t^2 / (sum noise^2 + eps) + eps).
// pass 1: fp64 sums [B][5] = (sum e, sum t, sum e^2, sum t^2, sum e*t)
__global__ void __launch_bounds__(256) snr_sums_kernel(const float* __restrict__ est, const float* __restrict__ tgt, int L, double* sums) {
    __shared__ double sh[5][8];
    const int b = blockIdx.y;
    double s[5] = {0, 0, 0, 0, 0};
    float f[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
        const float e = __ldg(est + (long long)b * L + i), t = __ldg(tgt + (long long)b * L + i);
        f[0] += e;
        f[1] += t;
        f[2] = fmaf(e, e, f[2]);
        f[3] = fmaf(t, t, f[3]);
        f[4] = fmaf(e, t, f[4]);
        if (++cnt == 32) {  // short fp32 runs, fp64 across them
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                s[k] += (double)f[k];
                f[k] = 0.f;
            }
            cnt = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        s[k] += (double)f[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double tot = 0;
        for (int w = 0; w < 8; ++w) tot += sh[threadIdx.x][w];
        atomicAdd(sums + 5 * b + threadIdx.x, tot);
    }
}

// pass 2: loss[b] and (optionally) dest = scale * d loss_b / d est
__global__ void __launch_bounds__(256) snr_apply_kernel(const float* __restrict__ est, const float* __restrict__ tgt, int L, const double* __restrict__ sums,
                                                        float* loss, float* dest, float scale) {
    const int b = blockIdx.y;
    const double se = sums[5 * b], stt = sums[5 * b + 1], see = sums[5 * b + 2], st2 = sums[5 * b + 3], set = sums[5 * b + 4];
    const double eps = 1e-8;
    const double en_t = st2 - stt * stt / L;                             // sum of squares of the zero-mean target
    const double dm = (se - stt) / L;                                    // mean of est - target
    const double en_n = (see - 2.0 * set + st2) - (se - stt) * (se - stt) / L;  // sum of squares of the zero-mean noise
    const double sdr = en_t / (en_n + eps);
    if (blockIdx.x == 0 && threadIdx.x == 0) loss[b] = (float)(-10.0 * log10(sdr + eps));
    if (dest == nullptr) return;
    // d/d est[n] of -10 log10(sdr + eps) = (10/ln 10) / (sdr + eps) * en_t * 2 noise[n] / (en_n + eps)^2
    const float coef = (float)((10.0 / 2.302585092994046) / (sdr + eps) * en_t * 2.0 / ((en_n + eps) * (en_n + eps)) * scale);
    const float fm = (float)dm;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
        const long long o = (long long)b * L + i;
        dest[o] = coef * (__ldg(est + o) - __ldg(tgt + o) - fm);
    }
}

// ------------------------------------------------------------------------------------------------------------ AdamW
// out[0] += sum g^2 (fp64)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* out) {
    __shared__ double sh[8];
    double s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = __ldg(g + i);
        s += (double)v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0;
        for (int w = 0; w < 8; ++w) tot += sh[w];
        atomicAdd(out, tot);
    }
}

// clip_grad_norm_(max_norm) fused into torch.optim.AdamW's update (decoupled weight decay)
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                                                    float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                                                    const double* __restrict__ gnorm_sq, float max_norm, float grad_scale) {
    float clip = grad_scale;
    if (max_norm > 0.f && gnorm_sq != nullptr) {
        const float norm = (float)sqrt(*gnorm_sq) * grad_scale;
        const float coef = max_norm / (norm + 1e-6f);
        clip *= coef < 1.f ? coef : 1.f;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * clip;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        pi -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
        p[i] = pi;
    }
}

}  // namespace rtfs
