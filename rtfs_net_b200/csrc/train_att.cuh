// Backward / training-step kernels, part 4: TF self-attention (MultiHeadSelfAttention2D.forward, layers/attention.py:149-189;
// ConvActNorm = 1x1 conv -> PReLU -> LayerNormalization4D over (E,F), layers/conv_layers.py:201-205).
//
// Tape of the forward: Q, K (B,H,Tc,256), V (B,H,Tc,1024) (token rows, inner index f*E+e), AO (B,Tc,64,64; channel = h*16+j).
//   att_proj_bwd : d(out) -> d(pre-activation of attn_concat_proj) per frame (LN over (C,F) + PReLU backward, conv recomputed)
//   (GEMM) dAO = dpre_o W_o ; regroup to per-head token rows
//   (bgemm) S = Q K^T/16, dP = dO V^T ; softmax_bwd: P, dS ; (bgemm) dQ = dS K, dK = dS^T Q, dV = P^T dO
//   att_qkv_bwd  : dQ/dK/dV -> d(pre-activation of the 12 head convs) per frame ; (GEMM) dg += dpre W_qkv
// The frame kernels loop over frames with a fixed element -> thread map, so the LN affine gradients (per (f,e) element) and the
// PReLU slope gradients accumulate in registers and reach memory once per CTA.
#pragma once
#include "common.cuh"

namespace rtfs {

// block-wide sum of two values, result broadcast to every thread (scratch: >= 2*nwarps + 2 floats)
DEVINL void block_sum2(float& a, float& b, float* scratch) {
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) {
        scratch[2 * w] = a;
        scratch[2 * w + 1] = b;
    }
    __syncthreads();
    if (w == 0) {
        float x = lane < nw ? scratch[2 * lane] : 0.f, y = lane < nw ? scratch[2 * lane + 1] : 0.f;
        x = warp_sum(x);
        y = warp_sum(y);
        if (lane == 0) {
            scratch[2 * nw] = x;
            scratch[2 * nw + 1] = y;
        }
    }
    __syncthreads();
    a = scratch[2 * nw];
    b = scratch[2 * nw + 1];
}

// ------------------------------------------------------------------------------------------------ concat projection
struct AttProjBwdArgs {
    const float* ao;     // (B*Tc, 64 f, 64 c)
    const float* dout;   // (B*Tc, 64 f, 64 j) gradient w.r.t. the attention output (before the residual is split off)
    const float* W;      // [64 j][64 c] fp32
    const float* bias;   // [64]
    const float* slope;  // [1]
    const float* gamma;  // [f*64 + j]
    float* dpre;         // (B*Tc, 64, 64) gradient w.r.t. W*ao + b
    float* dgamma;       // [4096] accumulated
    float* dbeta;        // [4096]
    float* dslope;       // [1]
    int nframes;
};

__global__ void __launch_bounds__(256) att_proj_bwd_kernel(AttProjBwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* Xs = sm;             // [64][65]
    float* Ws = Xs + 64 * 65;   // [64][65]
    float* scratch = Ws + 64 * 65;  // [32]
    const int tid = threadIdx.x, j = tid & 63, fb = tid >> 6;
    for (int i = tid; i < 4096; i += 256) Ws[(i >> 6) * 65 + (i & 63)] = __ldg(a.W + i);
    const float bj = __ldg(a.bias + j), sl = __ldg(a.slope);
    float gmm[16], dg[16], db[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        gmm[i] = __ldg(a.gamma + (fb + 4 * i) * 64 + j);
        dg[i] = db[i] = 0.f;
    }
    float dsl = 0.f;
    for (int fr = blockIdx.x; fr < a.nframes; fr += gridDim.x) {
        __syncthreads();
        const float* xin = a.ao + (long long)fr * 4096;
        for (int i = tid; i < 4096; i += 256) Xs[(i >> 6) * 65 + (i & 63)] = __ldg(xin + i);
        __syncthreads();
        float pre[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pre[i] = bj;
        for (int c = 0; c < 64; ++c) {
            const float w = Ws[j * 65 + c];
#pragma unroll
            for (int i = 0; i < 16; ++i) pre[i] = fmaf(Xs[(fb + 4 * i) * 65 + c], w, pre[i]);
        }
        float s = 0.f, dummy = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += prelu(pre[i], sl);
        block_sum2(s, dummy, scratch);
        const float mu = s * (1.f / 4096.f);
        float q = 0.f;
        dummy = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float d = prelu(pre[i], sl) - mu;
            q += d * d;
        }
        block_sum2(q, dummy, scratch);
        const float rs = 1.f / sqrtf(q * (1.f / 4096.f) + RTFS_EPS);
        float dy[16], s1 = 0.f, s2 = 0.f;
        const float* din = a.dout + (long long)fr * 4096;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            dy[i] = __ldg(din + (fb + 4 * i) * 64 + j);
            const float xh = (prelu(pre[i], sl) - mu) * rs;
            dg[i] += dy[i] * xh;
            db[i] += dy[i];
            const float gy = dy[i] * gmm[i];
            s1 += gy;
            s2 += gy * xh;
        }
        block_sum2(s1, s2, scratch);
        s1 *= (1.f / 4096.f);
        s2 *= (1.f / 4096.f);
        float* dp = a.dpre + (long long)fr * 4096;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float xh = (prelu(pre[i], sl) - mu) * rs;
            float da = (dy[i] * gmm[i] - s1 - xh * s2) * rs;
            if (pre[i] < 0.f) {
                dsl += da * pre[i];
                da *= sl;
            }
            dp[(fb + 4 * i) * 64 + j] = da;
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        atomicAdd(a.dgamma + (fb + 4 * i) * 64 + j, dg[i]);
        atomicAdd(a.dbeta + (fb + 4 * i) * 64 + j, db[i]);
    }
    float dummy = 0.f;
    block_sum2(dsl, dummy, scratch);
    if (tid == 0 && dsl != 0.f) atomicAdd(a.dslope, dsl);
}

// dAO (B,Tc,64 f,64 ch = h*16+j) -> dO (B,H,Tc,1024: f*16+j)
__global__ void __launch_bounds__(256) att_regroup_kernel(const float* __restrict__ dao, float* __restrict__ d_o, int B, int Tc, int H) {
    const long long total4 = (long long)B * Tc * 64 * 16;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total4; idx += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx & 15);            // channel quad: h = c4 >> 2, j = (c4 & 3) * 4
        const long long pos = idx >> 4;            // (b*Tc + t)*64 + f
        const int f = (int)(pos & 63);
        const long long bt = pos >> 6;
        const int t = (int)(bt % Tc);
        const long long b = bt / Tc;
        const int h = c4 >> 2, j = (c4 & 3) * 4;
        const float4 v = ldg4(dao + idx * 4);
        *reinterpret_cast<float4*>(d_o + (((b * H + h) * Tc + t) * 1024) + f * 16 + j) = v;
    }
}

// one warp per score row: S (already scaled) -> P = softmax(S) (in place); dP -> dS = P * (dP - sum(dP*P)) * scale (in place)
__global__ void __launch_bounds__(256) softmax_bwd_kernel(float* S, float* dP, long long rows, int n, float scale) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float* s = S + row * n;
    float* d = dP + row * n;
    float mx = -INFINITY;
    for (int i = lane; i < n; i += 32) mx = fmaxf(mx, s[i]);
    mx = warp_max(mx);
    float den = 0.f;
    for (int i = lane; i < n; i += 32) den += expf(s[i] - mx);
    den = warp_sum(den);
    const float inv = 1.f / den;
    float dot = 0.f;
    for (int i = lane; i < n; i += 32) {
        const float p = expf(s[i] - mx) * inv;
        s[i] = p;
        dot += p * d[i];
    }
    dot = warp_sum(dot);
    for (int i = lane; i < n; i += 32) d[i] = s[i] * (d[i] - dot) * scale;
}

// ------------------------------------------------------------------------------------------------ head convs
struct AttQkvBwdArgs {
    const float* x;      // (B*Tc, 64 f, 64 c) attention input
    const float* W;      // [96][64] fp32
    const float* bias;   // [96]
    const float* slope;  // [12]
    const float* gamma;  // groups concatenated, [f*E+e] inside a group
    const float* dq;     // (B,H,Tc,256)
    const float* dk;
    const float* dv;     // (B,H,Tc,1024)
    float* dpre;         // (B*Tc, 64, 96)
    float* dgamma;       // [6144] accumulated
    float* dbeta;
    float* dslope;       // [12]
    int nframes, Tc, H;
};

DEVINL void qkv_group(int j, int& grp, int& col0, int& E, int& goff) {
    if (j < 16) { grp = j >> 2; col0 = grp * 4; E = 4; goff = grp * 256; }
    else if (j < 32) { grp = 4 + ((j - 16) >> 2); col0 = 16 + (grp - 4) * 4; E = 4; goff = 1024 + (grp - 4) * 256; }
    else { grp = 8 + ((j - 32) >> 4); col0 = 32 + (grp - 8) * 16; E = 16; goff = 2048 + (grp - 8) * 1024; }
}

// 384 threads: thread = (column j of the 96 conv outputs, row phase fb): it owns rows f = fb + 4i, i < 16, of its column, so the
// LN affine gradients of its 16 (f, j) elements stay in registers over the frame loop and the per-group LN sums are formed from
// one partial per thread (shared-memory table [4][96], no atomics).
__global__ void __launch_bounds__(384) att_qkv_bwd_kernel(AttQkvBwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* Xs = sm;               // [64][65]
    float* Ws = Xs + 64 * 65;     // [96][65]
    float* part = Ws + 96 * 65;   // [2][4][96] per-thread partial sums (two quantities)
    float* dsl_s = part + 768;    // [12]
    const int tid = threadIdx.x, j = tid % 96, fb = tid / 96;
    for (int i = tid; i < 96 * 64; i += 384) Ws[(i >> 6) * 65 + (i & 63)] = __ldg(a.W + i);
    if (tid < 12) dsl_s[tid] = 0.f;
    int grp, col0, E, goff;
    qkv_group(j, grp, col0, E, goff);
    const float bj = __ldg(a.bias + j), sl = __ldg(a.slope + grp);
    const float invn = 1.f / (float)(64 * E);
    const int h = grp & 3, e = j - col0;
    float gmm[16], dg[16], db[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        gmm[i] = __ldg(a.gamma + goff + (fb + 4 * i) * E + e);
        dg[i] = db[i] = 0.f;
    }
    float dsl = 0.f;
    // sum over the group's E columns x 4 row phases of one partial per thread
    auto group_sum = [&](const float* tab) {
        float s = 0.f;
        for (int q = 0; q < 4; ++q)
            for (int c = 0; c < E; ++c) s += tab[q * 96 + col0 + c];
        return s;
    };
    for (int fr = blockIdx.x; fr < a.nframes; fr += gridDim.x) {
        __syncthreads();
        const float* xin = a.x + (long long)fr * 4096;
        for (int i = tid; i < 4096; i += 384) Xs[(i >> 6) * 65 + (i & 63)] = __ldg(xin + i);
        __syncthreads();
        float pre[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pre[i] = bj;
        for (int c = 0; c < 64; ++c) {
            const float w = Ws[j * 65 + c];
#pragma unroll
            for (int i = 0; i < 16; ++i) pre[i] = fmaf(Xs[(fb + 4 * i) * 65 + c], w, pre[i]);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += prelu(pre[i], sl);
        part[fb * 96 + j] = s;
        __syncthreads();
        const float mu = group_sum(part) * invn;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float d = prelu(pre[i], sl) - mu;
            q += d * d;
        }
        part[384 + fb * 96 + j] = q;
        __syncthreads();
        const float rs = 1.f / sqrtf(group_sum(part + 384) * invn + RTFS_EPS);
        const int b = fr / a.Tc, t = fr - b * a.Tc;
        const long long tok = ((long long)b * a.H + h) * a.Tc + t;
        const float* src = grp < 4 ? a.dq + tok * 256 : (grp < 8 ? a.dk + tok * 256 : a.dv + tok * 1024);
        float dy[16], s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            dy[i] = __ldg(src + (fb + 4 * i) * E + e);
            const float xh = (prelu(pre[i], sl) - mu) * rs;
            dg[i] += dy[i] * xh;
            db[i] += dy[i];
            const float gy = dy[i] * gmm[i];
            s1 += gy;
            s2 += gy * xh;
        }
        __syncthreads();  // every thread has read the centred-square table
        part[fb * 96 + j] = s1;
        part[384 + fb * 96 + j] = s2;
        __syncthreads();
        const float m1 = group_sum(part) * invn, m2 = group_sum(part + 384) * invn;
        float* dp = a.dpre + (long long)fr * (64 * 96);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float xh = (prelu(pre[i], sl) - mu) * rs;
            float da = (dy[i] * gmm[i] - m1 - xh * m2) * rs;
            if (pre[i] < 0.f) {
                dsl += da * pre[i];
                da *= sl;
            }
            dp[(fb + 4 * i) * 96 + j] = da;
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int gi = goff + (fb + 4 * i) * E + e;
        atomicAdd(a.dgamma + gi, dg[i]);
        atomicAdd(a.dbeta + gi, db[i]);
    }
    atomicAdd(dsl_s + grp, dsl);
    __syncthreads();
    if (tid < 12 && dsl_s[tid] != 0.f) atomicAdd(a.dslope + tid, dsl_s[tid]);
}

}  // namespace rtfs
