// Backward / training-step kernels of the RTFS-Net path, part 1: normalisations, depthwise stencils, TF-AR units, pooling.
//
// The training step (BASELINE configs[2]: RTFS-Net-6, SNR loss, gradient all-reduce; reference call chain
// src/system/core.py:94-117 -> AVNet.forward -> loss.backward()) runs the production forward kernels with a per-pass
// tape (every pre-normalisation tensor and the gLN statistics of a block pass stay resident) and the kernels below for the
// backward.  Each kernel is the adjoint of one forward piece and cites it:
//   gLN  = GlobalLayerNorm (layers/normalizations.py:8-17), depthwise 4x4 convs of ConvNormAct (layers/conv_layers.py:65-129),
//   TF-AR = InjectionMultiSum (layers/fusion.py:54-69), adaptive_avg_pool2d + sum (separators/tdanet.py:117-118).
// Layout as in the forward: channels-last (B,T,F,C) fp32; gLN statistics are fp64 (sum, sum of squares) per sample.
// All parameter gradients are ACCUMULATED (atomicAdd) into caller-zeroed buffers: the block is shared by every pass.
#pragma once
#include "common.cuh"

namespace rtfs {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_PRELU = 2, ACT_SIGMOID = 3 };

DEVINL float sigmoid_exact(float x) { return 1.f / (1.f + expf(-x)); }

// Per-channel partial sums of a CTA -> global: shared-memory atomics first, one global atomic per channel and CTA.
// `sh` holds n floats, zeroed by the caller before the accumulation phase (with a __syncthreads in between).
DEVINL void flush_shared_to_global(const float* sh, float* dst, int n) {
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = sh[i];
        if (v != 0.f) atomicAdd(dst + i, v);
    }
}

// ---------------------------------------------------------------------------------------------------------------- gLN
// forward materialisation: y = act(gLN(x))           (re-forms p = PReLU(gLN(p_pre)), d0 = gLN(d0_pre), d1 = gLN(d1_pre))
template <int C, int ACT>
__global__ void __launch_bounds__(256) gln_apply_kernel(const float* __restrict__ x, GlnRef gln, const float* __restrict__ slope,
                                                        float* __restrict__ y, long long n4_per_sample) {
    const int b = blockIdx.y;
    float mean, rstd;
    gln_mean_rstd(gln.sums, b, gln.inv_n, mean, rstd);
    const float a = ACT == ACT_PRELU ? __ldg(slope) : 0.f;
    const float4* xp = reinterpret_cast<const float4*>(x) + (long long)b * n4_per_sample;
    float4* yp = reinterpret_cast<float4*>(y) + (long long)b * n4_per_sample;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4_per_sample; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % (C / 4)) * 4;
        const float4 v = __ldg(xp + i), gm = ldg4(gln.gamma + c), be = ldg4(gln.beta + c);
        float4 o;
        o.x = (v.x - mean) * rstd * gm.x + be.x;
        o.y = (v.y - mean) * rstd * gm.y + be.y;
        o.z = (v.z - mean) * rstd * gm.z + be.z;
        o.w = (v.w - mean) * rstd * gm.w + be.w;
        if (ACT == ACT_PRELU) {
            o.x = prelu(o.x, a);
            o.y = prelu(o.y, a);
            o.z = prelu(o.z, a);
            o.w = prelu(o.w, a);
        }
        yp[i] = o;
    }
}

// backward, pass 1.  dy (in) = gradient w.r.t. act(gLN(x_pre)); dy (out, in place) = dn = gradient w.r.t. the gLN output.
//   red[b] += (sum gamma*dn, sum gamma*dn*xhat)  [fp64]    dgamma[c] += sum dn*xhat    dbeta[c] += sum dn
//   ACT_PRELU: dslope += sum dy*n*[n<0] ;  ACT_RELU: dn = dy*[n>0] ;  ACT_SIGMOID: dn = dy*s(1-s)
struct GlnBwdArgs {
    float* dy;
    const float* x_pre;
    GlnRef gln;
    const float* slope;
    double* red;  // [B][2], zeroed by the caller
    float* dgamma;
    float* dbeta;
    float* dslope;
    long long n4_per_sample;
};

template <int C, int ACT>
__global__ void __launch_bounds__(256) gln_bwd_reduce_kernel(GlnBwdArgs a) {
    __shared__ float sh[2 * C + 1];
    __shared__ float scratch[16];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < 2 * C + 1; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    float mean, rstd;
    gln_mean_rstd(a.gln.sums, b, a.gln.inv_n, mean, rstd);
    const float sl = ACT == ACT_PRELU ? __ldg(a.slope) : 0.f;
    const float4* xp = reinterpret_cast<const float4*>(a.x_pre) + (long long)b * a.n4_per_sample;
    float4* dp = reinterpret_cast<float4*>(a.dy) + (long long)b * a.n4_per_sample;
    // 256 threads and C/4 channel quads: a thread's channel quad is the same in every iteration (256 % (C/4) == 0 and the
    // grid stride is a multiple of 256)
    const int c = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) % (C / 4)) * 4;
    const float4 gm = ldg4(a.gln.gamma + c), be = ldg4(a.gln.beta + c);
    float s1 = 0.f, s2 = 0.f, dsl = 0.f;
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4_per_sample; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(xp + i);
        float4 d = dp[i];
        const float h0 = (v.x - mean) * rstd, h1 = (v.y - mean) * rstd, h2 = (v.z - mean) * rstd, h3 = (v.w - mean) * rstd;
        if (ACT != ACT_NONE) {
            const float n0 = h0 * gm.x + be.x, n1 = h1 * gm.y + be.y, n2 = h2 * gm.z + be.z, n3 = h3 * gm.w + be.w;
            if (ACT == ACT_RELU) {
                d.x = n0 > 0.f ? d.x : 0.f;
                d.y = n1 > 0.f ? d.y : 0.f;
                d.z = n2 > 0.f ? d.z : 0.f;
                d.w = n3 > 0.f ? d.w : 0.f;
            } else if (ACT == ACT_PRELU) {
                if (n0 < 0.f) { dsl += d.x * n0; d.x *= sl; }
                if (n1 < 0.f) { dsl += d.y * n1; d.y *= sl; }
                if (n2 < 0.f) { dsl += d.z * n2; d.z *= sl; }
                if (n3 < 0.f) { dsl += d.w * n3; d.w *= sl; }
            } else {
                const float g0 = sigmoid_exact(n0), g1 = sigmoid_exact(n1), g2 = sigmoid_exact(n2), g3 = sigmoid_exact(n3);
                d.x *= g0 * (1.f - g0);
                d.y *= g1 * (1.f - g1);
                d.z *= g2 * (1.f - g2);
                d.w *= g3 * (1.f - g3);
            }
            dp[i] = d;
        }
        const float w0 = gm.x * d.x, w1 = gm.y * d.y, w2 = gm.z * d.z, w3 = gm.w * d.w;
        s1 += w0 + w1 + w2 + w3;
        s2 += w0 * h0 + w1 * h1 + w2 * h2 + w3 * h3;
        dg.x += d.x * h0;
        dg.y += d.y * h1;
        dg.z += d.z * h2;
        dg.w += d.w * h3;
        db.x += d.x;
        db.y += d.y;
        db.z += d.z;
        db.w += d.w;
    }
    atomicAdd(sh + c, dg.x);
    atomicAdd(sh + c + 1, dg.y);
    atomicAdd(sh + c + 2, dg.z);
    atomicAdd(sh + c + 3, dg.w);
    atomicAdd(sh + C + c, db.x);
    atomicAdd(sh + C + c + 1, db.y);
    atomicAdd(sh + C + c + 2, db.z);
    atomicAdd(sh + C + c + 3, db.w);
    if (ACT == ACT_PRELU) {
        dsl = warp_sum(dsl);
        if ((threadIdx.x & 31) == 0) atomicAdd(sh + 2 * C, dsl);
    }
    block_stats_atomic(s1, s2, a.red + 2 * b, scratch);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        atomicAdd(a.dgamma + i, sh[i]);
        atomicAdd(a.dbeta + i, sh[C + i]);
    }
    if (ACT == ACT_PRELU && threadIdx.x == 0) atomicAdd(a.dslope, sh[2 * C]);
}

// backward, pass 2: dx = (gamma*dn - red0/n - xhat*red1/n) * rstd ; dst = dx or dst += dx (dst may alias dn)
template <int C>
__global__ void __launch_bounds__(256) gln_bwd_apply_kernel(const float* dn, const float* __restrict__ x_pre, GlnRef gln,
                                                            const double* __restrict__ red, float* dst, int accumulate, long long n4_per_sample) {
    const int b = blockIdx.y;
    float mean, rstd;
    gln_mean_rstd(gln.sums, b, gln.inv_n, mean, rstd);
    const float m1 = (float)(red[2 * b] * gln.inv_n), m2 = (float)(red[2 * b + 1] * gln.inv_n);
    const float4* xp = reinterpret_cast<const float4*>(x_pre) + (long long)b * n4_per_sample;
    const float4* dp = reinterpret_cast<const float4*>(dn) + (long long)b * n4_per_sample;
    float4* op = reinterpret_cast<float4*>(dst) + (long long)b * n4_per_sample;
    const int c = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) % (C / 4)) * 4;
    const float4 gm = ldg4(gln.gamma + c);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4_per_sample; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(xp + i), d = dp[i];
        float4 o;
        o.x = (gm.x * d.x - m1 - (v.x - mean) * rstd * m2) * rstd;
        o.y = (gm.y * d.y - m1 - (v.y - mean) * rstd * m2) * rstd;
        o.z = (gm.z * d.z - m1 - (v.z - mean) * rstd * m2) * rstd;
        o.w = (gm.w * d.w - m1 - (v.w - mean) * rstd * m2) * rstd;
        if (accumulate) {
            const float4 p = op[i];
            o.x += p.x;
            o.y += p.y;
            o.z += p.z;
            o.w += p.w;
        }
        op[i] = o;
    }
}

// ------------------------------------------------------------------------------------------------- depthwise 4x4 convs
// Forward (dwroll.cuh): y[to][fo] = sum_{i,j} x[to*s + i - 1][fo*s + j - 1] * w[i*4+j]  (+ bias); s = 1: 'same' (pad 1 before,
// 2 after); s = 2: pad 1.  Weights tap-major [16][64].
// data gradient: dx[t][f] = sum over the (to,fo,i,j) with to*s+i-1 == t, fo*s+j-1 == f of dy[to][fo] * w[i*4+j];
// up to NW convs that read the same input are summed (dg3 collects four, df1 two).
template <int NW>
struct DwBwdDataArgs {
    const float* dy[NW];
    const float* w[NW];
    float* dx;
    int accumulate;
    int Ti, Fi, To, Fo, stride;
    long long total4;  // B*Ti*Fi*16
};

template <int NW>
__global__ void __launch_bounds__(256) dw_bwd_data_kernel(DwBwdDataArgs<NW> a) {
    const int row_off = a.Fo * 64;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < a.total4; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx & 15) * 4;
        const int pos = (int)(idx >> 4);
        const int f = pos % a.Fi;
        const int bt = pos / a.Fi;
        const int t = bt % a.Ti;
        const int b = bt / a.Ti;
        // output rows / columns that read input (t, f) through tap i / j: to = (t + 1 - i) / s when divisible and in range
        int to[4], fo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int v = t + 1 - i, w = f + 1 - i;
            if (a.stride == 2) {
                to[i] = (v & 1) ? -1 : (v >> 1);
                fo[i] = (w & 1) ? -1 : (w >> 1);
            } else {
                to[i] = v;
                fo[i] = w;
            }
            if ((unsigned)to[i] >= (unsigned)a.To) to[i] = -1;
            if ((unsigned)fo[i] >= (unsigned)a.Fo) fo[i] = -1;
        }
        const long long bbase = (long long)b * a.To * a.Fo * 64 + c;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (to[i] < 0) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (fo[j] < 0) continue;
                const long long o = bbase + to[i] * row_off + fo[j] * 64;
#pragma unroll
                for (int k = 0; k < NW; ++k) {
                    const float4 d = ldg4(a.dy[k] + o), w = ldg4(a.w[k] + (i * 4 + j) * 64 + c);
                    acc.x = fmaf(d.x, w.x, acc.x);
                    acc.y = fmaf(d.y, w.y, acc.y);
                    acc.z = fmaf(d.z, w.z, acc.z);
                    acc.w = fmaf(d.w, w.w, acc.w);
                }
            }
        }
        float4* o = reinterpret_cast<float4*>(a.dx) + idx;
        if (a.accumulate) {
            const float4 p = *o;
            acc.x += p.x;
            acc.y += p.y;
            acc.z += p.z;
            acc.w += p.w;
        }
        *o = acc;
    }
}

// weight gradient: dw[i*4+j][c] += sum_{b,to,fo} dy[b][to][fo][c] * x[b][to*s+i-1][fo*s+j-1][c] ; dbias[c] += sum dy
struct DwBwdWArgs {
    const float* dy;  // (B,To,Fo,64)
    const float* x;   // (B,Ti,Fi,64) the conv input (materialised)
    float* dw;        // [16][64]
    float* dbias;     // [64] or null
    int B, Ti, Fi, To, Fo, stride;
};

__global__ void __launch_bounds__(256) dw_bwd_weight_kernel(DwBwdWArgs a) {
    __shared__ float sh[17 * 64];
    for (int i = threadIdx.x; i < 17 * 64; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    const int c = (threadIdx.x & 15) * 4;
    const int npos = a.B * a.To * a.Fo;  // < 2^31 / 64 by the library's size limit
    const int row_off = a.Fi * 64;       // element offset between input rows
    float4 acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ab = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int pos = blockIdx.x * 16 + (threadIdx.x >> 4); pos < npos; pos += gridDim.x * 16) {
        const int fo = pos % a.Fo;
        const int bt = pos / a.Fo;
        const int to = bt % a.To;
        const int b = bt / a.To;
        const float4 d = ldg4(a.dy + (long long)pos * 64 + c);
        ab.x += d.x;
        ab.y += d.y;
        ab.z += d.z;
        ab.w += d.w;
        const int t0 = to * a.stride - 1, f0 = fo * a.stride - 1;
        // one 64-bit base per position, 32-bit tap offsets, 4 + 4 validity flags
        const float* base = a.x + (((long long)b * a.Ti + t0) * a.Fi + f0) * 64 + c;
        bool vt[4], vf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            vt[i] = (unsigned)(t0 + i) < (unsigned)a.Ti;
            vf[i] = (unsigned)(f0 + i) < (unsigned)a.Fi;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (vt[i] && vf[j]) {
                    const float4 x = ldg4(base + i * row_off + j * 64);
                    acc[i * 4 + j].x = fmaf(d.x, x.x, acc[i * 4 + j].x);
                    acc[i * 4 + j].y = fmaf(d.y, x.y, acc[i * 4 + j].y);
                    acc[i * 4 + j].z = fmaf(d.z, x.z, acc[i * 4 + j].z);
                    acc[i * 4 + j].w = fmaf(d.w, x.w, acc[i * 4 + j].w);
                }
            }
        }
    }
    // lanes l and l + 16 of a warp hold the same channel quad: fold them before the shared-memory atomics
    const bool lead = (threadIdx.x & 16) == 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, 16);
        acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, 16);
        acc[k].z += __shfl_xor_sync(0xffffffffu, acc[k].z, 16);
        acc[k].w += __shfl_xor_sync(0xffffffffu, acc[k].w, 16);
        if (lead) {
            atomicAdd(sh + k * 64 + c, acc[k].x);
            atomicAdd(sh + k * 64 + c + 1, acc[k].y);
            atomicAdd(sh + k * 64 + c + 2, acc[k].z);
            atomicAdd(sh + k * 64 + c + 3, acc[k].w);
        }
    }
    ab.x += __shfl_xor_sync(0xffffffffu, ab.x, 16);
    ab.y += __shfl_xor_sync(0xffffffffu, ab.y, 16);
    ab.z += __shfl_xor_sync(0xffffffffu, ab.z, 16);
    ab.w += __shfl_xor_sync(0xffffffffu, ab.w, 16);
    if (lead) {
        atomicAdd(sh + 16 * 64 + c, ab.x);
        atomicAdd(sh + 16 * 64 + c + 1, ab.y);
        atomicAdd(sh + 16 * 64 + c + 2, ab.z);
        atomicAdd(sh + 16 * 64 + c + 3, ab.w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) atomicAdd(a.dw + i, sh[i]);
    if (a.dbias != nullptr)
        for (int i = threadIdx.x; i < 64; i += blockDim.x) atomicAdd(a.dbias + i, sh[16 * 64 + i]);
}

// ------------------------------------------------------------------------------------------------------------ TF-AR
// forward (re-materialisation of f0, f1, e for the weight gradients of the convs that consumed them):
//   out = gLN_l(l_pre) * sigmoid(gLN_g(g_pre))[up] + gLN_e(e_pre)[up] (+ gLN_d(d_pre))       layers/fusion.py:54-69
struct TfarArgs {
    const float* l_pre;  // (B,T,F,64)
    const float* g_pre;  // (B,Tc,Fc,64)
    const float* e_pre;  // (B,Tc,Fc,64)
    const float* d_pre;  // optional (B,T,F,64)
    GlnRef nl, ng, ne, nd;
    int T, F, Tc, Fc;
    long long total4;  // B*T*F*16
};

DEVINL float4 gln4(float4 v, float mean, float rstd, float4 gm, float4 be) {
    return make_float4((v.x - mean) * rstd * gm.x + be.x, (v.y - mean) * rstd * gm.y + be.y, (v.z - mean) * rstd * gm.z + be.z,
                       (v.w - mean) * rstd * gm.w + be.w);
}

__global__ void __launch_bounds__(256) tfar_fwd_kernel(TfarArgs a, float* __restrict__ out) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < a.total4; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx & 15) * 4;
        const long long pos = idx >> 4;
        const int f = (int)(pos % a.F);
        const long long bt = pos / a.F;
        const int t = (int)(bt % a.T);
        const int b = (int)(bt / a.T);
        const int tc = nearest_src32(t, a.Tc, a.T), fc = nearest_src32(f, a.Fc, a.F);
        const long long og = ((((long long)b * a.Tc + tc) * a.Fc) + fc) * 64 + c;
        float ml, rl, mg, rg, me, re;
        gln_mean_rstd(a.nl.sums, b, a.nl.inv_n, ml, rl);
        gln_mean_rstd(a.ng.sums, b, a.ng.inv_n, mg, rg);
        gln_mean_rstd(a.ne.sums, b, a.ne.inv_n, me, re);
        const float4 l = gln4(ldg4(a.l_pre + idx * 4), ml, rl, ldg4(a.nl.gamma + c), ldg4(a.nl.beta + c));
        const float4 g = gln4(ldg4(a.g_pre + og), mg, rg, ldg4(a.ng.gamma + c), ldg4(a.ng.beta + c));
        const float4 e = gln4(ldg4(a.e_pre + og), me, re, ldg4(a.ne.gamma + c), ldg4(a.ne.beta + c));
        float4 o;
        o.x = l.x * sigmoid_exact(g.x) + e.x;
        o.y = l.y * sigmoid_exact(g.y) + e.y;
        o.z = l.z * sigmoid_exact(g.z) + e.z;
        o.w = l.w * sigmoid_exact(g.w) + e.w;
        if (a.d_pre != nullptr) {
            float md, rd;
            gln_mean_rstd(a.nd.sums, b, a.nd.inv_n, md, rd);
            const float4 d = gln4(ldg4(a.d_pre + idx * 4), md, rd, ldg4(a.nd.gamma + c), ldg4(a.nd.beta + c));
            o.x += d.x;
            o.y += d.y;
            o.z += d.z;
            o.w += d.w;
        }
        *reinterpret_cast<float4*>(out + idx * 4) = o;
    }
}

// backward, local side: dl[pos] = dout[pos] * sigmoid(gLN_g(g_pre))[up(pos)]   (gradient w.r.t. gLN_l(l_pre))
__global__ void __launch_bounds__(256) tfar_bwd_local_kernel(TfarArgs a, const float* __restrict__ dout, float* __restrict__ dl) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < a.total4; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx & 15) * 4;
        const long long pos = idx >> 4;
        const int f = (int)(pos % a.F);
        const long long bt = pos / a.F;
        const int t = (int)(bt % a.T);
        const int b = (int)(bt / a.T);
        const int tc = nearest_src32(t, a.Tc, a.T), fc = nearest_src32(f, a.Fc, a.F);
        const long long og = ((((long long)b * a.Tc + tc) * a.Fc) + fc) * 64 + c;
        float mg, rg;
        gln_mean_rstd(a.ng.sums, b, a.ng.inv_n, mg, rg);
        const float4 g = gln4(ldg4(a.g_pre + og), mg, rg, ldg4(a.ng.gamma + c), ldg4(a.ng.beta + c));
        const float4 d = ldg4(dout + idx * 4);
        *reinterpret_cast<float4*>(dl + idx * 4) =
            make_float4(d.x * sigmoid_exact(g.x), d.y * sigmoid_exact(g.y), d.z * sigmoid_exact(g.z), d.w * sigmoid_exact(g.w));
    }
}

// backward, global side: every (tc,fc) collects the local positions that nearest-up-sampling mapped to it:
//   dg[tc,fc] = s(1-s) * sum dout*gLN_l(l_pre)   (gradient w.r.t. gLN_g(g_pre), sigmoid derivative included)
//   de[tc,fc] = sum dout                         (gradient w.r.t. gLN_e(e_pre))
DEVINL int ceil_div_i(int a, int b) { return (a + b - 1) / b; }

__global__ void __launch_bounds__(256) tfar_bwd_global_kernel(TfarArgs a, const float* __restrict__ dout, float* __restrict__ dg, float* __restrict__ de) {
    const long long totalg = a.total4 / ((long long)a.T * a.F) * ((long long)a.Tc * a.Fc);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < totalg; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx & 15) * 4;
        const long long pos = idx >> 4;
        const int fc = (int)(pos % a.Fc);
        const long long bt = pos / a.Fc;
        const int tc = (int)(bt % a.Tc);
        const int b = (int)(bt / a.Tc);
        // local indices t with floor(t*Tc/T) == tc  <=>  ceil(tc*T/Tc) <= t < ceil((tc+1)*T/Tc)
        const int t0 = ceil_div_i(tc * a.T, a.Tc), t1 = min(a.T, ceil_div_i((tc + 1) * a.T, a.Tc));
        const int f0 = ceil_div_i(fc * a.F, a.Fc), f1 = min(a.F, ceil_div_i((fc + 1) * a.F, a.Fc));
        float ml, rl, mg, rg;
        gln_mean_rstd(a.nl.sums, b, a.nl.inv_n, ml, rl);
        gln_mean_rstd(a.ng.sums, b, a.ng.inv_n, mg, rg);
        const float4 gml = ldg4(a.nl.gamma + c), bel = ldg4(a.nl.beta + c);
        float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), se = sg;
        for (int t = t0; t < t1; ++t)
            for (int f = f0; f < f1; ++f) {
                const long long o = ((((long long)b * a.T + t) * a.F) + f) * 64 + c;
                const float4 d = ldg4(dout + o);
                const float4 l = gln4(ldg4(a.l_pre + o), ml, rl, gml, bel);
                sg.x = fmaf(d.x, l.x, sg.x);
                sg.y = fmaf(d.y, l.y, sg.y);
                sg.z = fmaf(d.z, l.z, sg.z);
                sg.w = fmaf(d.w, l.w, sg.w);
                se.x += d.x;
                se.y += d.y;
                se.z += d.z;
                se.w += d.w;
            }
        const float4 g = gln4(ldg4(a.g_pre + idx * 4), mg, rg, ldg4(a.ng.gamma + c), ldg4(a.ng.beta + c));
        const float s0 = sigmoid_exact(g.x), s1 = sigmoid_exact(g.y), s2 = sigmoid_exact(g.z), s3 = sigmoid_exact(g.w);
        *reinterpret_cast<float4*>(dg + idx * 4) = make_float4(sg.x * s0 * (1.f - s0), sg.y * s1 * (1.f - s1), sg.z * s2 * (1.f - s2), sg.w * s3 * (1.f - s3));
        *reinterpret_cast<float4*>(de + idx * 4) = se;
    }
}

// ------------------------------------------------------------------------------------------------------ pooled sum
// g0 = gLN(d1_pre) + adaptive_avg_pool2d(d0, (Tc,Fc))   (tdanet.py:117-118).  Backward of the pool term, accumulated into the
// gradient of d0: window i of the adaptive pool covers [floor(i*T/Tc), ceil((i+1)*T/Tc)).
__global__ void __launch_bounds__(256) pool_bwd_kernel(const float* __restrict__ dg0, float* dd0, int T, int F, int Tc, int Fc, long long total4) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total4; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx & 15) * 4;
        const int pos = (int)(idx >> 4);
        const int f = pos % F;
        const int bt = pos / F;
        const int t = bt % T;
        const int b = bt / T;
        // windows that contain t: i in {ic - 1, ic, ic + 1} with ic = floor(t * Tc / T); 32-bit arithmetic (T, F < 65536)
        const int ic = (int)(((unsigned)t * (unsigned)Tc) / (unsigned)T), jc = (int)(((unsigned)f * (unsigned)Fc) / (unsigned)F);
        int iw[3], jw[3];
        float wt[3], wf[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int i = ic - 1 + k, j = jc - 1 + k;
            iw[k] = -1;
            jw[k] = -1;
            wt[k] = wf[k] = 0.f;
            if (i >= 0 && i < Tc) {
                const int ts = (int)(((unsigned)i * (unsigned)T) / (unsigned)Tc), te = (int)(((unsigned)(i + 1) * (unsigned)T + Tc - 1) / (unsigned)Tc);
                if (t >= ts && t < te) {
                    iw[k] = i;
                    wt[k] = 1.f / (float)(te - ts);
                }
            }
            if (j >= 0 && j < Fc) {
                const int fs = (int)(((unsigned)j * (unsigned)F) / (unsigned)Fc), fe = (int)(((unsigned)(j + 1) * (unsigned)F + Fc - 1) / (unsigned)Fc);
                if (f >= fs && f < fe) {
                    jw[k] = j;
                    wf[k] = 1.f / (float)(fe - fs);
                }
            }
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* base = dg0 + (long long)b * Tc * Fc * 64 + c;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (iw[a] < 0) continue;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (jw[k] < 0) continue;
                const float w = wt[a] * wf[k];
                const float4 d = ldg4(base + (iw[a] * Fc + jw[k]) * 64);
                acc.x = fmaf(d.x, w, acc.x);
                acc.y = fmaf(d.y, w, acc.y);
                acc.z = fmaf(d.z, w, acc.z);
                acc.w = fmaf(d.w, w, acc.w);
            }
        }
        float4* o = reinterpret_cast<float4*>(dd0) + idx;
        const float4 p = *o;
        *o = make_float4(p.x + acc.x, p.y + acc.y, p.z + acc.z, p.w + acc.w);
    }
}

// ------------------------------------------------------------------------------------------------------ small helpers
// out[c] += sum over rows of x[row][c]   (bias gradients of the 1x1 convs)
template <int C>
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long rows, float* out) {
    __shared__ float sh[C];
    for (int i = threadIdx.x; i < C; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    constexpr int Q = C / 4;             // float4 per row
    constexpr int RPB = 256 / Q > 0 ? 256 / Q : 1;
    const int q = threadIdx.x % Q, rl = threadIdx.x / Q;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl < RPB) {
        for (long long r = (long long)blockIdx.x * RPB + rl; r < rows; r += (long long)gridDim.x * RPB) {
            const float4 v = ldg4(x + r * C + q * 4);
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
        }
        atomicAdd(sh + q * 4, acc.x);
        atomicAdd(sh + q * 4 + 1, acc.y);
        atomicAdd(sh + q * 4 + 2, acc.z);
        atomicAdd(sh + q * 4 + 3, acc.w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(out + i, sh[i]);
}

// y = a + b (float4 streams; y may alias a)
__global__ void __launch_bounds__(256) add_kernel(const float* a, const float* __restrict__ b, float* y, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 u = reinterpret_cast<const float4*>(a)[i], v = __ldg(reinterpret_cast<const float4*>(b) + i);
        reinterpret_cast<float4*>(y)[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
    }
}

}  // namespace rtfs
