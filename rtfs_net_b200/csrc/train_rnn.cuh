// Backward / training-step kernels, part 3: the dual-path RNN (DualPathRNN.forward, layers/rnn_layers.py:136-162; SRU
// recurrence as restated in SURVEY.md App. C -- the upstream package keeps c for its own backward kernel the same way).
//
// Training forward (api.cu run_dprnn_train) = the unfused kernel chain with a tape: n (LN output, sequence-major), per
// layer U_l (gate pre-activations, rows seq*S + t), c_l (cell states) and h_l (layer outputs; the last one zero-padded for
// the ConvTranspose1d GEMM).  Backward of one path:
//   dz (B,Tc,Fc,64) --seq_pad--> dzp [seq][S+7][64]       (zero rows s >= S: the transposed-conv GEMM view reads past S)
//   dh_3 = unfold-GEMM(dzp) ; dW_ct, db_ct                    (train_gemm.cuh)
//   l = 3..0: sru_scan_bwd(U_l, c_l, dh_l) -> dU_l, d(highway input), dv, db ; dh_{l-1} = dU_l W_l^T + highway ; dW_l
//   dXunf = dU_0 W_0^T  [rows seq*S + l][tap*64 + c]
//   dprnn_ln_bwd: fold the 8 taps, LayerNorm-over-C backward, + residual path  -> dg_in ; dgamma, dbeta
#pragma once
#include "dprnn.cuh"

namespace rtfs {

// natural layout (B,Tc,Fc,64) -> sequence-major rows [seq][S+7][64]; rows S..S+6 of every sequence are zero
//   frequency path: seq = (b,t), s = f ; time path: seq = (b,f), s = t
__global__ void __launch_bounds__(256) seq_pad_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int Tc, int Fc, int time_path) {
    const int S = time_path ? Tc : Fc, n_other = time_path ? Fc : Tc;
    const long long total = (long long)B * n_other * (S + 7) * 16;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx & 15) * 4;
        const long long row = idx >> 4;
        const int s = (int)(row % (S + 7));
        const long long seq = row / (S + 7);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < S) {
            const int o = (int)(seq % n_other);
            const long long b = seq / n_other;
            const int t = time_path ? s : o, f = time_path ? o : s;
            v = ldg4(x + (((b * Tc + t) * Fc) + f) * 64 + c);
        }
        *reinterpret_cast<float4*>(out + row * 64 + c) = v;
    }
}

// Reverse-time sweep of one bidirectional SRU layer.  Thread = (sequence, column) as in sru_scan_kernel.
//   forward:  f = s(u1 + vf c' + bf) ; r = s(u2 + vr c' + br) ; c = f c' + (1-f) u0 ; h = r c + (1-r) x'      (c' = c_{prev})
//   backward: dr = dh (c - x') ; dx' = dh (1-r) ; dc += dh r ; df = dc (c' - u0) ; du0 = dc (1-f) ;
//             du2 = dr r(1-r) ; du1 = df f(1-f) ; dc' = dc f + du2 vr + du1 vf ; dv += du c' ; db += du
struct ScanBwdArgs {
    const float* U;     // rows seq*S + t, ldu = 64*k, columns m*64 + col
    int ldu;
    const float* C;     // cell states, rows seq*S + t
    const float* xin;   // k == 3: layer input h_{l-1}, rows seq*S + t
    const float* dh;    // rows seq*dh_stride + t
    int dh_stride;
    const float* wc;
    const float* bias;
    float* dU;          // rows seq*S + t (rows t >= L are zeroed: the weight-gradient GEMMs run over all S rows)
    float* dxin;        // k == 3: (1-r)*dh, rows seq*S + t, rows t >= L zeroed
    float* dwc;         // [128] accumulated
    float* dbias;       // [128] accumulated
    int nseq, S, L, k;
};

__global__ void __launch_bounds__(256) sru_scan_bwd_kernel(ScanBwdArgs a) {
    __shared__ float sh[4 * 64];
    sh[threadIdx.x] = 0.f;
    __syncthreads();
    const int col = threadIdx.x & 63;
    const int seq = blockIdx.x * 4 + (threadIdx.x >> 6);
    float dvf = 0.f, dvr = 0.f, dbf = 0.f, dbr = 0.f;
    if (seq < a.nseq) {
        const bool rev = col >= 32;
        const bool k4 = a.k == 4;
        const float vf = __ldg(a.wc + col), vr = __ldg(a.wc + 64 + col);
        const float bf = __ldg(a.bias + col), br = __ldg(a.bias + 64 + col);
        const float* Ub = a.U + (long long)seq * a.S * a.ldu + col;
        const float* Cb = a.C + (long long)seq * a.S * 64 + col;
        const float* Xb = k4 ? nullptr : a.xin + (long long)seq * a.S * 64 + col;
        const float* Db = a.dh + (long long)seq * a.dh_stride * 64 + col;
        float* dUb = a.dU + (long long)seq * a.S * a.ldu + col;
        float* dXb = k4 ? nullptr : a.dxin + (long long)seq * a.S * 64 + col;
        const int L = a.L;
        for (int t = L; t < a.S; ++t) {  // rows past the last unfolded step
            float* p = dUb + (long long)t * a.ldu;
            p[0] = 0.f;
            p[64] = 0.f;
            p[128] = 0.f;
            if (k4) p[192] = 0.f;
            else dXb[(long long)t * 64] = 0.f;
        }
        float dc = 0.f;
        constexpr int UN = 4;
        // scan step s (forward direction: t = s ; reverse direction: t = L-1-s); walk s = L-1 .. 0
        for (int s0 = L - 1; s0 >= 0; s0 -= UN) {
            float u0[UN], u1[UN], u2[UN], xp[UN], cc[UN], cp[UN], dh[UN];
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                const int s = s0 - i;
                if (s >= 0) {
                    const int t = rev ? (L - 1 - s) : s;
                    const int tp = rev ? t + 1 : t - 1;  // time index of scan step s-1
                    const float* p = Ub + (long long)t * a.ldu;
                    u0[i] = __ldg(p);
                    u1[i] = __ldg(p + 64);
                    u2[i] = __ldg(p + 128);
                    xp[i] = k4 ? __ldg(p + 192) : __ldg(Xb + (long long)t * 64);
                    cc[i] = __ldg(Cb + (long long)t * 64);
                    cp[i] = s > 0 ? __ldg(Cb + (long long)tp * 64) : 0.f;
                    dh[i] = __ldg(Db + (long long)t * 64);
                } else {
                    u0[i] = u1[i] = u2[i] = xp[i] = cc[i] = cp[i] = dh[i] = 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                const int s = s0 - i;
                if (s >= 0) {
                    const int t = rev ? (L - 1 - s) : s;
                    const float f = sigmoidf_fast(u1[i] + vf * cp[i] + bf);
                    const float r = sigmoidf_fast(u2[i] + vr * cp[i] + br);
                    const float dr = dh[i] * (cc[i] - xp[i]);
                    const float dxp = dh[i] * (1.f - r);
                    const float dct = dc + dh[i] * r;
                    const float df = dct * (cp[i] - u0[i]);
                    const float du0 = dct * (1.f - f);
                    const float du2 = dr * r * (1.f - r);
                    const float du1 = df * f * (1.f - f);
                    dc = dct * f + du2 * vr + du1 * vf;
                    dvf += du1 * cp[i];
                    dvr += du2 * cp[i];
                    dbf += du1;
                    dbr += du2;
                    float* p = dUb + (long long)t * a.ldu;
                    p[0] = du0;
                    p[64] = du1;
                    p[128] = du2;
                    if (k4) p[192] = dxp;
                    else dXb[(long long)t * 64] = dxp;
                }
            }
        }
    }
    // the four sequences of the CTA -> one atomic per column and quantity
    atomicAdd(sh + col, dvf);
    atomicAdd(sh + 64 + col, dvr);
    atomicAdd(sh + 128 + col, dbf);
    atomicAdd(sh + 192 + col, dbr);
    __syncthreads();
    if (threadIdx.x < 128) {
        atomicAdd(a.dwc + threadIdx.x, sh[threadIdx.x]);
        atomicAdd(a.dbias + threadIdx.x, sh[128 + threadIdx.x]);
    }
}

// Fold of the unfold(8) gradient + LayerNormalization4D((C,1)) backward + residual path (rnn_layers.py:145-147,156).
//   dn[seq][s][c] = sum_{k=0..7, 0 <= s-k < L} dXunf[seq*S + s-k][k*64 + c]
//   dz = (gamma*dn - mean_c(gamma*dn) - xhat*mean_c(gamma*dn*xhat)) * rstd + dout       (out = z' + z)
struct LnBwdArgs {
    const float* dxunf;  // [nseq*S][512]
    const float* z;      // layer input, natural layout (B,Tc,Fc,64)
    const float* dout;   // gradient w.r.t. the DualPathRNN output, natural layout
    const float* gamma;
    float* dz;           // natural layout
    float* dgamma;       // [64] accumulated
    float* dbeta;
    int B, Tc, Fc, time_path, L;
};

__global__ void __launch_bounds__(256) dprnn_ln_bwd_kernel(LnBwdArgs a) {
    __shared__ float sh[128];
    if (threadIdx.x < 128) sh[threadIdx.x] = 0.f;
    __syncthreads();
    const int l16 = threadIdx.x & 15, c = l16 * 4;
    const long long npos = (long long)a.B * a.Tc * a.Fc;
    const int S = a.time_path ? a.Tc : a.Fc;
    const unsigned gmask = 0xFFFFu << (threadIdx.x & 16);
    const float4 gm = ldg4(a.gamma + c);
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
    // every 16-lane group walks its own positions; the loop bound is rounded up so that both groups of a warp stay converged
    const long long iters = (npos + (long long)gridDim.x * 16 - 1) / ((long long)gridDim.x * 16);
    for (long long it = 0; it < iters; ++it) {
        const long long pos = (it * gridDim.x + blockIdx.x) * 16 + (threadIdx.x >> 4);
        const bool valid = pos < npos;
        float4 dn = make_float4(0.f, 0.f, 0.f, 0.f), v = dn, d_o = dn;
        if (valid) {
            const int f = (int)(pos % a.Fc);
            const long long bt = pos / a.Fc;
            const int t = (int)(bt % a.Tc);
            const long long b = bt / a.Tc;
            const long long seq = a.time_path ? b * a.Fc + f : bt;
            const int s = a.time_path ? t : f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int l = s - k;
                if (l >= 0 && l < a.L) {
                    const float4 u = ldg4(a.dxunf + (seq * S + l) * 512 + k * 64 + c);
                    dn.x += u.x;
                    dn.y += u.y;
                    dn.z += u.z;
                    dn.w += u.w;
                }
            }
            v = ldg4(a.z + pos * 64 + c);
            d_o = ldg4(a.dout + pos * 64 + c);
        }
        float sm = v.x + v.y + v.z + v.w;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sm += __shfl_xor_sync(gmask, sm, o);
        const float mu = sm * (1.f / 64.f);
        const float x0 = v.x - mu, x1 = v.y - mu, x2 = v.z - mu, x3 = v.w - mu;
        float q = x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(gmask, q, o);
        const float rs = 1.f / sqrtf(q * (1.f / 64.f) + RTFS_EPS);
        const float h0 = x0 * rs, h1 = x1 * rs, h2 = x2 * rs, h3 = x3 * rs;
        const float w0 = gm.x * dn.x, w1 = gm.y * dn.y, w2 = gm.z * dn.z, w3 = gm.w * dn.w;
        float s1 = w0 + w1 + w2 + w3, s2 = w0 * h0 + w1 * h1 + w2 * h2 + w3 * h3;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(gmask, s1, o);
            s2 += __shfl_xor_sync(gmask, s2, o);
        }
        s1 *= (1.f / 64.f);
        s2 *= (1.f / 64.f);
        if (valid) {
            float4 o;
            o.x = (w0 - s1 - h0 * s2) * rs + d_o.x;
            o.y = (w1 - s1 - h1 * s2) * rs + d_o.y;
            o.z = (w2 - s1 - h2 * s2) * rs + d_o.z;
            o.w = (w3 - s1 - h3 * s2) * rs + d_o.w;
            *reinterpret_cast<float4*>(a.dz + pos * 64 + c) = o;
            dg.x += dn.x * h0;
            dg.y += dn.y * h1;
            dg.z += dn.z * h2;
            dg.w += dn.w * h3;
            db.x += dn.x;
            db.y += dn.y;
            db.z += dn.z;
            db.w += dn.w;
        }
    }
    atomicAdd(sh + c, dg.x);
    atomicAdd(sh + c + 1, dg.y);
    atomicAdd(sh + c + 2, dg.z);
    atomicAdd(sh + c + 3, dg.w);
    atomicAdd(sh + 64 + c, db.x);
    atomicAdd(sh + 64 + c + 1, db.y);
    atomicAdd(sh + 64 + c + 2, db.z);
    atomicAdd(sh + 64 + c + 3, db.w);
    __syncthreads();
    if (threadIdx.x < 64) {
        atomicAdd(a.dgamma + threadIdx.x, sh[threadIdx.x]);
        atomicAdd(a.dbeta + threadIdx.x, sh[64 + threadIdx.x]);
    }
}

}  // namespace rtfs
