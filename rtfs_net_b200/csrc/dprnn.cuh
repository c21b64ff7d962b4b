// Dual-path RNN pieces (reference: DualPathRNN.forward, layers/rnn_layers.py:136-162; the SRU
// recurrence follows the third-party `sru` package).
//   dprnn_prep : [gLN(d1_pre)+pool ->] g ; LayerNormalization4D over C ; write n sequence-major
//   (GEMMs)    : unfold(8) o Linear == GEMM over the overlapping row view of n (gemm.cuh, lda=64,K=512)
//   sru_scan   : element-wise recurrence, one thread per (sequence, direction, hidden unit)
//   (GEMM)     : ConvTranspose1d(64,64,8) == GEMM over the overlapping view of the zero-padded h
#pragma once
#include "common.cuh"

namespace rtfs {

struct PrepArgs {
    const float* g_in;    // (B,Tc,Fc,64) when first == 0
    const float* d1_pre;  // first == 1: g = gLN(d1_pre) + pool
    const float* pool;
    GlnRef gln;
    const float* ln_gamma;  // [64]
    const float* ln_beta;
    float* g_out;  // written when first == 1
    float* n_out;  // sequence-major normalised rows
    int B, Tc, Fc;
    int time_path;  // 0: seq=(b,t) s=f (same linear order) ; 1: seq=(b,f) s=t
    int first;
};

__global__ void __launch_bounds__(256) dprnn_prep_kernel(PrepArgs a) {
    const int l16 = threadIdx.x & 15;
    const long long pos = (long long)blockIdx.x * 16 + (threadIdx.x >> 4);
    const long long npos = (long long)a.B * a.Tc * a.Fc;
    if (pos >= npos) return;  // whole 16-lane groups exit together; shuffles below stay within a group
    const int c = l16 * 4;
    float4 v;
    if (a.first) {
        const int b = (int)(pos / ((long long)a.Tc * a.Fc));
        float mean, rstd;
        gln_mean_rstd(a.gln.sums, b, a.gln.inv_n, mean, rstd);
        const float4 x = ldg4(a.d1_pre + pos * 64 + c), p = ldg4(a.pool + pos * 64 + c);
        const float4 gm = ldg4(a.gln.gamma + c), be = ldg4(a.gln.beta + c);
        v.x = (x.x - mean) * rstd * gm.x + be.x + p.x;
        v.y = (x.y - mean) * rstd * gm.y + be.y + p.y;
        v.z = (x.z - mean) * rstd * gm.z + be.z + p.z;
        v.w = (x.w - mean) * rstd * gm.w + be.w + p.w;
        *reinterpret_cast<float4*>(a.g_out + pos * 64 + c) = v;
    } else {
        v = ldg4(a.g_in + pos * 64 + c);
    }
    const unsigned gmask = 0xFFFFu << (threadIdx.x & 16);
    float s = v.x + v.y + v.z + v.w;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(gmask, s, o);
    const float mu = s * (1.f / 64.f);
    const float dx = v.x - mu, dy = v.y - mu, dz = v.z - mu, dw = v.w - mu;
    float q = dx * dx + dy * dy + dz * dz + dw * dw;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(gmask, q, o);
    const float rs = 1.f / sqrtf(q * (1.f / 64.f) + RTFS_EPS);
    const float4 gm = ldg4(a.ln_gamma + c), be = ldg4(a.ln_beta + c);
    float4 n;
    n.x = dx * rs * gm.x + be.x;
    n.y = dy * rs * gm.y + be.y;
    n.z = dz * rs * gm.z + be.z;
    n.w = dw * rs * gm.w + be.w;
    long long npos_out = pos;
    if (a.time_path) {
        const int f = (int)(pos % a.Fc);
        const long long bt = pos / a.Fc;
        const int t = (int)(bt % a.Tc);
        const int b = (int)(bt / a.Tc);
        npos_out = ((long long)b * a.Fc + f) * a.Tc + t;
    }
    *reinterpret_cast<float4*>(a.n_out + npos_out * 64 + c) = n;
}

// One bidirectional SRU layer's recurrence.  U: rows seq*S + t, columns m*64 + col (m-major so a
// warp reads 128 contiguous bytes per gate), col = dir*32 + j.  k == 4: x' = U3 ; k == 3: x' = xin.
struct ScanArgs {
    const float* U;
    int ldu;            // 64*k
    const float* xin;   // previous layer output (k == 3), rows seq*S + t, 64 cols
    const float* wc;    // [128] = v_f | v_r
    const float* bias;  // [128] = b_f | b_r
    float* hout;        // rows seq*out_stride + out_off + t
    int nseq, S, L, k;
    int out_stride, out_off;
    int zero_pad;  // also write 7 zero rows before and after the L valid rows (conv-transpose input)
    float* cout = nullptr;  // training tape: the cell state c_t, rows seq*S + t (read again by sru_scan_bwd_kernel)
};

__global__ void __launch_bounds__(256) sru_scan_kernel(ScanArgs a) {
    const int col = threadIdx.x & 63;
    const int seq = blockIdx.x * 4 + (threadIdx.x >> 6);
    if (seq >= a.nseq) return;
    const bool rev = col >= 32;
    const float vf = __ldg(a.wc + col), vr = __ldg(a.wc + 64 + col);
    const float bf = __ldg(a.bias + col), br = __ldg(a.bias + 64 + col);
    const float* Ub = a.U + (long long)seq * a.S * a.ldu + col;
    const float* Xb = a.xin ? a.xin + (long long)seq * a.S * 64 + col : nullptr;
    float* Hb = a.hout + ((long long)seq * a.out_stride + a.out_off) * 64 + col;
    float* Cb = a.cout ? a.cout + (long long)seq * a.S * 64 + col : nullptr;
    if (a.zero_pad) {
        for (int i = 0; i < 7; ++i) {
            Hb[(long long)(i - 7) * 64] = 0.f;
            Hb[(long long)(a.L + i) * 64] = 0.f;
        }
    }
    const int L = a.L;
    const bool k4 = a.k == 4;
    float c = 0.f;
    constexpr int UN = 4;
    for (int s0 = 0; s0 < L; s0 += UN) {
        float u0[UN], u1[UN], u2[UN], xp[UN];
#pragma unroll
        for (int i = 0; i < UN; ++i) {
            const int s = s0 + i;
            const int t = rev ? (L - 1 - s) : s;
            if (s < L) {
                const float* p = Ub + (long long)t * a.ldu;
                u0[i] = __ldg(p);
                u1[i] = __ldg(p + 64);
                u2[i] = __ldg(p + 128);
                xp[i] = k4 ? __ldg(p + 192) : __ldg(Xb + (long long)t * 64);
            } else {
                u0[i] = u1[i] = u2[i] = xp[i] = 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < UN; ++i) {
            const int s = s0 + i;
            if (s < L) {
                const int t = rev ? (L - 1 - s) : s;
                const float f = sigmoidf_fast(u1[i] + vf * c + bf);
                const float r = sigmoidf_fast(u2[i] + vr * c + br);
                c = f * c + (1.f - f) * u0[i];
                Hb[(long long)t * 64] = r * c + (1.f - r) * xp[i];
                if (Cb) Cb[(long long)t * 64] = c;
            }
        }
    }
}

}  // namespace rtfs
