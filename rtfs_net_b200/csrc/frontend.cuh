// STFT front end and iSTFT back end of the path, fp32 (the reference runs cuFFT in fp32; only
// the convolutions/matmuls are TF32 there), so the DFTs are done with fp32 FMAs against an
// exact 256-entry twiddle table.
//   stft_kernel      : torch.stft(n_fft=256, hop=128, hann periodic, center/reflect, onesided)
//                      -> spec (B,T,F,2)                     (TDAVNet/encoder.py:161-172)
//   dec_istft_kernel : sums the 9 shifted partial products of the 3x3 transposed conv
//                      (ConvTranspose2d 256->2, decoder.py:96-116, computed as a K=256,N=18 GEMM),
//                      irDFT, window, overlap-add, /sum(window^2), trim  (decoder.py:122-128)
#pragma once
#include "common.cuh"

namespace rtfs {

struct StftArgs {
    const float* wav;     // (B,L)
    const float* window;  // [256] hann periodic, fp32
    const float* costab;  // [256] cos(2*pi*k/256)
    const float* sintab;  // [256] sin(2*pi*k/256)
    float* spec;          // (B,T,129,2)
    int L, T;
};

__global__ void __launch_bounds__(288) stft_kernel(StftArgs a) {
    __shared__ float fr[256], ct[256], st[256];
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    if (tid < 256) {
        int s = t * 128 + tid - 128;
        if (s < 0) s = -s;
        if (s >= a.L) s = 2 * (a.L - 1) - s;
        fr[tid] = __ldg(a.wav + (long long)b * a.L + s) * __ldg(a.window + tid);
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
        a.spec[(((long long)b * a.T + t) * 129 + f) * 2 + part] = part ? -acc : acc;
    }
}

struct IstftArgs {
    const float* q;       // (B,T,129,18): column o*9 + i*3 + j of the transposed-conv partials
    const float* window;  // [256]
    const float* costab;
    const float* sintab;
    float* out;  // (B,L)
    int L, T;
};

__global__ void __launch_bounds__(256) dec_istft_kernel(IstftArgs a) {
    __shared__ float yre[2][129], yim[2][129], ct[256], st[256], part0[128];
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int F = 129;
    ct[tid] = __ldg(a.costab + tid);
    st[tid] = __ldg(a.sintab + tid);
    // frames: sel 0 -> t0 = h (second half of the frame), sel 1 -> t1 = h + 1 (first half)
    for (int i = tid; i < 2 * 129 * 2; i += 256) {
        const int sel = i / 258, r = i - sel * 258;
        const int f = r >> 1, o = r & 1;
        const int t = h + sel;
        float acc = 0.f;
        if (t < a.T) {
#pragma unroll
            for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    const int ts = t + 1 - ii, fs = f + 1 - jj;
                    if (ts >= 0 && ts < a.T && fs >= 0 && fs < F)
                        acc += __ldg(a.q + (((long long)b * a.T + ts) * F + fs) * 18 + o * 9 + ii * 3 + jj);
                }
        }
        if (o == 0) yre[sel][f] = acc;
        else yim[sel][f] = acc;
    }
    __syncthreads();
    const int sel = tid >> 7, m = tid & 127;
    const int mm = sel == 0 ? m + 128 : m;  // sample index inside the frame
    const int t = h + sel;
    float x = 0.f;
    if (t < a.T) {
        float acc = yre[sel][0] + ((mm & 1) ? -yre[sel][128] : yre[sel][128]);
#pragma unroll 4
        for (int f = 1; f < 128; ++f) {
            const int idx = (f * mm) & 255;
            acc = fmaf(2.f * yre[sel][f], ct[idx], acc);
            acc = fmaf(-2.f * yim[sel][f], st[idx], acc);
        }
        x = acc * (1.f / 256.f) * __ldg(a.window + mm);
    }
    if (sel == 0) part0[m] = x;
    __syncthreads();
    if (sel == 1) {
        const int n = h * 128 + m;
        if (n < a.L) {
            const float w1 = __ldg(a.window + m), w0 = __ldg(a.window + m + 128);
            float env = 0.f;
            if (h + 1 < a.T) env += w1 * w1;
            env += w0 * w0;  // frame h always exists for n < L
            a.out[(long long)b * a.L + n] = (x + part0[m]) / env;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Second generation: 16 frames per CTA.  A frame is two 128-sample blocks (hop = n_fft / 2), so 16 frames share 17
// blocks; thread f keeps 16 (re, im) accumulators and walks m = 0..127 once, using cos/sin(2 pi (m + 128) f / 256) =
// (-1)^f cos/sin(2 pi m f / 256): per m it reads 17 block samples (broadcast float4 loads of the transposed block
// table) and 2 twiddles for 64 FMAs, instead of 2 conflicting shared loads per FMA in a 256-long dependent chain.
constexpr int FE_FR = 16;   // frames (hops) per CTA
constexpr int FE_LD = 20;   // floats per row of the transposed tables (17 used)

__global__ void __launch_bounds__(160) stft16_kernel(StftArgs a) {
    __shared__ __align__(16) float bT[128 * FE_LD];  // bT[m][blk] = sample m of block t0 - 1 + blk (reflect-padded)
    __shared__ float ct[256], st[256], wn[256];
    const int t0 = blockIdx.x * FE_FR, b = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < 256; i += 160) {
        ct[i] = __ldg(a.costab + i);
        st[i] = __ldg(a.sintab + i);
        wn[i] = __ldg(a.window + i);
    }
    for (int i = tid; i < 17 * 128; i += 160) {
        const int blk = i >> 7, m = i & 127;
        int s = (t0 - 1 + blk) * 128 + m;
        if (s < 0) s = -s;
        if (s >= a.L) s = 2 * (a.L - 1) - s;
        bT[m * FE_LD + blk] = (s >= 0 && s < a.L) ? __ldg(a.wav + (long long)b * a.L + s) : 0.f;
    }
    __syncthreads();
    const int f = tid;
    if (f > 128) return;
    const float sgn = (f & 1) ? -1.f : 1.f;
    float re[FE_FR], im[FE_FR];
#pragma unroll
    for (int i = 0; i < FE_FR; ++i) re[i] = im[i] = 0.f;
#pragma unroll 2
    for (int m = 0; m < 128; ++m) {
        const int idx = (m * f) & 255;
        const float c = ct[idx], sn = st[idx];
        const float w0 = wn[m], w1 = wn[m + 128] * sgn;
        const float c0 = c * w0, s0 = sn * w0, c1 = c * w1, s1 = sn * w1;
        float bl[FE_LD];
#pragma unroll
        for (int j = 0; j < FE_LD / 4; ++j) *reinterpret_cast<float4*>(bl + 4 * j) = *reinterpret_cast<const float4*>(bT + m * FE_LD + 4 * j);
#pragma unroll
        for (int i = 0; i < FE_FR; ++i) {
            re[i] = fmaf(bl[i], c0, fmaf(bl[i + 1], c1, re[i]));
            im[i] = fmaf(bl[i], s0, fmaf(bl[i + 1], s1, im[i]));
        }
    }
#pragma unroll
    for (int i = 0; i < FE_FR; ++i) {
        const int t = t0 + i;
        if (t < a.T) *reinterpret_cast<float2*>(a.spec + (((long long)b * a.T + t) * 129 + f) * 2) = make_float2(re[i], -im[i]);
    }
}

// iSTFT, 16 hops (17 frames) per CTA: (A) the nine shifted partial products of the transposed conv are summed into
// Y[f][frame] (transposed tables in shared memory), (B) thread n computes sample n of all 17 frames with one pass over
// f = 1..127 (2 twiddles + 10 broadcast float4 loads for 68 FMAs), (C) window, overlap-add of the two half frames,
// division by the window envelope.
__global__ void __launch_bounds__(256) dec_istft16_kernel(IstftArgs a) {
    __shared__ __align__(16) float yre[129 * FE_LD], yim[129 * FE_LD];  // [f][frame]
    __shared__ __align__(16) float ov[FE_FR * 128];                     // second halves of frames t0 .. t0+15
    __shared__ float ct[256], st[256];
    const int t0 = blockIdx.x * FE_FR, b = blockIdx.y, tid = threadIdx.x;
    const int F = 129;
    ct[tid] = __ldg(a.costab + tid);
    st[tid] = __ldg(a.sintab + tid);
    for (int i = tid; i < 17 * 129; i += 256) {
        const int fr = i / 129, f = i - fr * 129;
        const int t = t0 + fr;
        float ar = 0.f, ai = 0.f;
        if (t < a.T) {
#pragma unroll
            for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    const int ts = t + 1 - ii, fs = f + 1 - jj;
                    if (ts >= 0 && ts < a.T && fs >= 0 && fs < F) {
                        const float* qp = a.q + (((long long)b * a.T + ts) * F + fs) * 18 + ii * 3 + jj;
                        ar += __ldg(qp);
                        ai += __ldg(qp + 9);
                    }
                }
        }
        yre[f * FE_LD + fr] = ar;
        yim[f * FE_LD + fr] = ai;
    }
    __syncthreads();
    const int n = tid;
    float acc[17];
    {
        const float sg = (n & 1) ? -1.f : 1.f;
#pragma unroll
        for (int i = 0; i < 17; ++i) acc[i] = yre[i] + sg * yre[128 * FE_LD + i];
    }
#pragma unroll 2
    for (int f = 1; f < 128; ++f) {
        const int idx = (f * n) & 255;
        const float c2 = 2.f * ct[idx], s2 = -2.f * st[idx];
        float yr[FE_LD], yi[FE_LD];
#pragma unroll
        for (int j = 0; j < FE_LD / 4; ++j) {
            *reinterpret_cast<float4*>(yr + 4 * j) = *reinterpret_cast<const float4*>(yre + f * FE_LD + 4 * j);
            *reinterpret_cast<float4*>(yi + 4 * j) = *reinterpret_cast<const float4*>(yim + f * FE_LD + 4 * j);
        }
#pragma unroll
        for (int i = 0; i < 17; ++i) acc[i] = fmaf(yr[i], c2, fmaf(yi[i], s2, acc[i]));
    }
    const float w = __ldg(a.window + n) * (1.f / 256.f);
    if (n >= 128) {  // second half of frame t0 + i -> hop t0 + i
#pragma unroll
        for (int i = 0; i < FE_FR; ++i) ov[i * 128 + (n - 128)] = acc[i] * w;
    }
    __syncthreads();
    if (n < 128) {   // first half of frame t0 + i + 1 -> hop t0 + i
        const float w1 = __ldg(a.window + n), w0 = __ldg(a.window + n + 128);
#pragma unroll
        for (int i = 0; i < FE_FR; ++i) {
            const int h = t0 + i, smp = h * 128 + n;
            if (smp < a.L) {
                float env = w0 * w0;  // frame h always exists for smp < L
                float x = ov[i * 128 + n];
                if (h + 1 < a.T) {
                    env += w1 * w1;
                    x += acc[i + 1] * w;
                }
                a.out[(long long)b * a.L + smp] = x / env;
            }
        }
    }
}

}  // namespace rtfs
