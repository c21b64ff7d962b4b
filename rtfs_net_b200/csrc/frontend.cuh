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

}  // namespace rtfs
