// Device side of the input pipeline (reference: src/datas/avspeech_dataset.py:10-14,115-143 and the lip-ROI transforms
// src/datas/transform.py:63-167): what the reference's DataLoader workers do in numpy per utterance runs here per batch,
// so that the host only moves the raw uint8 mouth ROIs and fp32 waveforms (pinned memory -> one async copy each).
//
//   mouth_preprocess_kernel : Normalize(0, 255) -> CenterCrop / RandomCrop(crop) -> HorizontalFlip -> Normalize(mean, std):
//                             roi (B, T, H, W) uint8 -> out (B, 1, T, crop, crop) fp32; the crop offsets and the flip decision are
//                             per utterance (drawn on the host in the reference's order), NULL = centre crop, no flip
//   wav_normalize_kernel    : normalize_tensor_wav: mixture -> (x - mean) / (std + eps) with the UNBIASED std of the mixture,
//                             every source -> (s - mean_s) / (std_mixture + eps)
// Both are pure streaming kernels: 16-byte stores, coalesced byte loads, fp64 accumulation of the moments.
#pragma once
#include "common.cuh"

namespace rtfs {

struct MouthPrepArgs {
    const unsigned char* roi;  // (B, T, H, W)
    float* out;                // (B, 1, T, crop, crop)
    const int* off_y;          // [B] or null (centre)
    const int* off_x;
    const int* flip;           // [B] or null
    int B, T, H, W, crop;
    float mean, std;
};

// one thread = 4 consecutive output pixels of one row (crop % 4 == 0)
__global__ void __launch_bounds__(256) mouth_preprocess_kernel(MouthPrepArgs a) {
    const int q = a.crop >> 2;
    const long long total = (long long)a.B * a.T * a.crop * q;
    // the reference divides in float64 and the result is cast to the waveform's dtype afterwards (core.py:89 mouth.type_as(wav));
    // fp32 with one correctly rounded division per step stays within 1 ulp of that
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int xq = (int)(i % q);
        long long r = i / q;
        const int y = (int)(r % a.crop);
        r /= a.crop;
        const int t = (int)(r % a.T), b = (int)(r / a.T);
        // CenterCrop: delta = int(round(w - tw) / 2.0) (transform.py:98-99)
        const int dy = a.off_y ? a.off_y[b] : (a.H - a.crop) / 2;
        const int dx = a.off_x ? a.off_x[b] : (a.W - a.crop) / 2;
        const bool fl = a.flip != nullptr && a.flip[b] != 0;
        const unsigned char* src = a.roi + (((long long)b * a.T + t) * a.H + (dy + y)) * a.W + dx;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = 4 * xq + k;
            const int xs = fl ? a.crop - 1 - x : x;  // cv2.flip(frame, 1) of the cropped frame
            const float u = __fdiv_rn((float)src[xs], 255.f);
            v[k] = __fdiv_rn(u - a.mean, a.std);
        }
        *reinterpret_cast<float4*>(a.out + (((long long)b * a.T + t) * a.crop + y) * a.crop + 4 * xq) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

struct WavNormArgs {
    const float* mix;  // (B, L)
    const float* src;  // (B, n_src, L) or null
    float* mix_out;
    float* src_out;
    int B, L, n_src;
    float eps;
};

DEVINL double block_sum_f64(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;
}

// grid (1 + n_src, B): row 0 = the mixture, row j = source j-1; every CTA recomputes the mixture's std (L floats from L2)
__global__ void __launch_bounds__(512) wav_normalize_kernel(WavNormArgs a) {
    __shared__ double red[16];
    const int b = blockIdx.y, row = blockIdx.x, tid = threadIdx.x;
    const float* mix = a.mix + (long long)b * a.L;
    double s = 0.0;
    for (int i = tid; i < a.L; i += blockDim.x) s += (double)mix[i];
    const double mean_m = block_sum_f64(s, red) / (double)a.L;
    double q = 0.0;
    for (int i = tid; i < a.L; i += blockDim.x) {
        const double d = (double)mix[i] - mean_m;
        q += d * d;
    }
    const double var = block_sum_f64(q, red) / (double)(a.L > 1 ? a.L - 1 : 1);  // torch.std: unbiased
    const float inv = 1.f / ((float)sqrt(var) + a.eps);
    const float* x = row == 0 ? mix : a.src + ((long long)b * a.n_src + (row - 1)) * a.L;
    float* y = row == 0 ? a.mix_out + (long long)b * a.L : a.src_out + ((long long)b * a.n_src + (row - 1)) * a.L;
    double mean_x = mean_m;
    if (row > 0) {
        double sx = 0.0;
        for (int i = tid; i < a.L; i += blockDim.x) sx += (double)x[i];
        mean_x = block_sum_f64(sx, red) / (double)a.L;
    }
    const float mx = (float)mean_x;
    for (int i = tid; i < a.L; i += blockDim.x) y[i] = (x[i] - mx) * inv;
}

}  // namespace rtfs
