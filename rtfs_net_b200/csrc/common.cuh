// Shared device helpers for the RTFS-Net forward kernels (sm_100a).
// Physical layout of every activation on the path is channels-last: a logical (B,C,T,F) tensor
// is stored (B,T,F,C) fp32, so that every contraction/recurrence/attention on the path
// (all of which run over C) sees contiguous rows of C floats.  See DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DEVINL __device__ __forceinline__
#define RTFS_EPS 1e-5f

// ------------------------------------------------------------------ per-device launch configuration
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE attribute of a kernel: a process that runs models on
// several devices must set it on each.  One cache per kernel instantiation (a `static SmemCfg` in its launch helper)
// remembers the largest size configured on every device; racing threads at worst set the same attribute twice.
namespace rtfs {
constexpr int kMaxDevices = 64;
struct SmemCfg {
    int bytes[kMaxDevices] = {};
};
template <class K>
inline cudaError_t ensure_smem(K kern, int smem, SmemCfg& cfg) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const bool cached = dev >= 0 && dev < kMaxDevices;
    if (cached && cfg.bytes[dev] >= smem && cfg.bytes[dev] > 0) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess && cached) cfg.bytes[dev] = smem;
    return e;
}
// SM count of the current device (148 on B200), cached per device; grids of the persistent kernels are sized from it
inline int sm_count() {
    static int cache[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < kMaxDevices && cache[dev] > 0) return cache[dev];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev >= 0 && dev < kMaxDevices) cache[dev] = n;
    return n;
}
}  // namespace rtfs

// ------------------------------------------------------------------ TF32 tensor-core MMA (legacy path)
// round-to-nearest (ties away) to TF32.  cvt.rna.tf32.f32 is emulated on sm_100a with three instructions (finite test +
// IADD + LOP3); the two-instruction integer form gives the same bits for every finite input and +-inf.
DEVINL uint32_t f2tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
DEVINL uint32_t f2tf32_cvt(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
DEVINL float tf32r(float x) { return __uint_as_float(f2tf32(x)); }
// cvt.rna.tf32.f32 is emulated on sm_100a (FSETP finite test + IADD + LOP3); without the NaN-payload guard it is two
// integer instructions with the same result for every finite input and +-inf (round half away: add half an ulp of the
// 10-bit mantissa to the magnitude bits, truncate)
DEVINL float tf32r_fast(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

// D(16x8) += A(16x8,row) * B(8x8,col);  lane = 4*g + t
//   a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4);  b0=(k=t,n=g) b1=(k=t+4,n=g)
//   d0=(g,2t) d1=(g,2t+1) d2=(g+8,2t) d3=(g+8,2t+1)
DEVINL void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ------------------------------------------------------------------ cp.async
DEVINL void cp_async16(void* smem, const void* gmem, bool valid) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
DEVINL void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// ------------------------------------------------------------------ reductions
DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-level (sum, sumsq) -> one fp64 atomicAdd pair.  `scratch` holds >= 2*nwarps floats.
// All threads of the block must call.  dst == nullptr skips the atomic.
DEVINL void block_stats_atomic(float s, float ss, double* dst, float* scratch) {
    s = warp_sum(s);
    ss = warp_sum(ss);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) {
        scratch[2 * w] = s;
        scratch[2 * w + 1] = ss;
    }
    __syncthreads();
    if (w == 0) {
        float a = lane < nw ? scratch[2 * lane] : 0.f;
        float b = lane < nw ? scratch[2 * lane + 1] : 0.f;
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0 && dst != nullptr) {
            atomicAdd(dst, (double)a);
            atomicAdd(dst + 1, (double)b);
        }
    }
}

// gLN statistics: sums[2*b] = sum, sums[2*b+1] = sum of squares over the n elements of sample b.
DEVINL void gln_mean_rstd(const double* __restrict__ sums, int b, double inv_n, float& mean, float& rstd) {
    const double s = sums[2 * b], ss = sums[2 * b + 1];
    const double m = s * inv_n;
    double var = ss * inv_n - m * m;
    var = var < 0.0 ? 0.0 : var;
    mean = (float)m;
    rstd = (float)rsqrt(var + 1e-5);
}

// A gLN to be applied on load: y = x * sc[c] + sh[c] with sc = rstd*gamma, sh = beta - mean*rstd*gamma
struct GlnRef {
    const double* sums;   // [B][2]
    const float* gamma;   // [C]
    const float* beta;    // [C]
    double inv_n;
};

DEVINL float prelu(float x, float a) { return x >= 0.f ? x : a * x; }
DEVINL float sigmoidf_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// F.interpolate(mode='nearest'): src = min(floor(dst*in/out), in-1)
DEVINL int nearest_src(int dst, int n_in, int n_out) {
    int s = (int)(((long long)dst * n_in) / n_out);
    return s < n_in - 1 ? s : n_in - 1;
}
// same, 32-bit arithmetic: exact while dst * n_in < 2^32 (every (T, F) this library accepts: both < 65536)
DEVINL int nearest_src32(int dst, int n_in, int n_out) {
    const int s = (int)(((unsigned)dst * (unsigned)n_in) / (unsigned)n_out);
    return s < n_in - 1 ? s : n_in - 1;
}

DEVINL float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
DEVINL float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
