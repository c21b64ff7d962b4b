// Backward / training-step kernels, part 2: contractions.
//   * data gradients of the 1x1 convs / SRU input maps / ConvTranspose1d are row-tile GEMMs dX = dY * W: they run on the
//     row-tile kernel of gemm.cuh (launch_gemm) with transposed weight images and the epilogues below;
//   * weight gradients dW[n][k] = sum_rows dY[row][n] * X[row][k] (reduction over ~10^6 rows into a <= 256 x 512 matrix)
//     run on wgrad_kernel: 64 x 64 output tiles, rows split over CTAs, fp32 atomics into the caller-zeroed gradient.
//     X and dY come through the same loader functors as the forward GEMMs, so the transformed operand of a fused forward
//     prologue (gLN+ReLU of the bottleneck, gateway+PReLU of the projection, PReLU of the mask head, im2col of the encoder,
//     the overlapping unfold views of the dual-path RNN) is re-formed on load instead of being stored in the forward;
//   * bgemm_kernel: batched small fp32 GEMM with arbitrary strides (attention score / context gradients per (b, head)).
#pragma once
#include <cstdlib>

#include "gemm.cuh"

namespace rtfs {

// ------------------------------------------------------------------------------------------------------------ wgrad
constexpr int WG_LD = 72;  // 64 + 8: fragment reads of 8 consecutive columns x 4 rows hit 32 distinct banks

// CTA tile: 64 (n) x 64*KT (k) outputs, rows streamed in 32-row sub-chunks of 128-row groups (the loaders' tile contract).
//   KT = 1: 8 warps = 4 (n, 16 rows) x 2 (k, 32 columns)           -- the K = 64 shapes
//   KT = 4: 8 warps = 2 (n, 32 rows) x 4 (k, 64 columns)           -- K >= 256: dY is re-read K/256 instead of K/64 times
// Operands are rounded to TF32 when they are staged (one rounding per element, none in the MMA loop).  The products of a
// reduction over 10^4..10^6 rows are individually rounded to nearest, so their errors average out: plain TF32 is far below the
// TF32 noise the forward activations already carry.  PREC3 (3xTF32: hi/lo planes of both operands staged separately) is kept
// for the tiny-K encoder / decoder filters that touch the waveform directly, and as an A/B switch (RTFS_WGRAD_3X=1).
template <bool PREC3, int KT, class XL, class YL>
__global__ void __launch_bounds__(256) wgrad_kernel(XL xl, YL yl, float* __restrict__ dW, int ldw, int M, int N, int K, int groups_per_cta) {
    constexpr int LDX = 64 * KT + 8;
    constexpr int NP = PREC3 ? 2 : 1;           // operand planes: hi (, lo)
    constexpr int WM = KT == 1 ? 1 : 2;         // 16-row n tiles per warp
    constexpr int WN = KT == 1 ? 4 : 8;         // 8-column k tiles per warp
    extern __shared__ __align__(16) float smem[];
    float* Ys = smem;                            // [NP][32][WG_LD]
    float* Xs = Ys + NP * 32 * WG_LD;            // [NP][32][LDX]
    float* extra_x = Xs + NP * 32 * LDX;
    float* extra_y = extra_x + XL::kExtra;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int ntn = (N + 63) / 64;
    const int n0 = (blockIdx.x % ntn) * 64, k0 = (blockIdx.x / ntn) * (64 * KT);
    const int wn0 = KT == 1 ? (warp & 3) * 16 : (warp & 1) * 32;   // first n row of the warp
    const int wk0 = KT == 1 ? (warp >> 2) * 32 : (warp >> 1) * 64; // first k column of the warp
    float acc[WM][WN][4];
#pragma unroll
    for (int a = 0; a < WM; ++a)
#pragma unroll
        for (int b = 0; b < WN; ++b) acc[a][b][0] = acc[a][b][1] = acc[a][b][2] = acc[a][b][3] = 0.f;

    const int ngroups = (M + 127) / 128;
    const int g_begin = blockIdx.y * groups_per_cta;
    const int g_end = min(ngroups, g_begin + groups_per_cta);
    const int r_ld = tid >> 3, c_ld = (tid & 7) * 4;
    auto stage = [&](float* dst, int ld, int col, float4 v) {
        const float4 h = make_float4(tf32r(v.x), tf32r(v.y), tf32r(v.z), tf32r(v.w));
        *reinterpret_cast<float4*>(dst + r_ld * ld + col) = h;
        if (PREC3)
            *reinterpret_cast<float4*>(dst + 32 * ld + r_ld * ld + col) =
                make_float4(tf32r(v.x - h.x), tf32r(v.y - h.y), tf32r(v.z - h.z), tf32r(v.w - h.w));
    };
    for (int grp = g_begin; grp < g_end; ++grp) {
        __syncthreads();  // previous group's tables / tiles are no longer read
        xl.init(grp * 128, M, extra_x);
        yl.init(grp * 128, M, extra_y);
        __syncthreads();
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            if (grp * 128 + 32 * i >= M) break;
            // 32 rows x (64*KT | 64) columns: thread -> row r_ld, columns c_ld + 32*j
            float4 xv[2 * KT], yv[2];
#pragma unroll
            for (int j = 0; j < 2 * KT; ++j)
                xv[j] = (k0 + c_ld + 32 * j < K) ? xl.load(i, k0 + c_ld + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 2; ++j)
                yv[j] = (n0 + c_ld + 32 * j < N) ? yl.load(i, n0 + c_ld + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
            __syncthreads();  // the previous sub-chunk's MMAs are done
#pragma unroll
            for (int j = 0; j < 2 * KT; ++j) stage(Xs, LDX, c_ld + 32 * j, xv[j]);
#pragma unroll
            for (int j = 0; j < 2; ++j) stage(Ys, WG_LD, c_ld + 32 * j, yv[j]);
            __syncthreads();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                // A (16 n x 8 rows): A[n][r] = Ys[r][n] ; B (8 rows x 8 k): B[r][k] = Xs[r][k]
                uint32_t ah[WM][4], al[WM][4];
#pragma unroll
                for (int a = 0; a < WM; ++a) {
                    const float* ya = Ys + (ks * 8 + t) * WG_LD + wn0 + a * 16 + g;
                    ah[a][0] = __float_as_uint(ya[0]);
                    ah[a][1] = __float_as_uint(ya[8]);
                    ah[a][2] = __float_as_uint(ya[4 * WG_LD]);
                    ah[a][3] = __float_as_uint(ya[4 * WG_LD + 8]);
                    if (PREC3) {
                        const float* yl2 = ya + 32 * WG_LD;
                        al[a][0] = __float_as_uint(yl2[0]);
                        al[a][1] = __float_as_uint(yl2[8]);
                        al[a][2] = __float_as_uint(yl2[4 * WG_LD]);
                        al[a][3] = __float_as_uint(yl2[4 * WG_LD + 8]);
                    }
                }
#pragma unroll
                for (int b = 0; b < WN; ++b) {
                    const float* xb = Xs + (ks * 8 + t) * LDX + wk0 + b * 8 + g;
                    uint32_t bh[2] = {__float_as_uint(xb[0]), __float_as_uint(xb[4 * LDX])};
                    uint32_t bl[2] = {0u, 0u};
                    if (PREC3) {
                        bl[0] = __float_as_uint(xb[32 * LDX]);
                        bl[1] = __float_as_uint(xb[32 * LDX + 4 * LDX]);
                    }
#pragma unroll
                    for (int a = 0; a < WM; ++a) {
                        if (PREC3) {
                            mma_tf32(acc[a][b], al[a], bh);
                            mma_tf32(acc[a][b], ah[a], bl);
                        }
                        mma_tf32(acc[a][b], ah[a], bh);
                    }
                }
            }
        }
    }
    // d0=(g,2t) d1=(g,2t+1) d2=(g+8,2t) d3=(g+8,2t+1): row = n, column = k
#pragma unroll
    for (int a = 0; a < WM; ++a)
#pragma unroll
        for (int b = 0; b < WN; ++b) {
            const int n = n0 + wn0 + a * 16 + g, k = k0 + wk0 + b * 8 + 2 * t;
            if (k < K) {
                if (n < N) atomicAdd(dW + (long long)n * ldw + k, acc[a][b][0]);
                if (n + 8 < N) atomicAdd(dW + (long long)(n + 8) * ldw + k, acc[a][b][2]);
            }
            if (k + 1 < K) {
                if (n < N) atomicAdd(dW + (long long)n * ldw + k + 1, acc[a][b][1]);
                if (n + 8 < N) atomicAdd(dW + (long long)(n + 8) * ldw + k + 1, acc[a][b][3]);
            }
        }
}

template <bool PREC3, int KT, class XL, class YL>
inline cudaError_t launch_wgrad_t(const XL& xl, const YL& yl, float* dW, int ldw, int M, int N, int K, cudaStream_t st) {
    auto kern = wgrad_kernel<PREC3, KT, XL, YL>;
    const int smem = ((PREC3 ? 2 : 1) * 32 * (WG_LD + 64 * KT + 8) + XL::kExtra + YL::kExtra) * 4;
    static SmemCfg cfg;
    if (smem > 48 * 1024)
        if (cudaError_t e = ensure_smem(kern, smem, cfg); e != cudaSuccess) return e;
    const int tiles = ((N + 63) / 64) * ((K + 64 * KT - 1) / (64 * KT));
    const int ngroups = (M + 127) / 128;
    int splits = (sm_count() * 4 + tiles - 1) / tiles;  // ~4 CTAs per SM in total
    if (splits > ngroups) splits = ngroups;
    if (splits < 1) splits = 1;
    const int gpc = (ngroups + splits - 1) / splits;
    splits = (ngroups + gpc - 1) / gpc;
    kern<<<dim3(tiles, splits), 256, smem, st>>>(xl, yl, dW, ldw, M, N, K, gpc);
    return cudaGetLastError();
}

inline bool wgrad_3x() {  // RTFS_WGRAD_3X=1: error-compensated 3xTF32 in every weight-gradient reduction (A/B of the TF32 default)
    static const bool v = [] {
        const char* e = getenv("RTFS_WGRAD_3X");
        return e != nullptr && e[0] != '\0' && e[0] != '0';
    }();
    return v;
}

// dW[N][K] (row stride ldw) += sum_{row < M} Y[row][n] * X[row][k].  EXACT: 3xTF32 regardless of the switch.
template <bool EXACT, class XL, class YL>
inline cudaError_t launch_wgrad(const XL& xl, const YL& yl, float* dW, int ldw, int M, int N, int K, cudaStream_t st) {
    if (EXACT || wgrad_3x()) {
        if (K >= 256) return launch_wgrad_t<true, 4>(xl, yl, dW, ldw, M, N, K, st);
        return launch_wgrad_t<true, 1>(xl, yl, dW, ldw, M, N, K, st);
    }
    if (K >= 256) return launch_wgrad_t<false, 4>(xl, yl, dW, ldw, M, N, K, st);
    return launch_wgrad_t<false, 1>(xl, yl, dW, ldw, M, N, K, st);
}

// ------------------------------------------------------------------------------------------- epilogues of launch_gemm
// C = acc + addend (same shape), e.g. dh_{l-1} = dU_l * W_l + (1 - r) * dh_l   (identity highway of SRU layers 1-3)
struct AddEpi {
    float* C;
    long long ldc;
    const float* addend;  // may be null
    DEVINL void init(int, int) {}
    DEVINL void store(int row, int col, float v0, float v1) {
        const long long o = (long long)row * ldc + col;
        if (addend) {
            const float2 a = ldg2(addend + o);
            v0 += a.x;
            v1 += a.y;
        }
        *reinterpret_cast<float2*>(C + o) = make_float2(v0, v1);
    }
    DEVINL void finish(float*) {}
};

// mask head backward (mask_generator.py:45-60): acc = gradient w.r.t. PReLU(A); dA = acc * (A >= 0 ? 1 : a);
// dslope += sum acc * A * [A < 0]
struct PreluBwdEpi {
    float* dA;           // [M][256]
    const float* A;      // [M][256] the PReLU input (refined features)
    const float* slope;
    float* dslope;
    float a_, ds_;
    DEVINL void init(int, int) {
        a_ = __ldg(slope);
        ds_ = 0.f;
    }
    DEVINL void store(int row, int col, float v0, float v1) {
        const long long o = (long long)row * 256 + col;
        const float2 x = ldg2(A + o);
        if (x.x < 0.f) { ds_ += v0 * x.x; v0 *= a_; }
        if (x.y < 0.f) { ds_ += v1 * x.y; v1 *= a_; }
        *reinterpret_cast<float2*>(dA + o) = make_float2(v0, v1);
    }
    DEVINL void finish(float* scratch) {
        float s = warp_sum(ds_);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += scratch[w];
            if (tot != 0.f) atomicAdd(dslope, tot);
        }
    }
};

// gateway backward (tdanet.py:34-41,107,131): acc = W_p^T dp_pre (projection path); the gateway output r also feeds the
// residual connection, so dr = acc + dOut.  With pre = wg*x + bg:  dpre = dr * (pre >= 0 ? 1 : a),
//   dx = dpre * wg ;  dwg[c] += sum dpre * x ;  dbg[c] += sum dpre ;  dslope += sum dr * pre * [pre < 0].
// Launched with BN = 64: a thread owns 8 fixed columns (4 pairs), whose sums stay in registers across its rows.
struct GateBwdEpi {
    float* dx;          // [M][256]
    const float* dout;  // [M][256] gradient w.r.t. the block output
    const float* x;     // [M][256] block input
    const float* wg;
    const float* bg;
    const float* slope;
    float* dwg;
    float* dbg;
    float* dslope;
    float* dacc;    // optional: running gradient w.r.t. a1 (every block input is out_prev + a1): dacc (+)= dx
    int dacc_init;  // 1: dacc = dx (first contribution), 0: dacc += dx
    float a_, ds_;
    float sw_[4][2], sb_[4][2];
    int colbase_;
    DEVINL void init(int, int) {
        a_ = __ldg(slope);
        ds_ = 0.f;
        colbase_ = -1;
#pragma unroll
        for (int i = 0; i < 4; ++i) sw_[i][0] = sw_[i][1] = sb_[i][0] = sb_[i][1] = 0.f;
    }
    DEVINL void store(int row, int col, float v0, float v1) {
        const long long o = (long long)row * 256 + col;
        const float2 xx = ldg2(x + o), dd = ldg2(dout + o), w = ldg2(wg + col), b = ldg2(bg + col);
        const float p0 = fmaf(w.x, xx.x, b.x), p1 = fmaf(w.y, xx.y, b.y);
        float d0 = v0 + dd.x, d1 = v1 + dd.y;
        if (p0 < 0.f) { ds_ += d0 * p0; d0 *= a_; }
        if (p1 < 0.f) { ds_ += d1 * p1; d1 *= a_; }
        const float2 gx = make_float2(d0 * w.x, d1 * w.y);
        *reinterpret_cast<float2*>(dx + o) = gx;
        if (dacc != nullptr) {
            float2 acc = gx;
            if (!dacc_init) {
                const float2 prev = *reinterpret_cast<const float2*>(dacc + o);
                acc.x += prev.x;
                acc.y += prev.y;
            }
            *reinterpret_cast<float2*>(dacc + o) = acc;
        }
        if (colbase_ < 0) colbase_ = col & ~31;  // first store of this thread: columns are colbase_ + ni*8 + 2t
        const int ni = (col >> 3) & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i == ni) {
                sw_[i][0] += d0 * xx.x;
                sw_[i][1] += d1 * xx.y;
                sb_[i][0] += d0;
                sb_[i][1] += d1;
            }
    }
    DEVINL void finish(float* scratch) {
        // lanes with the same t (lane & 3) hold the same columns: reduce over g (lane >> 2), then over the warps via atomics
        const int t = threadIdx.x & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float a = sw_[i][j], b = sb_[i][j];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                }
                sw_[i][j] = a;
                sb_[i][j] = b;
            }
        // colbase_ may be unset in threads whose rows were all past M: take it from any lane of the warp that has it
        int cb = colbase_;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) cb = max(cb, __shfl_xor_sync(0xffffffffu, cb, o));
        if ((threadIdx.x & 31) < 4 && cb >= 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int col = cb + i * 8 + 2 * t + j;
                    if (sw_[i][j] != 0.f) atomicAdd(dwg + col, sw_[i][j]);
                    if (sb_[i][j] != 0.f) atomicAdd(dbg + col, sb_[i][j]);
                }
        }
        float s = warp_sum(ds_);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += scratch[w];
            if (tot != 0.f) atomicAdd(dslope, tot);
        }
    }
};

// ------------------------------------------------------------------------------------------------- batched small GEMM
// C[b][m][n] = alpha * sum_k A[b](m,k) * B[b](k,n) with element strides (am, ak), (bk, bn), row-major C (ldc);
// fp32 FMAs, 64 x 64 tiles, 16-wide k chunks, 256 threads x (4 x 4).  Used for the attention backward per (b, head):
// M, N, K are 125..1024.
struct BgemmArgs {
    const float* A;
    const float* B;
    float* C;
    long long sa, sb, sc;  // batch strides (elements)
    long long am, ak, bk, bn;
    int ldc;
    int M, N, K;
    float alpha;
};

__global__ void __launch_bounds__(256) bgemm_kernel(BgemmArgs a) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const float* A = a.A + (long long)blockIdx.z * a.sa;
    const float* B = a.B + (long long)blockIdx.z * a.sb;
    float* C = a.C + (long long)blockIdx.z * a.sc;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // loads: 64 x 16 elements per operand and chunk = 4 per thread; the thread->element map follows the unit stride
    const bool a_k_fast = a.ak == 1, b_n_fast = a.bn == 1;
    for (int kc = 0; kc < a.K; kc += 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + 256 * i;
            int m, k;
            if (a_k_fast) { k = e & 15; m = e >> 4; } else { m = e & 63; k = e >> 6; }
            const int gm = m0 + m, gk = kc + k;
            As[k][m] = (gm < a.M && gk < a.K) ? __ldg(A + gm * a.am + gk * a.ak) : 0.f;
            int n, k2;
            if (b_n_fast) { n = e & 63; k2 = e >> 6; } else { k2 = e & 15; n = e >> 4; }
            const int gn = n0 + n, gk2 = kc + k2;
            Bs[k2][n] = (gn < a.N && gk2 < a.K) ? __ldg(B + gk2 * a.bk + gn * a.bn) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < a.N) C[(long long)m * a.ldc + n] = a.alpha * acc[i][j];
        }
    }
}

inline cudaError_t launch_bgemm(const BgemmArgs& a, int batch, cudaStream_t st) {
    bgemm_kernel<<<dim3((a.N + 63) / 64, (a.M + 63) / 64, batch), 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace rtfs
