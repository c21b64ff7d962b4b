// Rolling-row depthwise 4x4 convolutions of the RTFS block on channels-last (B,T,F,64) tensors
// (reference: ConvNormAct with groups=C, layers/conv_layers.py:65-129; the down-samplers tdanet.py:61-76
// and the three TF-AR units layers/fusion.py:25-52).  Second generation of dwconv.cuh:
//
//   * one CTA owns (utterance b, a 16-channel group, a segment of output rows) and marches down T;
//   * every input row is fetched from HBM exactly once per CTA with 16-byte cp.async into a ring of
//     DR_NR shared-memory row slots (DR_NR-1 rows = 40-80 KB in flight per CTA, no registers tied up),
//     the ring slot of row r is recycled for row r+DR_NR one barrier later;
//   * thread = (channel pair, strip of 4 output columns): a 4-row x 7-column register window rolls down T,
//     so each shared-memory row is read once, when it becomes the newest row of the window;
//   * the input is produced on the fly by a functor (gLN-apply / PReLU / TF-AR combine with nearest
//     up-sampling), so normalised tensors are never materialised; zero padding is applied after it;
//   * NW convolutions can share one input (the four "global" convs of the TF-AR units); DUAL adds the
//     stride-2 down-sampler + the 3x3/2 adaptive average pool on the same window (tdanet.py:69-76,117);
//   * epilogue: per-sample (sum, sumsq) of every output for the following gLN, one fp64 atomic pair per CTA.
//
//   out[t][f][c] = bias[c] + sum_{i,j<4} w[c][i][j] * X[S*t-1+i][S*f-1+j][c]   (zero outside), S = 1 (or 2 for DUAL's second output)
#pragma once
#include <cstdlib>
#include "common.cuh"

namespace rtfs {

constexpr int DR_CG = 16;  // channels per CTA (one 64-byte piece of every position)
constexpr int DR_NR = 6;   // ring slots

// position -> slot index; swaps neighbours in every other group of 4 so that the two strips a
// half-warp covers hit different bank halves (positions are 64 bytes = 16 banks wide)
DEVINL int dr_phys(int f) { return f ^ ((f >> 2) & 1); }

DEVINL void cp_async16_always(void* smem, const void* gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
DEVINL void cp_async_wait_dyn(int n) {  // n in [0, DR_NR-2]
    switch (n) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        default: cp_async_wait<4>(); break;
    }
}

// copy the 16-channel piece of one (B,T,F,64) row into a slot: position f -> slot + dr_phys(f)*16
DEVINL void dr_issue_row(float* slot, const float* row_base /* + cg*16 applied */, int Fi, int tid, int nthreads) {
    for (int i = tid; i < Fi * 4; i += nthreads) {
        const int pos = i >> 2, c = i & 3;
        cp_async16_always(slot + dr_phys(pos) * 16 + c * 4, row_base + (long long)pos * 64 + c * 4);
    }
}

// ---------------------------------------------------------------- input functors
// contract: slot_floats() ; init(b, cg, pair, f0, tid, nthreads) ; issue(r, slot_u32) for a valid input row r ;
//           kPost / post(slot): optional in-place pass over the freshly landed row (one row ahead of its use) ;
//           fetch(slot, q) -> transformed channel pair at window column q (column clamped into range; the kernel
//           applies the zero padding) .
// The kernel is issue-bound (ncu: 55-70 % issue-active, 337 instructions per warp and row of which 67 FFMA2), so the
// functors keep every per-row address as base + immediate: per-thread cp.async piece tables, four window base offsets.

DEVINL void cp_async16_u32(uint32_t s, const void* gmem) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem)); }

// this thread's cp.async pieces of one row copy (F positions x 16 channels = F*4 pieces of 16 bytes; F*4 <= 2*nthreads)
struct DrPieces {
    int n_;
    uint32_t so_[2];  // byte offset inside the row area of a slot
    int go_[2];       // float offset inside the global row
    DEVINL void init(int F, int tid, int nthreads) {
        n_ = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int i = tid + j * nthreads;
            const int pos = i >> 2, c = i & 3;
            so_[j] = (uint32_t)(dr_phys(pos) * 16 + c * 4) * 4u;
            go_[j] = pos * 64 + c * 4;
            if (i < F * 4) n_ = j + 1;
        }
    }
    DEVINL void issue(uint32_t area_u32, const float* grow) const {
        if (n_ > 0) cp_async16_u32(area_u32 + so_[0], grow + go_[0]);
        if (n_ > 1) cp_async16_u32(area_u32 + so_[1], grow + go_[1]);
    }
};

// float offsets of the 7 window columns f0-1 .. f0+5 of a strip (f0 = 4*strip) under the dr_phys swizzle, as four bases
// + compile-time column offsets: columns of the strip's own group of 4 and of the two neighbouring groups (opposite parity)
struct DrWin {
    int e_, o_, ne_, no_;  // even / odd columns of the own group, of the neighbour groups
    DEVINL void init(int f0, int pair, int F) {
        const int p = (f0 >> 2) & 1, base = f0 * 16 + 2 * pair;
        e_ = base + 16 * p;
        o_ = base - 16 * p;
        ne_ = base + 16 * (1 - p);
        no_ = base - 16 * (1 - p);
        (void)F;
    }
    // q = 0: f0-1 (odd column 3 of the previous group); q = 1..4: own group; q = 5, 6: columns 0, 1 of the next group
    DEVINL int at(int q) const { return q == 0 ? no_ - 16 : q <= 4 ? ((q - 1) & 1 ? o_ : e_) + (q - 1) * 16 : (q == 5 ? ne_ : no_) + (q - 1) * 16; }
};

struct XrPlain {
    const float* x;  // [B][Ti][Fi][64]
    int Ti, Fi;
    static constexpr bool kPost = false;
    const float* base_;
    DrPieces pc_;
    DrWin wn_;
    __host__ __device__ __forceinline__ int slot_floats() const { return (Fi + 10) * 16; }  // 4 pad positions left, 6 right: the swizzled window columns of the edge strips
    DEVINL void init(int b, int cg, int pair, int f0, int tid, int nthreads) {
        base_ = x + (long long)b * Ti * Fi * 64 + cg * 16;
        pc_.init(Fi, tid, nthreads);
        wn_.init(f0, pair, Fi);
    }
    DEVINL void issue(int r, uint32_t slot_u32) const { pc_.issue(slot_u32 + 4 * 16 * 4, base_ + (long long)r * Fi * 64); }
    DEVINL void post(float*) const {}
    DEVINL float2 fetch(const float* slot, int q) const { return *reinterpret_cast<const float2*>(slot + 4 * 16 + wn_.at(q)); }
};

// X = act(gLN(x)) ; ACT 0 none / 2 PReLU(slope)
template <int ACT>
struct XrGln {
    const float* x;
    int Ti, Fi;
    GlnRef gln;
    const float* slope;
    static constexpr bool kPost = false;
    const float* base_;
    DrPieces pc_;
    DrWin wn_;
    float2 sc_, sh_;
    float a_;
    __host__ __device__ __forceinline__ int slot_floats() const { return (Fi + 10) * 16; }  // 4 pad positions left, 6 right: the swizzled window columns of the edge strips
    DEVINL void init(int b, int cg, int pair, int f0, int tid, int nthreads) {
        base_ = x + (long long)b * Ti * Fi * 64 + cg * 16;
        pc_.init(Fi, tid, nthreads);
        wn_.init(f0, pair, Fi);
        const int c = cg * 16 + 2 * pair;
        float mean, rstd;
        gln_mean_rstd(gln.sums, b, gln.inv_n, mean, rstd);
        const float2 g = ldg2(gln.gamma + c), be = ldg2(gln.beta + c);
        sc_ = make_float2(rstd * g.x, rstd * g.y);
        sh_ = make_float2(be.x - mean * sc_.x, be.y - mean * sc_.y);
        a_ = (ACT == 2) ? __ldg(slope) : 0.f;
    }
    DEVINL void issue(int r, uint32_t slot_u32) const { pc_.issue(slot_u32 + 4 * 16 * 4, base_ + (long long)r * Fi * 64); }
    DEVINL void post(float*) const {}
    DEVINL float2 fetch(const float* slot, int q) const {
        const float2 v = *reinterpret_cast<const float2*>(slot + 4 * 16 + wn_.at(q));
        float2 y = __ffma2_rn(v, sc_, sh_);
        if (ACT == 2) {
            y.x = prelu(y.x, a_);
            y.y = prelu(y.y, a_);
        }
        return y;
    }
};

// X = TF-AR output (layers/fusion.py:54-69) formed on the fly:
//   gLN_l(l)[t][f] * sigmoid(gLN_g(g))[near(t)][near(f)] + gLN_e(e)[near(t)][near(f)]
// l at (Ti,Fi); g, e at (Tg,Fg) (nearest up-sampling; identity when the sizes are equal).  Every row slot carries its
// own copy of the g / e row it up-samples from; sigmoid(gLN_g(.)) and gLN_e(.) are applied to that copy IN PLACE once
// per row by the thread that copied the piece (post), instead of once per window column and output position in fetch
// (7x fewer sigmoids at half resolution), which also frees the registers the packed kernel needs.
struct XrTfarP {
    const float* l;
    const float* g;
    const float* e;
    int Ti, Fi, Tg, Fg;
    GlnRef nl, ng, ne;
    static constexpr bool kPost = true;
    const float *bl_, *bg_, *be_;
    DrPieces pl_, pg_;
    DrWin wn_;
    int lfl_, gfl_;
    int gofs_[7];  // float offset (inside the g area) of the up-sampled column of each window column
    float2 scl_, shl_;
    const float4* ptab_;  // shared: [4 channel quads][scg, shg, sce, she] of this CTA's 16 channels (kept out of registers)
    __host__ __device__ __forceinline__ int slot_floats() const { return (Fi + 10) * 16 + 2 * (Fg + 2) * 16; }
    DEVINL void mk(const GlnRef& r, int b, int c, float& sc, float& sh) const {
        float mean, rstd;
        gln_mean_rstd(r.sums, b, r.inv_n, mean, rstd);
        sc = rstd * __ldg(r.gamma + c);
        sh = __ldg(r.beta + c) - mean * sc;
    }
    DEVINL void init(int b, int cg, int pair, int f0, int tid, int nthreads) {
        bl_ = l + (long long)b * Ti * Fi * 64 + cg * 16;
        bg_ = g + (long long)b * Tg * Fg * 64 + cg * 16;
        be_ = e + (long long)b * Tg * Fg * 64 + cg * 16;
        lfl_ = (Fi + 10) * 16;
        gfl_ = (Fg + 2) * 16;
        pl_.init(Fi, tid, nthreads);
        pg_.init(Fg, tid, nthreads);
        wn_.init(f0, pair, Fi);
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            int f = f0 - 1 + q;
            f = f < 0 ? 0 : (f > Fi - 1 ? Fi - 1 : f);
            gofs_[q] = dr_phys(nearest_src32(f, Fg, Fi)) * 16 + 2 * pair;
        }
        const int c = cg * 16 + 2 * pair;
        mk(nl, b, c, scl_.x, shl_.x);
        mk(nl, b, c + 1, scl_.y, shl_.y);
        __shared__ float4 ptab[4][4];
        if (tid < 16) {
            float sc, sh;
            float* tb = reinterpret_cast<float*>(&ptab[tid >> 2][0]) + (tid & 3);
            mk(ng, b, cg * 16 + tid, sc, sh);
            tb[0] = sc;
            tb[4] = sh;
            mk(ne, b, cg * 16 + tid, sc, sh);
            tb[8] = sc;
            tb[12] = sh;
        }
        ptab_ = &ptab[tid & 3][0];  // channels of this thread's copied pieces (visible after the kernel's first barrier)
    }
    DEVINL void issue(int r, uint32_t slot_u32) const {
        pl_.issue(slot_u32 + 4 * 16 * 4, bl_ + (long long)r * Fi * 64);
        const int rg = nearest_src32(r, Tg, Ti);
        pg_.issue(slot_u32 + lfl_ * 4, bg_ + (long long)rg * Fg * 64);
        pg_.issue(slot_u32 + (lfl_ + gfl_) * 4, be_ + (long long)rg * Fg * 64);
    }
    // the pieces this thread copied itself (its own cp.async group has completed when post runs)
    DEVINL void post(float* slot) const {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (j < pg_.n_) {
                float4* pgp = reinterpret_cast<float4*>(slot + lfl_ + (pg_.so_[j] >> 2));
                float4* pep = reinterpret_cast<float4*>(slot + lfl_ + gfl_ + (pg_.so_[j] >> 2));
                float4 vg = *pgp, ve = *pep;
                const float4 scg = ptab_[0], shg = ptab_[1], sce = ptab_[2], she = ptab_[3];
                vg.x = sigmoidf_fast(fmaf(vg.x, scg.x, shg.x));
                vg.y = sigmoidf_fast(fmaf(vg.y, scg.y, shg.y));
                vg.z = sigmoidf_fast(fmaf(vg.z, scg.z, shg.z));
                vg.w = sigmoidf_fast(fmaf(vg.w, scg.w, shg.w));
                ve.x = fmaf(ve.x, sce.x, she.x);
                ve.y = fmaf(ve.y, sce.y, she.y);
                ve.z = fmaf(ve.z, sce.z, she.z);
                ve.w = fmaf(ve.w, sce.w, she.w);
                *pgp = vg;
                *pep = ve;
            }
        }
    }
    DEVINL float2 fetch(const float* slot, int q) const {
        const float2 vl = *reinterpret_cast<const float2*>(slot + 4 * 16 + wn_.at(q));
        const float2 sg = *reinterpret_cast<const float2*>(slot + lfl_ + gofs_[q]);
        const float2 ve = *reinterpret_cast<const float2*>(slot + lfl_ + gfl_ + gofs_[q]);
        return __ffma2_rn(__ffma2_rn(vl, scl_, shl_), sg, ve);
    }
};

template <int NW>
struct DrArgs {
    int Ti, Fi;             // input = stride-1 output size
    int rows_per_seg;       // output rows per CTA
    const float* w[NW];     // [16][64] tap-major
    const float* bias[NW];  // [64] or null
    float* out[NW];         // [B][Ti][Fi][64]
    double* sums[NW];       // [B][2] or null
    // DUAL: stride-2 conv (pad 1) + adaptive average pool onto (To2, Fo2)
    const float* w2;
    const float* bias2;
    float* out2;
    double* sums2;
    float* pool;
    int To2, Fo2;
};

// NT threads = 8 channel pairs x NT/8 strips of 4 columns  (NT/8 >= ceil(Fi/4))
template <class XF, int NW, bool DUAL, int NT, int NR = DR_NR>
__global__ void __launch_bounds__(NT, (NT > 256 ? 2 : 4)) dwroll_kernel(XF xf, DrArgs<NW> a) {
    constexpr int NCONV = NW + (DUAL ? 1 : 0);
    constexpr int AHEAD = XF::kPost ? NR - 3 : NR - 2;  // cp.async groups left in flight when a row is consumed
    static_assert(AHEAD >= 1, "ring too short");
    extern __shared__ __align__(16) float dr_smem[];
    __shared__ float wsm[NCONV][16][DR_CG];  // filter taps of this channel group
    __shared__ float bsm[NCONV][DR_CG];
    __shared__ float scratch[2 * (NT / 32)];

    const int tid = threadIdx.x, pair = tid & 7, strip = tid >> 3;
    const int cg = blockIdx.x & 3, seg = blockIdx.x >> 2, b = blockIdx.y;
    const int Ti = a.Ti, Fi = a.Fi;
    const int f0 = 4 * strip;
    const bool active = f0 < Fi;
    const int t0 = seg * a.rows_per_seg;
    const int t1 = min(t0 + a.rows_per_seg, Ti);
    const int slot_fl = xf.slot_floats();

    for (int i = tid; i < NCONV * 16 * DR_CG; i += NT) {
        const int w = i / (16 * DR_CG), rem = i - w * 16 * DR_CG, tap = rem / DR_CG, c = rem - tap * DR_CG;
        const float* wp = (DUAL && w == NW) ? a.w2 : a.w[w < NW ? w : 0];
        wsm[w][tap][c] = __ldg(wp + tap * 64 + cg * DR_CG + c);
    }
    for (int i = tid; i < NCONV * DR_CG; i += NT) {
        const int w = i / DR_CG, c = i - w * DR_CG;
        const float* bp = (DUAL && w == NW) ? a.bias2 : a.bias[w < NW ? w : 0];
        bsm[w][c] = bp ? __ldg(bp + cg * DR_CG + c) : 0.f;
    }
    xf.init(b, cg, pair, active ? f0 : 0, tid, NT);
    // column validity of the 7 window columns (zero padding is applied AFTER the input transform)
    bool cv[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) cv[q] = active && (f0 - 1 + q >= 0) && (f0 - 1 + q < Fi);

    // input rows r = t0-1 .. t1+1 ; row r lives in slot (r - (t0-1)) % NR: running offsets instead of a modulo per row
    const int r_first = t0 - 1, r_last = t1 + 1;
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(dr_smem);
    const int ring_fl = NR * slot_fl;
    int r_issue = r_first, off_issue = 0;
    auto issue_next = [&]() {
        if (r_issue >= 0 && r_issue < Ti && r_issue <= r_last) xf.issue(r_issue, smem0 + 4u * off_issue);
        cp_async_commit();
        ++r_issue;
        off_issue += slot_fl;
        if (off_issue == ring_fl) off_issue = 0;
    };
#pragma unroll
    for (int i = 0; i < NR - 1; ++i) issue_next();
    if (XF::kPost) {  // the first row is post-processed before the loop, every later row one step ahead of its use
        cp_async_wait<NR - 2>();
        __syncthreads();  // the functor's shared tables are complete
        if (r_first >= 0) xf.post(dr_smem);
    }

    // win[slot][q]: 4 input rows x 7 columns (f0-1 .. f0+5).  The row loop is unrolled by 4 so that the slot a new
    // row lands in (step & 3) and the logical window order are compile-time: no register moves when the window rolls.
    float2 win[4][7];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 7; ++q) win[i][q] = make_float2(0.f, 0.f);
    float2 st_s[NCONV], st_q[NCONV];  // per channel of the pair; packed f32x2 math (FFMA2) throughout
#pragma unroll
    for (int w = 0; w < NCONV; ++w) st_s[w] = st_q[w] = make_float2(0.f, 0.f);

    const int wT = 2 + (Ti & 1), wF = 2 + (Fi & 1);
    const float pool_scale = 1.f / (float)(wT * wF);
    const int c0 = cg * DR_CG + 2 * pair;
    // output row pointers advance by one row per step
    float* orow[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) orow[w] = a.out[w] + (((long long)b * Ti + t0) * Fi + f0) * 64 + c0;
    const int ostride = Fi * 64;
    bool ov[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) ov[o] = f0 + o < Fi;

    int off_cur = 0;
    for (int rb = r_first; rb <= r_last; rb += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = rb + k;
            if (r > r_last) break;       // uniform over the CTA
            cp_async_wait<AHEAD>();      // this thread's pieces of row r (and of row r+1 when kPost) have landed
            __syncthreads();             // everyone's have; everyone is done reading row r-1's slot
            issue_next();                // -> into the slot of row r-1
            const float* slot = dr_smem + off_cur;
            off_cur += slot_fl;
            if (off_cur == ring_fl) off_cur = 0;
            if (XF::kPost && r + 1 >= 0 && r + 1 < Ti && r + 1 <= r_last) xf.post(dr_smem + off_cur);
            if (r >= 0 && r < Ti) {      // uniform
#pragma unroll
                for (int q = 0; q < 7; ++q) {
                    const float2 v = xf.fetch(slot, q);
                    win[k][q] = make_float2(cv[q] ? v.x : 0.f, cv[q] ? v.y : 0.f);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 7; ++q) win[k][q] = make_float2(0.f, 0.f);
            }
            const int t = r - 2;  // output row whose window (rows t-1..t+2) is now complete
            if (t < t0) continue;
            // logical window row i (input row t-1+i) = win[(k + 1 + i) & 3]
            // ---- stride-1 outputs
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                float2 acc[4];
                const float2 bs = *reinterpret_cast<const float2*>(&bsm[w][2 * pair]);
#pragma unroll
                for (int o = 0; o < 4; ++o) acc[o] = bs;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 wv = *reinterpret_cast<const float2*>(&wsm[w][i * 4 + j][2 * pair]);
#pragma unroll
                        for (int o = 0; o < 4; ++o) acc[o] = __ffma2_rn(wv, win[(k + 1 + i) & 3][o + j], acc[o]);
                    }
#pragma unroll
                for (int o = 0; o < 4; ++o)
                    if (ov[o]) {
                        *reinterpret_cast<float2*>(orow[w] + o * 64) = acc[o];
                        st_s[w] = __fadd2_rn(st_s[w], acc[o]);
                        st_q[w] = __ffma2_rn(acc[o], acc[o], st_q[w]);
                    }
                orow[w] += ostride;
            }
            // ---- stride-2 output row t/2 and the adaptive average pool
            if (DUAL && (t & 1) == 0 && (t >> 1) < a.To2) {
                const int to = t >> 1;
                const float2 bs = *reinterpret_cast<const float2*>(&bsm[NW][2 * pair]);
                float2 acc[2] = {bs, bs};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 wv = *reinterpret_cast<const float2*>(&wsm[NW][i * 4 + j][2 * pair]);
#pragma unroll
                        for (int o = 0; o < 2; ++o) acc[o] = __ffma2_rn(wv, win[(k + 1 + i) & 3][2 * o + j], acc[o]);
                    }
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const int fo = (f0 >> 1) + o;
                    if (active && fo < a.Fo2) {
                        const long long off = (((long long)b * a.To2 + to) * a.Fo2 + fo) * 64 + c0;
                        *reinterpret_cast<float2*>(a.out2 + off) = acc[o];
                        st_s[NW] = __fadd2_rn(st_s[NW], acc[o]);
                        st_q[NW] = __ffma2_rn(acc[o], acc[o], st_q[NW]);
                        float2 p = make_float2(0.f, 0.f);
#pragma unroll
                        for (int i = 1; i < 4; ++i)
#pragma unroll
                            for (int j = 1; j < 4; ++j)
                                if (i <= wT && j <= wF) p = __fadd2_rn(p, win[(k + 1 + i) & 3][2 * o + j]);
                        *reinterpret_cast<float2*>(a.pool + off) = make_float2(p.x * pool_scale, p.y * pool_scale);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int w = 0; w < NCONV; ++w) {
        double* dst = (DUAL && w == NW) ? a.sums2 : a.sums[w < NW ? w : 0];
        block_stats_atomic(st_s[w].x + st_s[w].y, st_q[w].x + st_q[w].y, dst ? dst + 2 * b : nullptr, scratch);
    }
}

// ---------------------------------------------------------------- first generation (A/B: RTFS_SCALAR_TFAR=1)
// X = TF-AR output (layers/fusion.py:54-69) formed on the fly:
//   gLN_l(l)[t][f] * sigmoid(gLN_g(g))[near(t)][near(f)] + gLN_e(e)[near(t)][near(f)]
// l at (Ti,Fi); g, e at (Tg,Fg) (nearest up-sampling; identity when the sizes are equal).
struct XrTfar {  // first-generation API (dwroll_scalar_kernel only)
    const float* l;
    const float* g;
    const float* e;
    int Ti, Fi, Tg, Fg;
    GlnRef nl, ng, ne;
    const float *bl_, *bg_, *be_;
    int f0_, pair_, lfl_, gfl_;
    int gofs_[7];  // slot offset of the up-sampled column of each window column
    float2 scl_, shl_, scg_, shg_, sce_, she_;
    __host__ __device__ __forceinline__ int slot_floats() const { return (Fi + 2) * 16 + 2 * (Fg + 2) * 16; }
    DEVINL void mk(const GlnRef& r, int b, int c, float2& sc, float2& sh) {
        float mean, rstd;
        gln_mean_rstd(r.sums, b, r.inv_n, mean, rstd);
        const float2 gm = ldg2(r.gamma + c), be = ldg2(r.beta + c);
        sc = make_float2(rstd * gm.x, rstd * gm.y);
        sh = make_float2(be.x - mean * sc.x, be.y - mean * sc.y);
    }
    DEVINL void init(int b, int cg, int pair, int f0) {
        bl_ = l + (long long)b * Ti * Fi * 64 + cg * 16;
        bg_ = g + (long long)b * Tg * Fg * 64 + cg * 16;
        be_ = e + (long long)b * Tg * Fg * 64 + cg * 16;
        f0_ = f0;
        pair_ = pair;
        lfl_ = (Fi + 2) * 16;
        gfl_ = (Fg + 2) * 16;
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            int f = f0 - 1 + q;
            f = f < 0 ? 0 : (f > Fi - 1 ? Fi - 1 : f);
            gofs_[q] = lfl_ + dr_phys(nearest_src32(f, Fg, Fi)) * 16 + 2 * pair;
        }
        const int c = cg * 16 + 2 * pair;
        mk(nl, b, c, scl_, shl_);
        mk(ng, b, c, scg_, shg_);
        mk(ne, b, c, sce_, she_);
    }
    DEVINL void issue(int r, float* slot, int tid, int nthreads) const {
        dr_issue_row(slot, bl_ + (long long)r * Fi * 64, Fi, tid, nthreads);
        const int rg = nearest_src32(r, Tg, Ti);
        dr_issue_row(slot + lfl_, bg_ + (long long)rg * Fg * 64, Fg, tid, nthreads);
        dr_issue_row(slot + lfl_ + gfl_, be_ + (long long)rg * Fg * 64, Fg, tid, nthreads);
    }
    DEVINL float2 fetch(const float* slot, int q) const {
        const int f = f0_ - 1 + q;
        const float2 vl = *reinterpret_cast<const float2*>(slot + dr_phys(f) * 16 + 2 * pair_);
        const float2 vg = *reinterpret_cast<const float2*>(slot + gofs_[q]);
        const float2 ve = *reinterpret_cast<const float2*>(slot + gofs_[q] + gfl_);
        float2 y;
        y.x = fmaf(vl.x, scl_.x, shl_.x) * sigmoidf_fast(fmaf(vg.x, scg_.x, shg_.x)) + fmaf(ve.x, sce_.x, she_.x);
        y.y = fmaf(vl.y, scl_.y, shl_.y) * sigmoidf_fast(fmaf(vg.y, scg_.y, shg_.y)) + fmaf(ve.y, sce_.y, she_.y);
        return y;
    }
};

// First version of the rolling kernel (scalar FFMA, window rolled with register moves); still the faster one for the
// TF-AR input functor at full resolution, where the packed version spills.
template <class XF, int NW, bool DUAL, int NT>
__global__ void __launch_bounds__(NT, (NT > 256 ? 2 : 4)) dwroll_scalar_kernel(XF xf, DrArgs<NW> a) {
    constexpr int NCONV = NW + (DUAL ? 1 : 0);
    extern __shared__ __align__(16) float dr_smem[];
    __shared__ float wsm[NCONV][16][DR_CG];  // filter taps of this channel group
    __shared__ float bsm[NCONV][DR_CG];
    __shared__ float scratch[2 * (NT / 32)];

    const int tid = threadIdx.x, pair = tid & 7, strip = tid >> 3;
    const int cg = blockIdx.x & 3, seg = blockIdx.x >> 2, b = blockIdx.y;
    const int Ti = a.Ti, Fi = a.Fi;
    const int f0 = 4 * strip;
    const bool active = f0 < Fi;
    const int t0 = seg * a.rows_per_seg;
    const int t1 = min(t0 + a.rows_per_seg, Ti);
    const int slot_fl = xf.slot_floats();

    for (int i = tid; i < NCONV * 16 * DR_CG; i += NT) {
        const int w = i / (16 * DR_CG), rem = i - w * 16 * DR_CG, tap = rem / DR_CG, c = rem - tap * DR_CG;
        const float* wp = (DUAL && w == NW) ? a.w2 : a.w[w < NW ? w : 0];
        wsm[w][tap][c] = __ldg(wp + tap * 64 + cg * DR_CG + c);
    }
    for (int i = tid; i < NCONV * DR_CG; i += NT) {
        const int w = i / DR_CG, c = i - w * DR_CG;
        const float* bp = (DUAL && w == NW) ? a.bias2 : a.bias[w < NW ? w : 0];
        bsm[w][c] = bp ? __ldg(bp + cg * DR_CG + c) : 0.f;
    }
    xf.init(b, cg, pair, active ? f0 : 0);

    // input rows r = t0-1 .. t1+1 ; row r lives in slot (r - (t0-1)) % DR_NR
    const int r_first = t0 - 1, r_last = t1 + 1;
    auto issue = [&](int r) {
        if (r >= 0 && r < Ti && r <= r_last) xf.issue(r, dr_smem + ((r - r_first) % DR_NR) * slot_fl, tid, NT);
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < DR_NR - 1; ++i) issue(r_first + i);

    float2 win[4][7];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 7; ++q) win[i][q] = make_float2(0.f, 0.f);
    float st_s[NCONV], st_q[NCONV];
#pragma unroll
    for (int w = 0; w < NCONV; ++w) st_s[w] = st_q[w] = 0.f;

    const int wT = 2 + (Ti & 1), wF = 2 + (Fi & 1);
    const float pool_scale = 1.f / (float)(wT * wF);
    const int c0 = cg * DR_CG + 2 * pair;

    for (int r = r_first; r <= r_last; ++r) {
        cp_async_wait<DR_NR - 2>();  // this thread's pieces of row r have landed
        __syncthreads();             // everyone's have; everyone is done reading row r-1's slot
        issue(r + DR_NR - 1);        // -> into the slot of row r-1
        // roll the window and bring in row r
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            win[0][q] = win[1][q];
            win[1][q] = win[2][q];
            win[2][q] = win[3][q];
        }
        const bool rvalid = r >= 0 && r < Ti;
        const float* slot = dr_smem + ((r - r_first) % DR_NR) * slot_fl;
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const int f = f0 - 1 + q;
            win[3][q] = (active && rvalid && f >= 0 && f < Fi) ? xf.fetch(slot, q) : make_float2(0.f, 0.f);
        }
        const int t = r - 2;  // output row whose window (rows t-1..t+2) is now complete
        if (t < t0 || !active) continue;
        // ---- stride-1 outputs
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            float2 acc[4];
            const float2 bs = *reinterpret_cast<const float2*>(&bsm[w][2 * pair]);
#pragma unroll
            for (int o = 0; o < 4; ++o) acc[o] = bs;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 wv = *reinterpret_cast<const float2*>(&wsm[w][i * 4 + j][2 * pair]);
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        acc[o].x = fmaf(wv.x, win[i][o + j].x, acc[o].x);
                        acc[o].y = fmaf(wv.y, win[i][o + j].y, acc[o].y);
                    }
                }
            float* orow = a.out[w] + (((long long)b * Ti + t) * Fi + f0) * 64 + c0;
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (f0 + o < Fi) {
                    *reinterpret_cast<float2*>(orow + o * 64) = acc[o];
                    st_s[w] += acc[o].x + acc[o].y;
                    st_q[w] += acc[o].x * acc[o].x + acc[o].y * acc[o].y;
                }
        }
        // ---- stride-2 output row t/2 and the adaptive average pool
        if (DUAL && (t & 1) == 0 && (t >> 1) < a.To2) {
            const int to = t >> 1;
            const float2 bs = *reinterpret_cast<const float2*>(&bsm[NW][2 * pair]);
            float2 acc[2] = {bs, bs};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 wv = *reinterpret_cast<const float2*>(&wsm[NW][i * 4 + j][2 * pair]);
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        acc[o].x = fmaf(wv.x, win[i][2 * o + j].x, acc[o].x);
                        acc[o].y = fmaf(wv.y, win[i][2 * o + j].y, acc[o].y);
                    }
                }
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                const int fo = (f0 >> 1) + o;
                if (fo < a.Fo2) {
                    const long long off = (((long long)b * a.To2 + to) * a.Fo2 + fo) * 64 + c0;
                    *reinterpret_cast<float2*>(a.out2 + off) = acc[o];
                    st_s[NW] += acc[o].x + acc[o].y;
                    st_q[NW] += acc[o].x * acc[o].x + acc[o].y * acc[o].y;
                    float2 p = make_float2(0.f, 0.f);
#pragma unroll
                    for (int i = 1; i < 4; ++i)
#pragma unroll
                        for (int j = 1; j < 4; ++j)
                            if (i <= wT && j <= wF) {
                                p.x += win[i][2 * o + j].x;
                                p.y += win[i][2 * o + j].y;
                            }
                    *reinterpret_cast<float2*>(a.pool + off) = make_float2(p.x * pool_scale, p.y * pool_scale);
                }
            }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int w = 0; w < NCONV; ++w) {
        double* dst = (DUAL && w == NW) ? a.sums2 : a.sums[w < NW ? w : 0];
        block_stats_atomic(st_s[w], st_q[w], dst ? dst + 2 * b : nullptr, scratch);
    }
}

// rows per CTA: as few segments as keep the 3-row halo re-read small while filling 148 SMs evenly
inline int dr_rows_per_seg(int T, int B, int ctas_per_sm) {
    const int slots = sm_count() * ctas_per_sm;
    int best = T;
    double best_eff = 0.0;
    for (int nseg = 1; nseg <= 24 && nseg <= T; ++nseg) {
        const int rows = (T + nseg - 1) / nseg;
        const int nseg_eff = (T + rows - 1) / rows;
        const long long ctas = (long long)B * 4 * nseg_eff;
        const long long waves = (ctas + slots - 1) / slots;
        const double eff = ((double)ctas / (double)(waves * slots)) * ((double)rows / (double)(rows + 3 + DR_NR - 1));
        if (eff > best_eff) {
            best_eff = eff;
            best = rows;
        }
    }
    return best;
}

// NR: ring slots (rows in flight + 1); the compressed-resolution instances use 4 so that 4 CTAs fit an SM
inline int dr_rows_env() {
    static const int rows_env = [] {
        const char* v = getenv("RTFS_DW_ROWS");  // tuning override: output rows per CTA
        return v ? atoi(v) : 0;
    }();
    return rows_env;
}
inline int dr_rows(int Ti, int B) {  // 32: measured best at T = 251 (A/B of 16..126)
    const int e = dr_rows_env();
    return e > 0 ? (e < Ti ? e : Ti) : (Ti >= 96 ? 32 : dr_rows_per_seg(Ti, B, 2));
}

template <class XF, int NW, bool DUAL, int NT, int NR = DR_NR>
inline cudaError_t launch_dwroll(const XF& xf, DrArgs<NW> a, int B, cudaStream_t st) {
    auto kern = dwroll_kernel<XF, NW, DUAL, NT, NR>;
    const int smem = NR * xf.slot_floats() * 4;
    static SmemCfg cfg;  // per instantiation, per device
    if (cudaError_t e = ensure_smem(kern, smem, cfg); e != cudaSuccess) return e;
    a.rows_per_seg = dr_rows(a.Ti, B);
    const int nseg = (a.Ti + a.rows_per_seg - 1) / a.rows_per_seg;
    kern<<<dim3(4 * nseg, B), NT, smem, st>>>(xf, a);
    return cudaGetLastError();
}

template <class XF, int NW, bool DUAL, int NT>
inline cudaError_t launch_dwroll_scalar(const XF& xf, DrArgs<NW> a, int B, cudaStream_t st) {
    auto kern = dwroll_scalar_kernel<XF, NW, DUAL, NT>;
    const int smem = DR_NR * xf.slot_floats() * 4;
    static SmemCfg cfg;  // per instantiation, per device
    if (cudaError_t e = ensure_smem(kern, smem, cfg); e != cudaSuccess) return e;
    a.rows_per_seg = dr_rows(a.Ti, B);
    const int nseg = (a.Ti + a.rows_per_seg - 1) / a.rows_per_seg;
    kern<<<dim3(4 * nseg, B), NT, smem, st>>>(xf, a);
    return cudaGetLastError();
}

}  // namespace rtfs
