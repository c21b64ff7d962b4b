// Training step orchestration (included at the end of api.cu: shares its Ctx / run_* helpers).
// Forward with a tape, backward through every stage of the path; see include/rtfs_b200.h for the contract and
// train_kernels.cuh / train_gemm.cuh / train_rnn.cuh / train_att.cuh / train_misc.cuh for the kernels.
#include "train_att.cuh"
#include "train_gemm.cuh"
#include "train_kernels.cuh"
#include "train_misc.cuh"
#include "train_rnn.cuh"

namespace {

struct TapePlan {
    long long off[RTFS_TP_COUNT];
    long long pass_bytes, total;
};

long long align256(long long bytes) { return ((bytes + 255) / 256) * 256; }

TapePlan make_tape_plan(const Dims& d, int R) {
    const long long B = d.B, A = B * d.P * 256 * 4;
    long long sz[RTFS_TP_COUNT];
    sz[RTFS_TP_SPEC] = B * d.P * 2 * 4;
    sz[RTFS_TP_A0] = sz[RTFS_TP_A1] = sz[RTFS_TP_BLK0] = sz[RTFS_TP_REFINED] = sz[RTFS_TP_M] = sz[RTFS_TP_Z] = A;
    sz[RTFS_TP_X] = A * (R > 1 ? R - 1 : 0);
    sz[RTFS_TP_Q18] = B * d.P * 18 * 4;
    sz[RTFS_TP_VK] = sz[RTFS_TP_ATT] = B * (long long)(d.Tv > 0 ? d.Tv : 1) * 256 * 4;
    sz[RTFS_TP_CAFSUM] = 256 * 2 * 8;
    sz[RTFS_TP_PASS0] = 0;
    TapePlan p;
    long long o = 0;
    for (int i = 0; i < RTFS_TP_COUNT; ++i) {
        p.off[i] = o;
        o += align256(sz[i]);
    }
    p.pass_bytes = make_plan(d, true).total;
    p.total = o + p.pass_bytes * R;
    return p;
}

struct BwdPlan {
    long long off[RTFS_BW_COUNT];
    long long total;
};

BwdPlan make_bwd_plan(const Dims& d) {
    const long long B = d.B, A = B * d.P * 256, H = B * d.P * 64, G = B * d.Pc * 64;
    const long long hp_f = (long long)d.Tc * (d.Fc + 7), hp_t = (long long)d.Fc * (d.Tc + 7);
    const long long hp = B * (hp_f > hp_t ? hp_f : hp_t) * 64 + 16 * 64;
    long long sz[RTFS_BW_COUNT];  // floats
    for (int i = 0; i < RTFS_BW_COUNT; ++i) sz[i] = G;
    sz[RTFS_BW_DA] = sz[RTFS_BW_DB] = sz[RTFS_BW_DA1] = sz[RTFS_BW_DM] = sz[RTFS_BW_DA0] = A;
    sz[RTFS_BW_DSPEC] = B * d.P * 2;
    sz[RTFS_BW_HE] = sz[RTFS_BW_HDE] = sz[RTFS_BW_HF0] = sz[RTFS_BW_HDF0] = sz[RTFS_BW_HT] = H;
    sz[RTFS_BW_GDQ] = sz[RTFS_BW_GDK] = B * d.Pc * 16;
    sz[RTFS_BW_GDPRE] = B * d.Pc * 96;
    sz[RTFS_BW_DZP] = hp;
    sz[RTFS_BW_DHA] = sz[RTFS_BW_DHB] = hp;
    sz[RTFS_BW_DU] = B * d.Pc * 256 + 8 * 256;
    sz[RTFS_BW_DXIN] = G + 8 * 64;
    sz[RTFS_BW_DXUNF] = B * d.Pc * 512 + 8 * 512;
    sz[RTFS_BW_RED] = 32 * B * 2 * 2;
    sz[RTFS_BW_DVK] = sz[RTFS_BW_DATT] = B * (long long)(d.Tv > 0 ? d.Tv : 1) * 256;
    sz[RTFS_BW_CSUM] = 256 * 4 * 2;
    BwdPlan p;
    long long o = 0;
    for (int i = 0; i < RTFS_BW_COUNT; ++i) {
        p.off[i] = o;
        o += align256(sz[i] * 4);
    }
    p.total = o;
    return p;
}

struct TrainCtx {
    const float* const* P;
    float* const* G;  // gradient table (may be null for forward-only use)
    Dims d;
    cudaStream_t st;
    char* tape;
    TapePlan tp;
    char* scr;
    BwdPlan bp;
    int R;
    float* tbuf(int i) const { return reinterpret_cast<float*>(tape + tp.off[i]); }
    float* xin(int pass) const {  // block input of pass `pass`
        return pass == 0 ? tbuf(RTFS_TP_A1) : tbuf(RTFS_TP_X) + (long long)(pass - 1) * d.B * d.P * 256;
    }
    char* pass_ws(int pass) const { return tape + tp.off[RTFS_TP_PASS0] + (long long)pass * tp.pass_bytes; }
    float* sbuf(int i) const { return reinterpret_cast<float*>(scr + bp.off[i]); }
    double* red(int unit) const { return reinterpret_cast<double*>(scr + bp.off[RTFS_BW_RED]) + (long long)unit * d.B * 2; }
    float* grad(int slot) const { return G ? G[slot] : nullptr; }
};

Ctx pass_ctx(const TrainCtx& t, char* ws) {
    Ctx c;
    c.P = t.P;
    c.d = t.d;
    c.pl = make_plan(t.d, true);
    c.ws = ws;
    c.st = t.st;
    return c;
}

inline unsigned grid_for(long long work_items, int per_sm = 8) {
    long long blocks = (work_items + 255) / 256;
    const long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

#define NEED_GRAD(slot)                                                                             \
    do {                                                                                            \
        if (t.grad(slot) == nullptr) return fail_msg("backward: gradient buffer missing for slot " #slot); \
    } while (0)

// gLN backward unit: dy (in) -> dn (in place) -> dst (=|+=) gradient w.r.t. the pre-normalisation tensor
template <int C, int ACT>
int gln_unit_bwd(const TrainCtx& t, float* dy, const float* x_pre, const GlnRef& gln, const float* slope, int unit, float* dgamma, float* dbeta,
                 float* dslope, long long n_per_sample, float* dst, int accumulate) {
    const long long n4 = n_per_sample / 4;
    long long chunks = (n4 + 255) / 256;
    const long long cap = ((long long)sm_count() * 8 + t.d.B - 1) / t.d.B;
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    GlnBwdArgs a{dy, x_pre, gln, slope, t.red(unit), dgamma, dbeta, dslope, n4};
    gln_bwd_reduce_kernel<C, ACT><<<dim3((unsigned)chunks, t.d.B), 256, 0, t.st>>>(a);
    CK(cudaGetLastError());
    gln_bwd_apply_kernel<C><<<dim3((unsigned)chunks, t.d.B), 256, 0, t.st>>>(dy, x_pre, gln, t.red(unit), dst, accumulate, n4);
    CK(cudaGetLastError());
    return 0;
}

template <int C, int ACT>
int gln_apply(const TrainCtx& t, const float* x, const GlnRef& gln, const float* slope, float* y, long long n_per_sample) {
    const long long n4 = n_per_sample / 4;
    long long chunks = (n4 + 255) / 256;
    const long long cap = ((long long)sm_count() * 8 + t.d.B - 1) / t.d.B;
    if (chunks > cap) chunks = cap;
    gln_apply_kernel<C, ACT><<<dim3((unsigned)chunks, t.d.B), 256, 0, t.st>>>(x, gln, slope, y, n4);
    CK(cudaGetLastError());
    return 0;
}

int dw_wgrad(const TrainCtx& t, const float* dy, const float* x, float* dw, float* dbias, int Ti, int Fi, int To, int Fo, int stride) {
    DwBwdWArgs a{dy, x, dw, dbias, t.d.B, Ti, Fi, To, Fo, stride};
    const long long npos = (long long)t.d.B * To * Fo;
    long long blocks = (npos + 15) / 16;
    const long long cap = (long long)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    dw_bwd_weight_kernel<<<(unsigned)blocks, 256, 0, t.st>>>(a);
    CK(cudaGetLastError());
    return 0;
}

template <int NW>
int dw_dgrad(const TrainCtx& t, DwBwdDataArgs<NW> a) {
    a.total4 = (long long)t.d.B * a.Ti * a.Fi * 16;
    dw_bwd_data_kernel<NW><<<grid_for(a.total4, 16), 256, 0, t.st>>>(a);
    CK(cudaGetLastError());
    return 0;
}

template <int C>
int colsum(const TrainCtx& t, const float* x, long long rows, float* out) {
    colsum_kernel<C><<<grid_for(rows * (C / 4), 4), 256, 0, t.st>>>(x, rows, out);
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ dual-path RNN
struct RnnTape {
    float *n, *u[4], *c[4], *h[3], *hp;
};
RnnTape rnn_tape(const Ctx& c, int which) {
    const int base = which == 0 ? RTFS_WS_TF_N : RTFS_WS_TT_N;
    RnnTape r;
    r.n = c.buf(base);
    for (int l = 0; l < 4; ++l) r.u[l] = c.buf(base + 1 + l);
    for (int l = 0; l < 4; ++l) r.c[l] = c.buf(base + 5 + l);
    for (int l = 0; l < 3; ++l) r.h[l] = c.buf(base + 9 + l);
    r.hp = c.buf(base + 12);
    return r;
}

// DualPathRNN forward with the tape (the unfused kernel chain of run_dprnn, every intermediate kept)
int run_dprnn_train(const Ctx& c, int which, bool first, const float* g_in, float* g_first, float* g_out) {
    const Dims& d = c.d;
    const int base = which == 0 ? RTFS_P_RF_LNG : RTFS_P_RT_LNG;
    const int basei = which == 0 ? RTFS_P_RF_WI0 : RTFS_P_RT_WI0;
    const int S = which == 0 ? d.Fc : d.Tc;
    const int n_other = which == 0 ? d.Tc : d.Fc;
    const int nseq = d.B * n_other;
    const int L = S - 7;
    if (L < 1) return fail_msg("dual-path RNN needs at least 8 steps along the scanned axis");
    const int M = nseq * S;
    const RnnTape tp = rnn_tape(c, which);
    // rows past the end that the overlapping GEMM views touch must be finite
    CKN(cudaMemsetAsync(tp.n + (long long)M * 64, 0, 8 * 64 * 4, c.st));
    CKN(cudaMemsetAsync(tp.hp + (long long)nseq * (S + 7) * 64, 0, 8 * 64 * 4, c.st));
    PrepArgs pa;
    pa.g_in = g_in;
    pa.d1_pre = c.buf(RTFS_WS_D1_PRE);
    pa.pool = c.buf(RTFS_WS_POOL);
    pa.gln = c.gln(RTFS_ST_D1, RTFS_P_D1_GAMMA, RTFS_P_D1_BETA, d.Pc * 64);
    pa.ln_gamma = c.P[base + 0];
    pa.ln_beta = c.P[base + 1];
    pa.g_out = g_first;
    pa.n_out = tp.n;
    pa.B = d.B;
    pa.Tc = d.Tc;
    pa.Fc = d.Fc;
    pa.time_path = which;
    pa.first = first ? 1 : 0;
    const long long npos = d.B * d.Pc;
    dprnn_prep_kernel<<<(unsigned)((npos + 15) / 16), 256, 0, c.st>>>(pa);
    CK(cudaGetLastError());
    const float* resid = first ? g_first : g_in;
    {
        StoreEpi4 ep{tp.u[0], 256, nullptr};
        CK((launch_gemm_tc_unfold<256, 3, 1>(tp.n, c.P[basei + 0], ep, M, c.st)));
        ScanArgs sa{tp.u[0], 256, nullptr, c.P[base + 3], c.P[base + 4], tp.h[0], nseq, S, L, 4, S, 0, 0, tp.c[0]};
        sru_scan_kernel<<<(nseq + 3) / 4, 256, 0, c.st>>>(sa);
        CK(cudaGetLastError());
    }
    for (int l = 1; l <= 3; ++l) {
        const int pw = base + 2 + 3 * l;
        const float* hin = tp.h[l - 1];
        PlainLoader al{hin, 64, 64};
        StoreEpi4 ep{tp.u[l], 192, nullptr};
        CK((launch_gemm_tc<192, 64, 2, 2, 2, 256>(al, c.P[basei + l], ep, M, c.st)));
        const bool last = l == 3;
        float* hout = last ? tp.hp : tp.h[l];
        ScanArgs sa{tp.u[l], 192, hin, c.P[pw + 1], c.P[pw + 2], hout, nseq, S, L, 3, last ? S + 7 : S, last ? 7 : 0, last ? 1 : 0, tp.c[l]};
        sru_scan_kernel<<<(nseq + 3) / 4, 256, 0, c.st>>>(sa);
        CK(cudaGetLastError());
    }
    ConvTEpi4 ep{g_out, resid, c.P[base + 15], S, n_other, which, d.Tc, d.Fc};
    CK((launch_gemm_tc_unfold<64, 4, 2>(tp.hp, c.P[basei + 4], ep, nseq * (S + 7), c.st)));
    return 0;
}

// backward of one DualPathRNN: d_out (natural layout) -> d_in ; parameter gradients accumulated
int run_dprnn_bwd(const TrainCtx& t, const Ctx& c, int which, const float* g_in, const float* d_out, float* d_in) {
    const Dims& d = c.d;
    const int base = which == 0 ? RTFS_P_RF_LNG : RTFS_P_RT_LNG;
    const int baset = which == 0 ? RTFS_P_RF_W0T : RTFS_P_RT_W0T;  // W0T, W1T, W2T, W3T, CTWB
    const int S = which == 0 ? d.Fc : d.Tc;
    const int n_other = which == 0 ? d.Tc : d.Fc;
    const int nseq = d.B * n_other;
    const int L = S - 7;
    const int M = nseq * S, Mp = nseq * (S + 7);
    const RnnTape tp = rnn_tape(c, which);
    for (int i = 0; i < 16; ++i) NEED_GRAD(base + i);
    for (int i = 0; i < 5; ++i)
        if (t.P[baset + i] == nullptr) return fail_msg("backward: transposed SRU weight images missing (training-only parameter slots)");
    float *dzp = t.sbuf(RTFS_BW_DZP), *dhA = t.sbuf(RTFS_BW_DHA), *dhB = t.sbuf(RTFS_BW_DHB), *dU = t.sbuf(RTFS_BW_DU);
    float *dxin = t.sbuf(RTFS_BW_DXIN), *dxunf = t.sbuf(RTFS_BW_DXUNF);
    // 1. padded sequence-major copy of d_out (+ zero tail rows for the overlapping view)
    seq_pad_kernel<<<grid_for((long long)Mp * 16, 16), 256, 0, t.st>>>(d_out, dzp, d.B, d.Tc, d.Fc, which);
    CK(cudaGetLastError());
    CKN(cudaMemsetAsync(dzp + (long long)Mp * 64, 0, 16 * 64 * 4, t.st));
    // 2. ConvTranspose1d: dh_3[l] = sum_tap W[:, :, tap] dz[l + tap] ; dW_ct ; db_ct
    {
        PlainLoader al{dzp, 64, 512};
        StoreEpi ep{dhA, 64, nullptr};
        CK((launch_gemm<64, 512, false>(al, t.P[baset + 4], ep, Mp, 64, t.st)));
        PlainLoader xl{tp.hp, 64, 512}, yl{dzp, 64, 64};
        CK((launch_wgrad<false>(xl, yl, t.grad(base + 14), 512, Mp, 64, 512, t.st)));
        RUN(colsum<64>(t, dzp, Mp, t.grad(base + 15)));
    }
    // 3. SRU layers 3..1 (identity highway)
    float* dh = dhA;
    int dh_stride = S + 7;
    for (int l = 3; l >= 1; --l) {
        const int pw = base + 2 + 3 * l;
        ScanBwdArgs sa{tp.u[l], 192, tp.c[l], tp.h[l - 1], dh, dh_stride, t.P[pw + 1], t.P[pw + 2], dU, dxin, t.grad(pw + 1), t.grad(pw + 2), nseq, S, L, 3};
        sru_scan_bwd_kernel<<<(nseq + 3) / 4, 256, 0, t.st>>>(sa);
        CK(cudaGetLastError());
        PlainLoader xl{tp.h[l - 1], 64, 64}, yl{dU, 192, 192};
        CK((launch_wgrad<false>(xl, yl, t.grad(pw), 64, M, 192, 64, t.st)));
        float* dprev = dh == dhA ? dhB : dhA;
        PlainLoader al{dU, 192, 192};
        AddEpi ep{dprev, 64, dxin};
        CK((launch_gemm<64, 192, false>(al, t.P[baset + l], ep, M, 64, t.st)));
        dh = dprev;
        dh_stride = S;
    }
    // 4. layer 0 (k = 4: highway through the fourth gate column) ; unfold o Linear
    {
        const int pw = base + 2;
        ScanBwdArgs sa{tp.u[0], 256, tp.c[0], nullptr, dh, dh_stride, t.P[pw + 1], t.P[pw + 2], dU, nullptr, t.grad(pw + 1), t.grad(pw + 2), nseq, S, L, 4};
        sru_scan_bwd_kernel<<<(nseq + 3) / 4, 256, 0, t.st>>>(sa);
        CK(cudaGetLastError());
        PlainLoader xl{tp.n, 64, 512}, yl{dU, 256, 256};
        CK((launch_wgrad<false>(xl, yl, t.grad(pw), 512, M, 256, 512, t.st)));
        PlainLoader al{dU, 256, 256};
        StoreEpi ep{dxunf, 512, nullptr};
        CK((launch_gemm<128, 256, false>(al, t.P[baset + 0], ep, M, 512, t.st)));
    }
    // 5. fold + LayerNorm backward + residual
    LnBwdArgs la{dxunf, g_in, d_out, t.P[base + 0], d_in, t.grad(base + 0), t.grad(base + 1), d.B, d.Tc, d.Fc, which, L};
    const long long npos = (long long)d.B * d.Pc;
    long long blocks = (npos + 15) / 16;
    if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
    dprnn_ln_bwd_kernel<<<(unsigned)blocks, 256, 0, t.st>>>(la);
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ attention
int run_mhsa_bwd(const TrainCtx& t, const Ctx& c, const float* g_in, const float* d_out, float* d_in) {
    const Dims& d = c.d;
    const int H = 4, Tc = d.Tc;
    const int nframes = d.B * Tc;
    const int M = nframes * 64;
    if ((long long)Tc * Tc * H > (long long)d.Pc * 64) return fail_msg("attention backward: too many frames for the score scratch");
    const int slots[] = {RTFS_P_AT_WQKV, RTFS_P_AT_BQKV, RTFS_P_AT_SLOPE, RTFS_P_AT_GAMMA, RTFS_P_AT_BETA,
                         RTFS_P_AT_WO,   RTFS_P_AT_BO,   RTFS_P_AT_SLOPEO, RTFS_P_AT_GAMMAO, RTFS_P_AT_BETAO};
    for (int s : slots) NEED_GRAD(s);
    if (t.P[RTFS_P_AT_WQKVT] == nullptr || t.P[RTFS_P_AT_WOT] == nullptr) return fail_msg("backward: transposed attention weight images missing");
    float *q = c.buf(RTFS_WS_Q), *k = c.buf(RTFS_WS_K), *v = c.buf(RTFS_WS_V), *ao = c.buf(RTFS_WS_AO);
    float *ga1 = t.sbuf(RTFS_BW_GA1), *ga2 = t.sbuf(RTFS_BW_GA2), *ga3 = t.sbuf(RTFS_BW_GA3), *gs = t.sbuf(RTFS_BW_GS), *gdp = t.sbuf(RTFS_BW_GDP);
    float *gdq = t.sbuf(RTFS_BW_GDQ), *gdk = t.sbuf(RTFS_BW_GDK), *gdpre = t.sbuf(RTFS_BW_GDPRE);
    const int frame_grid = nframes < sm_count() * 2 ? nframes : sm_count() * 2;
    // 1. concat projection: LN(C,F) + PReLU backward
    {
        AttProjBwdArgs a{ao, d_out, t.P[RTFS_P_AT_WO], t.P[RTFS_P_AT_BO], t.P[RTFS_P_AT_SLOPEO], t.P[RTFS_P_AT_GAMMAO], ga1,
                         t.grad(RTFS_P_AT_GAMMAO), t.grad(RTFS_P_AT_BETAO), t.grad(RTFS_P_AT_SLOPEO), nframes};
        const int smem = (2 * 64 * 65 + 64) * 4;
        static SmemCfg cfg;
        CKN(ensure_smem(att_proj_bwd_kernel, smem, cfg));
        att_proj_bwd_kernel<<<frame_grid, 256, smem, t.st>>>(a);
        CK(cudaGetLastError());
        PlainLoader al{ga1, 64, 64};
        StoreEpi ep{ga2, 64, nullptr};
        CK((launch_gemm<64, 64, false>(al, t.P[RTFS_P_AT_WOT], ep, M, 64, t.st)));
        PlainLoader xl{ao, 64, 64}, yl{ga1, 64, 64};
        CK((launch_wgrad<false>(xl, yl, t.grad(RTFS_P_AT_WO), 64, M, 64, 64, t.st)));
        RUN(colsum<64>(t, ga1, M, t.grad(RTFS_P_AT_BO)));
    }
    // 2. per-head token rows of dO ; scores and their gradients
    att_regroup_kernel<<<grid_for((long long)M * 16, 16), 256, 0, t.st>>>(ga2, ga3, d.B, Tc, H);
    CK(cudaGetLastError());
    const int BH = d.B * H;
    const float scale = 1.f / sqrtf(4.f * 64.f);
    {
        BgemmArgs s{q, k, gs, (long long)Tc * 256, (long long)Tc * 256, (long long)Tc * Tc, 256, 1, 1, 256, Tc, Tc, Tc, 256, scale};
        CK(launch_bgemm(s, BH, t.st));
        BgemmArgs p{ga3, v, gdp, (long long)Tc * 1024, (long long)Tc * 1024, (long long)Tc * Tc, 1024, 1, 1, 1024, Tc, Tc, Tc, 1024, 1.f};
        CK(launch_bgemm(p, BH, t.st));
        const long long rows = (long long)BH * Tc;
        softmax_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, t.st>>>(gs, gdp, rows, Tc, scale);
        CK(cudaGetLastError());
        // dQ = dS K ; dK = dS^T Q ; dV = P^T dO
        BgemmArgs dq{gdp, k, gdq, (long long)Tc * Tc, (long long)Tc * 256, (long long)Tc * 256, Tc, 1, 256, 1, 256, Tc, 256, Tc, 1.f};
        CK(launch_bgemm(dq, BH, t.st));
        BgemmArgs dk{gdp, q, gdk, (long long)Tc * Tc, (long long)Tc * 256, (long long)Tc * 256, 1, Tc, 256, 1, 256, Tc, 256, Tc, 1.f};
        CK(launch_bgemm(dk, BH, t.st));
        BgemmArgs dv{gs, ga3, ga2, (long long)Tc * Tc, (long long)Tc * 1024, (long long)Tc * 1024, 1, Tc, 1024, 1, 1024, Tc, 1024, Tc, 1.f};
        CK(launch_bgemm(dv, BH, t.st));
    }
    // 3. head convs: LN(E,F) + PReLU backward, then the 1x1 convs
    {
        AttQkvBwdArgs a{g_in, t.P[RTFS_P_AT_WQKV], t.P[RTFS_P_AT_BQKV], t.P[RTFS_P_AT_SLOPE], t.P[RTFS_P_AT_GAMMA], gdq, gdk, ga2, gdpre,
                        t.grad(RTFS_P_AT_GAMMA), t.grad(RTFS_P_AT_BETA), t.grad(RTFS_P_AT_SLOPE), nframes, Tc, H};
        const int smem = (64 * 65 + 96 * 65 + 768 + 16) * 4;
        static SmemCfg cfg;
        CKN(ensure_smem(att_qkv_bwd_kernel, smem, cfg));
        att_qkv_bwd_kernel<<<frame_grid, 384, smem, t.st>>>(a);
        CK(cudaGetLastError());
        PlainLoader al{gdpre, 96, 96};
        AddEpi ep{d_in, 64, d_out};  // + the residual path (out = proj + g)
        CK((launch_gemm<64, 96, false>(al, t.P[RTFS_P_AT_WQKVT], ep, M, 64, t.st)));
        PlainLoader xl{g_in, 64, 64}, yl{gdpre, 96, 96};
        CK((launch_wgrad<false>(xl, yl, t.grad(RTFS_P_AT_WQKV), 64, M, 96, 64, t.st)));
        RUN(colsum<96>(t, gdpre, M, t.grad(RTFS_P_AT_BQKV)));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ RTFS block
int run_block_bwd(const TrainCtx& t, const Ctx& c, const float* x, const float* dout, float* dx, float* dacc, bool dacc_init) {
    const Dims& d = c.d;
    const float* const* P = c.P;
    const int M = (int)(d.B * d.P);
    const long long nfull = d.P * 64, ncomp = d.Pc * 64;
    float *p_pre = c.buf(RTFS_WS_P_PRE), *d0_pre = c.buf(RTFS_WS_D0_PRE), *d1_pre = c.buf(RTFS_WS_D1_PRE);
    float *g0 = c.buf(RTFS_WS_G0), *g1 = c.buf(RTFS_WS_G1), *g2 = c.buf(RTFS_WS_G2), *g3 = c.buf(RTFS_WS_G3);
    float *le0 = c.buf(RTFS_WS_LE0_PRE), *lec = c.buf(RTFS_WS_LEC_PRE), *le1 = c.buf(RTFS_WS_LE1);
    float *ge0 = c.buf(RTFS_WS_GE0), *gg0 = c.buf(RTFS_WS_GG0), *ge1 = c.buf(RTFS_WS_GE1), *gg1 = c.buf(RTFS_WS_GG1);
    float *gec = c.buf(RTFS_WS_GEC), *ggc = c.buf(RTFS_WS_GGC);
    float *hE = t.sbuf(RTFS_BW_HE), *hDE = t.sbuf(RTFS_BW_HDE), *hF0 = t.sbuf(RTFS_BW_HF0), *hDF0 = t.sbuf(RTFS_BW_HDF0), *hT = t.sbuf(RTFS_BW_HT);
    float *gF1 = t.sbuf(RTFS_BW_GF1), *gDF1 = t.sbuf(RTFS_BW_GDF1), *gT1 = t.sbuf(RTFS_BW_GT1), *gT2 = t.sbuf(RTFS_BW_GT2), *gT3 = t.sbuf(RTFS_BW_GT3);
    float *gT4 = t.sbuf(RTFS_BW_GT4), *gT5 = t.sbuf(RTFS_BW_GT5), *gD1N = t.sbuf(RTFS_BW_GD1N), *gDD1N = t.sbuf(RTFS_BW_GDD1N);
    float *gDG3 = t.sbuf(RTFS_BW_GDG3), *gDG2 = t.sbuf(RTFS_BW_GDG2), *gDG1 = t.sbuf(RTFS_BW_GDG1), *gDG0 = t.sbuf(RTFS_BW_GDG0);
    for (int s = RTFS_P_GW_W; s <= RTFS_P_D1_BETA; ++s) NEED_GRAD(s);
    for (int s = RTFS_P_F0_LW; s <= RTFS_P_RC_B; ++s) NEED_GRAD(s);
    if (P[RTFS_P_PJ_WT] == nullptr || P[RTFS_P_RC_WT] == nullptr) return fail_msg("backward: transposed 1x1-conv weight images missing");
    CKN(cudaMemsetAsync(t.red(0), 0, sizeof(double) * 2 * d.B * 32, t.st));
    const GlnRef nC0L = c.gln(RTFS_ST_C0L, RTFS_P_C0_LG, RTFS_P_C0_LB, nfull), nC0G = c.gln(RTFS_ST_C0G, RTFS_P_C0_GG, RTFS_P_C0_GB, ncomp);
    const GlnRef nC0E = c.gln(RTFS_ST_C0E, RTFS_P_C0_EG, RTFS_P_C0_EB, ncomp), nD0 = c.gln(RTFS_ST_D0, RTFS_P_D0_GAMMA, RTFS_P_D0_BETA, nfull);
    const GlnRef nF0L = c.gln(RTFS_ST_F0L, RTFS_P_F0_LG, RTFS_P_F0_LB, nfull), nF0G = c.gln(RTFS_ST_F0G, RTFS_P_F0_GG, RTFS_P_F0_GB, ncomp);
    const GlnRef nF0E = c.gln(RTFS_ST_F0E, RTFS_P_F0_EG, RTFS_P_F0_EB, ncomp), nF1L = c.gln(RTFS_ST_F1L, RTFS_P_F1_LG, RTFS_P_F1_LB, ncomp);
    const GlnRef nF1G = c.gln(RTFS_ST_F1G, RTFS_P_F1_GG, RTFS_P_F1_GB, ncomp), nF1E = c.gln(RTFS_ST_F1E, RTFS_P_F1_EG, RTFS_P_F1_EB, ncomp);
    const GlnRef nD1 = c.gln(RTFS_ST_D1, RTFS_P_D1_GAMMA, RTFS_P_D1_BETA, ncomp), nPJ = c.gln(RTFS_ST_PJ, RTFS_P_PJ_GAMMA, RTFS_P_PJ_BETA, nfull);
    const long long tot_full = (long long)d.B * d.P * 16, tot_comp = (long long)d.B * d.Pc * 16;
    TfarArgs aC0{lec, ggc, gec, d0_pre, nC0L, nC0G, nC0E, nD0, d.T, d.F, d.Tc, d.Fc, tot_full};
    TfarArgs aF0{le0, gg0, ge0, nullptr, nF0L, nF0G, nF0E, nD0, d.T, d.F, d.Tc, d.Fc, tot_full};
    TfarArgs aF1{le1, gg1, ge1, nullptr, nF1L, nF1G, nF1E, nD1, d.Tc, d.Fc, d.Tc, d.Fc, tot_comp};

    // B1-B2: e (re-materialised) ; residual conv: de = dout W_r, dW_r, db_r                      tdanet.py:131
    tfar_fwd_kernel<<<grid_for(tot_full, 16), 256, 0, t.st>>>(aC0, hE);
    CK(cudaGetLastError());
    {
        PlainLoader al{dout, 256, 256};
        StoreEpi ep{hDE, 64, nullptr};
        CK((launch_gemm<64, 256, false>(al, P[RTFS_P_RC_WT], ep, M, 64, t.st)));
        PlainLoader xl{hE, 64, 64}, yl{dout, 256, 256};
        CK((launch_wgrad<false>(xl, yl, t.grad(RTFS_P_RC_W), 64, M, 256, 64, t.st)));
        RUN(colsum<256>(t, dout, M, t.grad(RTFS_P_RC_B)));
    }
    // B3-B5: concat_layers.0 = TF-AR(f0, f1) + d0                                                tdanet.py:127-129
    tfar_bwd_local_kernel<<<grid_for(tot_full, 16), 256, 0, t.st>>>(aC0, hDE, hT);
    CK(cudaGetLastError());
    tfar_bwd_global_kernel<<<grid_for(tot_comp, 16), 256, 0, t.st>>>(aC0, hDE, gT1, gT2);
    CK(cudaGetLastError());
    float* hDD0 = hDE;  // from here on hDE accumulates the gradient w.r.t. d0 = gLN(d0_pre) (the "+ d0" term starts it)
    RUN((gln_unit_bwd<64, ACT_NONE>(t, hT, lec, nC0L, nullptr, 0, t.grad(RTFS_P_C0_LG), t.grad(RTFS_P_C0_LB), nullptr, nfull, hT, 0)));
    RUN((gln_unit_bwd<64, ACT_NONE>(t, gT1, ggc, nC0G, nullptr, 1, t.grad(RTFS_P_C0_GG), t.grad(RTFS_P_C0_GB), nullptr, ncomp, gT1, 0)));
    RUN((gln_unit_bwd<64, ACT_NONE>(t, gT2, gec, nC0E, nullptr, 2, t.grad(RTFS_P_C0_EG), t.grad(RTFS_P_C0_EB), nullptr, ncomp, gT2, 0)));
    tfar_fwd_kernel<<<grid_for(tot_full, 16), 256, 0, t.st>>>(aF0, hF0);  // f0
    CK(cudaGetLastError());
    RUN(dw_wgrad(t, hT, hF0, t.grad(RTFS_P_C0_LW), nullptr, d.T, d.F, d.T, d.F, 1));
    {
        DwBwdDataArgs<1> a{{hT}, {P[RTFS_P_C0_LW]}, hDF0, 0, d.T, d.F, d.T, d.F, 1, 0};
        RUN(dw_dgrad<1>(t, a));
    }
    tfar_fwd_kernel<<<grid_for(tot_comp, 16), 256, 0, t.st>>>(aF1, gF1);  // f1
    CK(cudaGetLastError());
    RUN(dw_wgrad(t, gT2, gF1, t.grad(RTFS_P_C0_EW), nullptr, d.Tc, d.Fc, d.Tc, d.Fc, 1));
    RUN(dw_wgrad(t, gT1, gF1, t.grad(RTFS_P_C0_GW), nullptr, d.Tc, d.Fc, d.Tc, d.Fc, 1));
    {
        DwBwdDataArgs<2> a{{gT2, gT1}, {P[RTFS_P_C0_EW], P[RTFS_P_C0_GW]}, gDF1, 0, d.Tc, d.Fc, d.Tc, d.Fc, 1, 0};
        RUN(dw_dgrad<2>(t, a));
    }
    // B6-B8: fusion_layers.0 = TF-AR(d0, g), fusion_layers.1 = TF-AR(d1, g)                      tdanet.py:124-126
    tfar_bwd_local_kernel<<<grid_for(tot_full, 16), 256, 0, t.st>>>(aF0, hDF0, hT);
    CK(cudaGetLastError());
    tfar_bwd_global_kernel<<<grid_for(tot_comp, 16), 256, 0, t.st>>>(aF0, hDF0, gT1, gT2);
    CK(cudaGetLastError());
    tfar_bwd_local_kernel<<<grid_for(tot_comp, 16), 256, 0, t.st>>>(aF1, gDF1, gT3);
    CK(cudaGetLastError());
    tfar_bwd_global_kernel<<<grid_for(tot_comp, 16), 256, 0, t.st>>>(aF1, gDF1, gT4, gT5);
    CK(cudaGetLastError());
    RUN((gln_unit_bwd<64, ACT_NONE>(t, hT, le0, nF0L, nullptr, 3, t.grad(RTFS_P_F0_LG), t.grad(RTFS_P_F0_LB), nullptr, nfull, hT, 0)));
    RUN((gln_unit_bwd<64, ACT_NONE>(t, gT1, gg0, nF0G, nullptr, 4, t.grad(RTFS_P_F0_GG), t.grad(RTFS_P_F0_GB), nullptr, ncomp, gT1, 0)));
    RUN((gln_unit_bwd<64, ACT_NONE>(t, gT2, ge0, nF0E, nullptr, 5, t.grad(RTFS_P_F0_EG), t.grad(RTFS_P_F0_EB), nullptr, ncomp, gT2, 0)));
    RUN((gln_unit_bwd<64, ACT_NONE>(t, gT3, le1, nF1L, nullptr, 6, t.grad(RTFS_P_F1_LG), t.grad(RTFS_P_F1_LB), nullptr, ncomp, gT3, 0)));
    RUN((gln_unit_bwd<64, ACT_NONE>(t, gT4, gg1, nF1G, nullptr, 7, t.grad(RTFS_P_F1_GG), t.grad(RTFS_P_F1_GB), nullptr, ncomp, gT4, 0)));
    RUN((gln_unit_bwd<64, ACT_NONE>(t, gT5, ge1, nF1E, nullptr, 8, t.grad(RTFS_P_F1_EG), t.grad(RTFS_P_F1_EB), nullptr, ncomp, gT5, 0)));
    RUN((gln_apply<64, ACT_NONE>(t, d0_pre, nD0, nullptr, hF0, nfull)));  // d0 = gLN(d0_pre)
    RUN(dw_wgrad(t, hT, hF0, t.grad(RTFS_P_F0_LW), nullptr, d.T, d.F, d.T, d.F, 1));
    {
        DwBwdDataArgs<1> a{{hT}, {P[RTFS_P_F0_LW]}, hDD0, 1, d.T, d.F, d.T, d.F, 1, 0};
        RUN(dw_dgrad<1>(t, a));
    }
    RUN(dw_wgrad(t, gT2, g3, t.grad(RTFS_P_F0_EW), nullptr, d.Tc, d.Fc, d.Tc, d.Fc, 1));
    RUN(dw_wgrad(t, gT1, g3, t.grad(RTFS_P_F0_GW), nullptr, d.Tc, d.Fc, d.Tc, d.Fc, 1));
    RUN(dw_wgrad(t, gT5, g3, t.grad(RTFS_P_F1_EW), nullptr, d.Tc, d.Fc, d.Tc, d.Fc, 1));
    RUN(dw_wgrad(t, gT4, g3, t.grad(RTFS_P_F1_GW), nullptr, d.Tc, d.Fc, d.Tc, d.Fc, 1));
    {
        DwBwdDataArgs<4> a{{gT2, gT1, gT5, gT4}, {P[RTFS_P_F0_EW], P[RTFS_P_F0_GW], P[RTFS_P_F1_EW], P[RTFS_P_F1_GW]}, gDG3, 0, d.Tc, d.Fc, d.Tc, d.Fc, 1, 0};
        RUN(dw_dgrad<4>(t, a));
    }
    RUN((gln_apply<64, ACT_NONE>(t, d1_pre, nD1, nullptr, gD1N, ncomp)));  // d1 = gLN(d1_pre)
    RUN(dw_wgrad(t, gT3, gD1N, t.grad(RTFS_P_F1_LW), nullptr, d.Tc, d.Fc, d.Tc, d.Fc, 1));
    {
        DwBwdDataArgs<1> a{{gT3}, {P[RTFS_P_F1_LW]}, gDD1N, 0, d.Tc, d.Fc, d.Tc, d.Fc, 1, 0};
        RUN(dw_dgrad<1>(t, a));
    }
    // B9-B10: global attention stack                                                             tdanet.py:121
    RUN(run_mhsa_bwd(t, c, g2, gDG3, gDG2));
    RUN(run_dprnn_bwd(t, c, 1, g1, gDG2, gDG1));
    RUN(run_dprnn_bwd(t, c, 0, g0, gDG1, gDG0));
    // B11: g0 = gLN(d1_pre) + adaptive_avg_pool2d(d0)                                             tdanet.py:117-118
    add_kernel<<<grid_for(tot_comp, 16), 256, 0, t.st>>>(gDD1N, gDG0, gDD1N, tot_comp);
    CK(cudaGetLastError());
    pool_bwd_kernel<<<grid_for(tot_full, 16), 256, 0, t.st>>>(gDG0, hDD0, d.T, d.F, d.Tc, d.Fc, tot_full);
    CK(cudaGetLastError());
    // B12: downsample_layers.1 (stride 2) on d0                                                   tdanet.py:69-76
    RUN((gln_unit_bwd<64, ACT_NONE>(t, gDD1N, d1_pre, nD1, nullptr, 9, t.grad(RTFS_P_D1_GAMMA), t.grad(RTFS_P_D1_BETA), nullptr, ncomp, gDD1N, 0)));
    RUN(dw_wgrad(t, gDD1N, hF0, t.grad(RTFS_P_D1_W), t.grad(RTFS_P_D1_B), d.T, d.F, d.Tc, d.Fc, 2));
    {
        DwBwdDataArgs<1> a{{gDD1N}, {P[RTFS_P_D1_W]}, hDD0, 1, d.T, d.F, d.Tc, d.Fc, 2, 0};
        RUN(dw_dgrad<1>(t, a));
    }
    // B13: downsample_layers.0 on p = PReLU(gLN(p_pre))                                           tdanet.py:61-68
    RUN((gln_unit_bwd<64, ACT_NONE>(t, hDD0, d0_pre, nD0, nullptr, 10, t.grad(RTFS_P_D0_GAMMA), t.grad(RTFS_P_D0_BETA), nullptr, nfull, hDD0, 0)));
    RUN((gln_apply<64, ACT_PRELU>(t, p_pre, nPJ, P[RTFS_P_PJ_A], hF0, nfull)));
    RUN(dw_wgrad(t, hDD0, hF0, t.grad(RTFS_P_D0_W), t.grad(RTFS_P_D0_B), d.T, d.F, d.T, d.F, 1));
    {
        DwBwdDataArgs<1> a{{hDD0}, {P[RTFS_P_D0_W]}, hT, 0, d.T, d.F, d.T, d.F, 1, 0};
        RUN(dw_dgrad<1>(t, a));
    }
    // B14-B15: projection (gLN + PReLU), gateway                                                  tdanet.py:34-49,107-108
    RUN((gln_unit_bwd<64, ACT_PRELU>(t, hT, p_pre, nPJ, P[RTFS_P_PJ_A], 11, t.grad(RTFS_P_PJ_GAMMA), t.grad(RTFS_P_PJ_BETA), t.grad(RTFS_P_PJ_A), nfull, hT, 0)));
    {
        GateLoader xl{x, P[RTFS_P_GW_W], P[RTFS_P_GW_B], P[RTFS_P_GW_A], 256};
        PlainLoader yl{hT, 64, 64};
        CK((launch_wgrad<false>(xl, yl, t.grad(RTFS_P_PJ_W), 256, M, 64, 256, t.st)));
        RUN(colsum<64>(t, hT, M, t.grad(RTFS_P_PJ_B)));
        PlainLoader al{hT, 64, 64};
        GateBwdEpi ep{dx, dout, x, P[RTFS_P_GW_W], P[RTFS_P_GW_B], P[RTFS_P_GW_A], t.grad(RTFS_P_GW_W), t.grad(RTFS_P_GW_B), t.grad(RTFS_P_GW_A), dacc, dacc_init ? 1 : 0};
        CK((launch_gemm<64, 64, false>(al, P[RTFS_P_PJ_WT], ep, M, 256, t.st)));
    }
    return 0;
}

bool make_train_ctx(TrainCtx& t, const float* const* params, float* const* grads, void* tape, void* scratch, int B, int T, int Tv, int R, void* stream) {
    if (params == nullptr || B < 1 || T < 16) {
        g_err = "bad arguments (params null, B < 1 or fewer than 16 frames)";
        return false;
    }
    t.P = params;
    t.G = grads;
    t.d = make_dims(B, T, Tv);
    if ((long long)B * t.d.P * 256 >= (1ll << 31)) {
        g_err = "batch too large for 32-bit row indexing (B*T*F*256 must be < 2^31 elements per call)";
        return false;
    }
    t.st = reinterpret_cast<cudaStream_t>(stream);
    t.tape = reinterpret_cast<char*>(tape);
    t.R = R;
    t.tp = make_tape_plan(t.d, R);
    t.scr = reinterpret_cast<char*>(scratch);
    t.bp = make_bwd_plan(t.d);
    return true;
}

}  // namespace

extern "C" {

long long rtfs_train_plan(int B, int L, int Tv, int repeats, long long* tape_offsets, long long* pass_bytes, long long* bwd_bytes, long long* bwd_offsets) {
    const Dims d = make_dims(B, L / 128 + 1, Tv);
    const TapePlan tp = make_tape_plan(d, repeats);
    const BwdPlan bp = make_bwd_plan(d);
    if (tape_offsets)
        for (int i = 0; i < RTFS_TP_COUNT; ++i) tape_offsets[i] = tp.off[i];
    if (pass_bytes) *pass_bytes = tp.pass_bytes;
    if (bwd_bytes) *bwd_bytes = bp.total;
    if (bwd_offsets)
        for (int i = 0; i < RTFS_BW_COUNT; ++i) bwd_offsets[i] = bp.off[i];
    return tp.total;
}

long long rtfs_train_pass_plan(int B, int L, long long* offsets) {
    const Dims d = make_dims(B, L / 128 + 1, 0);
    const Plan p = make_plan(d, true);
    if (offsets)
        for (int i = 0; i < RTFS_WS_COUNT; ++i) offsets[i] = p.off[i];
    return p.total;
}

int rtfs_avnet_train_forward(const float* const* params, const float* wav, const float* video, float* out, void* tape,
                             int B, int L, int Tv, int repeats, int phase, void* stream) {
    TrainCtx t;
    if (!make_train_ctx(t, params, nullptr, tape, nullptr, B, L / 128 + 1, Tv, repeats, stream)) return -2;
    if (repeats < 1 || Tv < 1) return fail_msg("rtfs_avnet_train_forward: repeats < 1 or Tv < 1");
    const long long rows = (long long)B * t.d.P;
    float *a0 = t.tbuf(RTFS_TP_A0), *a1 = t.tbuf(RTFS_TP_A1), *blk0 = t.tbuf(RTFS_TP_BLK0), *refined = t.tbuf(RTFS_TP_REFINED);
    // a Ctx whose SPEC / Q18 / VK / ATT / statistics live in the global tape: the encoder, CAF and decoder stages address their
    // buffers through Ctx::buf, so they get a small plan of their own over the pass-0 region + the global buffers
    Ctx c0 = pass_ctx(t, t.pass_ws(0));
    c0.d.Tv = Tv;
    c0.pl.off[RTFS_WS_SPEC] = (t.tape + t.tp.off[RTFS_TP_SPEC]) - c0.ws;
    c0.pl.off[RTFS_WS_Q18] = (t.tape + t.tp.off[RTFS_TP_Q18]) - c0.ws;
    c0.pl.off[RTFS_WS_VK] = (t.tape + t.tp.off[RTFS_TP_VK]) - c0.ws;
    c0.pl.off[RTFS_WS_ATT] = (t.tape + t.tp.off[RTFS_TP_ATT]) - c0.ws;
    if (phase == 0) {
        g_launches = 0;
        RUN(run_encoder(c0, wav, a0, L));
        RUN(run_bottleneck(c0, a0, a1, false));
        c0.train = true;
        RUN(run_block(c0, a1, nullptr, blk0));
        if (video != nullptr) RUN(run_caf_video(c0, video));  // (NULL: the caller submits the video-dependent part as phase 2)
        double* cs = reinterpret_cast<double*>(t.tape + t.tp.off[RTFS_TP_CAFSUM]);
        CKN(cudaMemsetAsync(cs, 0, 256 * 2 * 8, t.st));
        chan_stats_kernel<<<grid_for(rows * 64, 4), 256, 0, t.st>>>(blk0, rows, cs);
        CK(cudaGetLastError());
        return 0;
    }
    if (phase == 2) {
        if (video == nullptr) return fail_msg("rtfs_avnet_train_forward(2): video missing");
        RUN(run_caf_video(c0, video));
        return 0;
    }
    // phase 1: CAF with the batch-statistics scale / shift the host stored in the SK / TK / SV / TV slots
    const Dims& d = t.d;
    float* caf_out = repeats > 1 ? t.xin(1) : refined;
    {
        CafApplyArgs aa{blk0, repeats > 1 ? a1 : nullptr, c0.buf(RTFS_WS_VK), c0.buf(RTFS_WS_ATT), params[RTFS_P_CAF_SK], params[RTFS_P_CAF_TK],
                        params[RTFS_P_CAF_SV], params[RTFS_P_CAF_TV], caf_out, d.T, d.F, 256, d.Tv, d.B * d.P * 64};
        long long blocks = (aa.total4 + 255) / 256;
        if (blocks > sm_count() * 16) blocks = sm_count() * 16;
        caf_apply_kernel<<<(unsigned)blocks, 256, 0, t.st>>>(aa);
        CK(cudaGetLastError());
    }
    for (int i = 1; i < repeats; ++i) {
        Ctx ci = pass_ctx(t, t.pass_ws(i));
        ci.train = true;
        float* o = i + 1 < repeats ? t.xin(i + 1) : refined;
        RUN(run_block(ci, t.xin(i), i + 1 < repeats ? a1 : nullptr, o));
    }
    RUN(run_mask(c0, refined, a0, t.tbuf(RTFS_TP_Z), t.tbuf(RTFS_TP_M)));
    RUN(run_decoder(c0, t.tbuf(RTFS_TP_Z), out, L));
    return 0;
}

int rtfs_avnet_backward(const float* const* params, float* const* grads, const float* wav, const float* video, const float* d_out,
                        float* d_video, const float* caf_mu, const float* caf_c0, const float* caf_c1, void* tape, void* scratch,
                        int B, int L, int Tv, int repeats, int phase, void* stream) {
    (void)wav;
    TrainCtx t;
    if (!make_train_ctx(t, params, grads, tape, scratch, B, L / 128 + 1, Tv, repeats, stream)) return -2;
    if (grads == nullptr || scratch == nullptr) return fail_msg("rtfs_avnet_backward: grads / scratch missing");
    const Dims& d = t.d;
    const int M = (int)(d.B * d.P);
    const long long totA4 = (long long)M * 64;
    float *a0 = t.tbuf(RTFS_TP_A0), *a1 = t.tbuf(RTFS_TP_A1), *blk0 = t.tbuf(RTFS_TP_BLK0), *refined = t.tbuf(RTFS_TP_REFINED);
    float *dA = t.sbuf(RTFS_BW_DA), *dB = t.sbuf(RTFS_BW_DB), *dA1 = t.sbuf(RTFS_BW_DA1), *dM = t.sbuf(RTFS_BW_DM), *dA0 = t.sbuf(RTFS_BW_DA0);
    float* vk = t.tbuf(RTFS_TP_VK);
    float* att = t.tbuf(RTFS_TP_ATT);
    CafBwdArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.a = blk0;
    ca.vk = vk;
    ca.att = att;
    ca.sk = params[RTFS_P_CAF_SK];
    ca.tk = params[RTFS_P_CAF_TK];
    ca.sv = params[RTFS_P_CAF_SV];
    ca.tv = params[RTFS_P_CAF_TV];
    ca.csum = reinterpret_cast<double*>(t.scr + t.bp.off[RTFS_BW_CSUM]);
    ca.dvk = t.sbuf(RTFS_BW_DVK);
    ca.datt = t.sbuf(RTFS_BW_DATT);
    ca.T = d.T;
    ca.F = d.F;
    ca.Tv = d.Tv;
    // after phase 0 the gradient w.r.t. the CAF output lives in dB when R is even... keep it simple: it always ends in dB
    if (phase == 0) {
        g_launches = 0;
        const int need[] = {RTFS_P_DEC_WE, RTFS_P_MK_A, RTFS_P_MK_W, RTFS_P_MK_B};
        for (int s : need) NEED_GRAD(s);
        if (params[RTFS_P_DEC_WE] == nullptr || params[RTFS_P_MK_WT] == nullptr) return fail_msg("backward: training-only parameter slots missing");
        // decoder: iSTFT adjoint, then dz = conv2d(dspec, W_dec) (the adjoint of the transposed conv)   decoder.py:110-128
        float* dspec = t.sbuf(RTFS_BW_DSPEC);
        IstftBwdArgs ia{d_out, params[RTFS_P_WINDOW], params[RTFS_P_COSTAB], params[RTFS_P_SINTAB], dspec, L, d.T};
        istft_bwd_kernel<<<dim3(d.T, d.B), 288, 0, t.st>>>(ia);
        CK(cudaGetLastError());
        {
            Im2colLoader al{dspec, d.T, d.F};
            StoreEpi ep{dM, 256, nullptr};  // dz
            CK((launch_gemm<128, 32, true>(al, params[RTFS_P_DEC_WE], ep, M, 256, t.st)));
            Im2colLoader xl{dspec, d.T, d.F};
            PlainLoader yl{t.tbuf(RTFS_TP_Z), 256, 256};
            CK((launch_wgrad<true>(xl, yl, t.grad(RTFS_P_DEC_WE), 32, M, 256, 32, t.st)));
        }
        // S^3 mask                                                                                  mask_generator.py:67-99
        mask_bwd_kernel<<<grid_for((long long)M * 32, 16), 256, 0, t.st>>>(dM, a0, t.tbuf(RTFS_TP_M), dM, dA0, M);
        CK(cudaGetLastError());
        {
            PlainLoader al{dM, 256, 256};
            PreluBwdEpi ep{dA, refined, params[RTFS_P_MK_A], t.grad(RTFS_P_MK_A)};
            CK((launch_gemm<128, 256, false>(al, params[RTFS_P_MK_WT], ep, M, 256, t.st)));
            PreluLoader xl{refined, params[RTFS_P_MK_A], 256};
            PlainLoader yl{dM, 256, 256};
            CK((launch_wgrad<false>(xl, yl, t.grad(RTFS_P_MK_W), 256, M, 256, 256, t.st)));
            RUN(colsum<256>(t, dM, M, t.grad(RTFS_P_MK_B)));
        }
        // block passes R-1 .. 1: the gradient w.r.t. a pass's input is the gradient w.r.t. the previous pass's output AND is
        // added to the gradient w.r.t. a1 (x_i = out_{i-1} + a1)                                    refinement_module.py:45-62
        float *cur = dA, *other = dB;
        bool acc_init = true;
        for (int i = repeats - 1; i >= 1; --i) {
            Ctx ci = pass_ctx(t, t.pass_ws(i));
            RUN(run_block_bwd(t, ci, t.xin(i), cur, other, dA1, acc_init));
            acc_init = false;
            float* tmp = cur;
            cur = other;
            other = tmp;
        }
        if (acc_init) CKN(cudaMemsetAsync(dA1, 0, sizeof(float) * (size_t)M * 256, t.st));  // R == 1: nothing reached a1 yet
        if (cur != dB) {  // leave the gradient w.r.t. the CAF output in dB for phase 1
            CKN(cudaMemcpyAsync(dB, cur, sizeof(float) * (size_t)M * 256, cudaMemcpyDeviceToDevice, t.st));
        }
        // CAF reductions                                                                            layers/fusion.py:252-274
        ca.dout = dB;
        CKN(cudaMemsetAsync(ca.csum, 0, 256 * 4 * 8, t.st));
        CKN(cudaMemsetAsync(ca.dvk, 0, sizeof(float) * (size_t)d.B * d.Tv * 256, t.st));
        CKN(cudaMemsetAsync(ca.datt, 0, sizeof(float) * (size_t)d.B * d.Tv * 256, t.st));
        int tpc = (d.T * d.B + sm_count() * 4 - 1) / (sm_count() * 4);
        if (tpc < 1) tpc = 1;
        ca.t_per_cta = tpc;
        caf_bwd_reduce_kernel<<<dim3((d.T + tpc - 1) / tpc, d.B), 256, 0, t.st>>>(ca);
        CK(cudaGetLastError());
        const int caf_slots[] = {RTFS_P_CAF_WR, RTFS_P_CAF_BR, RTFS_P_CAF_GR, RTFS_P_CAF_BER, RTFS_P_CAF_WA, RTFS_P_CAF_BA, RTFS_P_CAF_GA, RTFS_P_CAF_BEA};
        for (int s : caf_slots) NEED_GRAD(s);
        if (d_video == nullptr) return fail_msg("rtfs_avnet_backward: d_video missing");
        CafVideoBwdArgs va{video, params[RTFS_P_CAF_WR], params[RTFS_P_CAF_BR], params[RTFS_P_CAF_GR], params[RTFS_P_CAF_BER],
                           params[RTFS_P_CAF_WA], params[RTFS_P_CAF_BA], params[RTFS_P_CAF_GA], params[RTFS_P_CAF_BEA],
                           ca.dvk, ca.datt, d_video,
                           t.grad(RTFS_P_CAF_WR), t.grad(RTFS_P_CAF_BR), t.grad(RTFS_P_CAF_GR), t.grad(RTFS_P_CAF_BER),
                           t.grad(RTFS_P_CAF_WA), t.grad(RTFS_P_CAF_BA), t.grad(RTFS_P_CAF_GA), t.grad(RTFS_P_CAF_BEA), 256, d.Tv};
        const int smem = d.Tv * 256 * 4;
        if (smem > 200 * 1024) return fail_msg("CAF backward: too many video frames for the shared-memory softmax");
        static SmemCfg cfg;
        if (smem > 48 * 1024) CKN(ensure_smem(caf_video_bwd_kernel, smem, cfg));
        caf_video_bwd_kernel<<<d.B, 256, smem, t.st>>>(va);
        CK(cudaGetLastError());
        return 0;
    }
    // phase 1
    if (caf_mu == nullptr || caf_c0 == nullptr || caf_c1 == nullptr) return fail_msg("rtfs_avnet_backward: CAF BatchNorm vectors missing");
    ca.dout = dB;
    ca.mu = caf_mu;
    ca.c0 = caf_c0;
    ca.c1 = caf_c1;
    ca.da = dA;
    caf_bwd_apply_kernel<<<grid_for(totA4, 16), 256, 0, t.st>>>(ca, totA4);
    CK(cudaGetLastError());
    if (repeats > 1) {  // X_1 = CAF(...) + a1: the addend's share was accumulated by pass 1's backward already (dacc)
    }
    {
        Ctx c0 = pass_ctx(t, t.pass_ws(0));
        RUN(run_block_bwd(t, c0, a1, dA, dB, dA1, false));  // dx_0 -> dB, dA1 += dx_0
    }
    // bottleneck: a1 = W relu(gLN(a0)) + b                                                           conv_layers.py:65-129, tdavnet.py:59
    {
        const int need[] = {RTFS_P_BN_GAMMA, RTFS_P_BN_BETA, RTFS_P_BN_W, RTFS_P_BN_B, RTFS_P_ENC_W};
        for (int s : need) NEED_GRAD(s);
        if (params[RTFS_P_BN_WT] == nullptr) return fail_msg("backward: transposed bottleneck weight image missing");
        Ctx c0 = pass_ctx(t, t.pass_ws(0));
        const GlnRef nA0 = c0.gln(RTFS_ST_A0, RTFS_P_BN_GAMMA, RTFS_P_BN_BETA, d.P * 256);
        GlnActLoader<256, 1> xl{a0, nA0, (int)d.P, d.B};
        PlainLoader yl{dA1, 256, 256};
        CK((launch_wgrad<false>(xl, yl, t.grad(RTFS_P_BN_W), 256, M, 256, 256, t.st)));
        RUN(colsum<256>(t, dA1, M, t.grad(RTFS_P_BN_B)));
        PlainLoader al{dA1, 256, 256};
        StoreEpi ep{dA, 256, nullptr};  // gradient w.r.t. relu(gLN(a0))
        CK((launch_gemm<128, 256, false>(al, params[RTFS_P_BN_WT], ep, M, 256, t.st)));
        CKN(cudaMemsetAsync(t.red(31), 0, sizeof(double) * 2 * d.B, t.st));
        RUN((gln_unit_bwd<256, ACT_RELU>(t, dA, a0, nA0, nullptr, 31, t.grad(RTFS_P_BN_GAMMA), t.grad(RTFS_P_BN_BETA), nullptr, d.P * 256, dA0, 1)));
        // encoder conv: dW[c][k] = sum da0[m][c] * im2col(spec)[m][k]                                  encoder.py:161-175
        Im2colLoader xe{t.tbuf(RTFS_TP_SPEC), d.T, d.F};
        PlainLoader ye{dA0, 256, 256};
        CK((launch_wgrad<true>(xe, ye, t.grad(RTFS_P_ENC_W), 32, M, 256, 32, t.st)));
    }
    return 0;
}

int rtfs_block_train_forward(const float* const* params, const float* x, const float* addend, float* out, void* pass_ws, int B, int T, void* stream) {
    Ctx c;
    if (!make_ctx(c, params, pass_ws, B, T, 0, stream, true)) return -2;
    if (x == out) return fail_msg("rtfs_block_train_forward: out must not alias x");
    c.train = true;
    return run_block(c, x, addend, out);
}

int rtfs_block_backward(const float* const* params, float* const* grads, const float* x, const float* d_out, float* d_x,
                        void* pass_ws, void* scratch, int B, int T, void* stream) {
    TrainCtx t;
    if (!make_train_ctx(t, params, grads, nullptr, scratch, B, T, 0, 1, stream)) return -2;
    Ctx c = pass_ctx(t, reinterpret_cast<char*>(pass_ws));
    return run_block_bwd(t, c, x, d_out, d_x, nullptr, false);
}

int rtfs_dprnn_train_forward(const float* const* params, int which, const float* g_in, float* g_out, void* pass_ws, int B, int T, void* stream) {
    Ctx c;
    if (!make_ctx(c, params, pass_ws, B, T, 0, stream, true)) return -2;
    return run_dprnn_train(c, which, false, g_in, nullptr, g_out);
}

int rtfs_dprnn_backward(const float* const* params, float* const* grads, int which, const float* g_in, const float* d_out, float* d_in,
                        void* pass_ws, void* scratch, int B, int T, void* stream) {
    TrainCtx t;
    if (!make_train_ctx(t, params, grads, nullptr, scratch, B, T, 0, 1, stream)) return -2;
    Ctx c = pass_ctx(t, reinterpret_cast<char*>(pass_ws));
    return run_dprnn_bwd(t, c, which, g_in, d_out, d_in);
}

int rtfs_mhsa_train_forward(const float* const* params, const float* g_in, float* g_out, void* pass_ws, int B, int T, void* stream) {
    Ctx c;
    if (!make_ctx(c, params, pass_ws, B, T, 0, stream, true)) return -2;
    return run_mhsa(c, g_in, g_out);
}

int rtfs_mhsa_backward(const float* const* params, float* const* grads, const float* g_in, const float* d_out, float* d_in,
                       void* pass_ws, void* scratch, int B, int T, void* stream) {
    TrainCtx t;
    if (!make_train_ctx(t, params, grads, nullptr, scratch, B, T, 0, 1, stream)) return -2;
    Ctx c = pass_ctx(t, reinterpret_cast<char*>(pass_ws));
    return run_mhsa_bwd(t, c, g_in, d_out, d_in);
}

int rtfs_snr_loss(const float* est, const float* target, float* loss, float* d_est, double* sums, int B, int L, float scale, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (est == nullptr || target == nullptr || loss == nullptr || sums == nullptr || B < 1 || L < 1) return fail_msg("rtfs_snr_loss: bad arguments");
    CKN(cudaMemsetAsync(sums, 0, sizeof(double) * 5 * B, st));
    int chunks = (L + 256 * 32 - 1) / (256 * 32);
    if (chunks < 1) chunks = 1;
    snr_sums_kernel<<<dim3(chunks, B), 256, 0, st>>>(est, target, L, sums);
    CK(cudaGetLastError());
    snr_apply_kernel<<<dim3(chunks, B), 256, 0, st>>>(est, target, L, sums, loss, d_est, scale);
    CK(cudaGetLastError());
    return 0;
}

int rtfs_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, float max_norm, float grad_scale, double* gnorm_sq, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (p == nullptr || g == nullptr || m == nullptr || v == nullptr || n < 1 || step < 1) return fail_msg("rtfs_adamw_step: bad arguments");
    const unsigned grid = grid_for(n, 4);
    if (max_norm > 0.f) {
        if (gnorm_sq == nullptr) return fail_msg("rtfs_adamw_step: gnorm_sq scratch missing");
        CKN(cudaMemsetAsync(gnorm_sq, 0, sizeof(double), st));
        sumsq_kernel<<<grid, 256, 0, st>>>(g, n, gnorm_sq);
        CK(cudaGetLastError());
    }
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    adamw_kernel<<<grid, 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), max_norm > 0.f ? gnorm_sq : nullptr, max_norm, grad_scale);
    CK(cudaGetLastError());
    return 0;
}

}  // extern "C"
