// librtfs_b200.so -- C ABI + launch orchestration of the RTFS-Net forward path (see
// include/rtfs_b200.h for the contract and the reference file:line each entry point replaces).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rtfs_b200.h"
#include "attention.cuh"
#include "att_tc.cuh"
#include "att_core_tc.cuh"
#include "caf.cuh"
#include "dprnn.cuh"
#include "dprnn_fused.cuh"
#include "dwconv.cuh"
#include "dwroll.cuh"
#include "frontend.cuh"
#include "gemm.cuh"
#include "gemm_tc.cuh"
#include "gemm_tcp.cuh"
#include "input.cuh"
#include "video.cuh"

using namespace rtfs;

namespace {

thread_local std::string g_err;

// Kernel generation switches (debug / A-B measurement): RTFS_LEGACY_GEMM=1 routes every contraction
// through the mma.sync kernels of gemm.cuh instead of the tcgen05 kernels of gemm_tc.cuh.
bool env_flag(const char* name) {
    const char* v = getenv(name);
    return v != nullptr && v[0] != '\0' && v[0] != '0';
}
bool use_roll() {  // RTFS_LEGACY_DW=1: first-generation depthwise kernels (dwconv.cuh)
    static const bool v = !env_flag("RTFS_LEGACY_DW");
    return v;
}
bool use_fused_dprnn() {  // RTFS_UNFUSED_DPRNN=1: prep + GEMM + scan kernels instead of dprnn_fused.cuh
    static const bool v = !env_flag("RTFS_UNFUSED_DPRNN");
    return v;
}
// Which full-resolution 1x1 convs run the persistent warp-specialised kernel (gemm_tcp.cuh) instead of the
// one-tile-per-CTA kernel (gemm_tc.cuh): bit 0 bottleneck, 1 gate+projection, 2 residual conv, 3 mask (default 11: all but the residual conv).
// RTFS_PERSIST_MASK overrides (A/B measurement).
bool use_persistent(int bit) {
    static const int mask = [] {
        const char* v = getenv("RTFS_PERSIST_MASK");
        return v ? atoi(v) : 15;  // all four full-resolution 1x1 convs persistent
    }();
    return (mask >> bit) & 1;
}
bool use_wide() {  // RTFS_WIDE_GEMM=1: 512-thread CTAs for the residual conv (64 registers/thread: spills; measured slower since the
                   // epilogue constants are cached per column block)
    static const bool v = env_flag("RTFS_WIDE_GEMM");
    return v;
}
bool use_unfold() {  // RTFS_NO_UNFOLD=1: overlapping-view GEMMs through the generic im2col-style loader
    static const bool v = !env_flag("RTFS_NO_UNFOLD");
    return v;
}
bool use_tc() {
    static const bool v = !env_flag("RTFS_LEGACY_GEMM");
    return v;
}
thread_local long long g_launches = 0;

// Optional per-stage device timing (bench.py roofline leg): cudaEvent pairs recorded on the launch
// stream around every stage; collected (and reset) by rtfs_profile_collect.
struct Profiler {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    std::vector<int> stage;  // stage id of pair i -> events pool[2i], pool[2i+1]
    size_t used = 0;
    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
        }
        return pool[used++];
    }
};
thread_local Profiler g_prof;

struct StageTimer {
    cudaStream_t st;
    cudaEvent_t stop = nullptr;
    StageTimer(int id, cudaStream_t s) : st(s) {
        if (!g_prof.on) return;
        cudaEvent_t start = g_prof.get();
        stop = g_prof.get();
        g_prof.stage.push_back(id);
        cudaEventRecord(start, st);
    }
    ~StageTimer() {
        if (stop) cudaEventRecord(stop, st);
    }
};
#define STAGE(id) StageTimer stage_timer__(id, c.st)

struct Dims {
    int B, T, F, Tc, Fc, Tv;
    long long P, Pc;
};

Dims make_dims(int B, int T, int Tv) {
    Dims d;
    d.B = B;
    d.T = T;
    d.F = RTFS_F;
    d.Tc = (T - 2) / 2 + 1;
    d.Fc = RTFS_FC;
    d.Tv = Tv;
    d.P = (long long)T * d.F;
    d.Pc = (long long)d.Tc * d.Fc;
    return d;
}

struct Plan {
    long long off[RTFS_WS_COUNT];
    long long total;
};

// train_pass: the per-pass region of the training tape -- no full-resolution 256-channel buffers (they live in the
// global part of the tape), plus the dual-path RNN tape buffers
Plan make_plan(const Dims& d, bool train_pass = false) {
    const long long B = d.B;
    const long long A = B * d.P * 256, H = B * d.P * 64, G = B * d.Pc * 64;
    long long sz[RTFS_WS_COUNT];
    for (int i = 0; i < RTFS_WS_COUNT; ++i) sz[i] = G;
    sz[RTFS_WS_SPEC] = B * d.P * 2;
    sz[RTFS_WS_A0] = sz[RTFS_WS_A1] = sz[RTFS_WS_XA] = sz[RTFS_WS_XB] = A;
    sz[RTFS_WS_P_PRE] = sz[RTFS_WS_D0_PRE] = sz[RTFS_WS_LE0_PRE] = sz[RTFS_WS_LEC_PRE] = H;
    sz[RTFS_WS_N] = sz[RTFS_WS_HA] = sz[RTFS_WS_HB] = G + 8 * 64;
    const long long hp_f = (long long)d.Tc * (d.Fc + 7), hp_t = (long long)d.Fc * (d.Tc + 7);
    sz[RTFS_WS_HPAD] = B * (hp_f > hp_t ? hp_f : hp_t) * 64 + 8 * 64;
    sz[RTFS_WS_U] = B * d.Pc * 256 + 8 * 256;
    sz[RTFS_WS_Q] = sz[RTFS_WS_K] = B * d.Pc * 16;
    sz[RTFS_WS_Q18] = B * d.P * 18;
    sz[RTFS_WS_STATS] = (long long)RTFS_ST_COUNT * B * 2 * 2;  // doubles, counted in floats
    sz[RTFS_WS_VK] = sz[RTFS_WS_ATT] = B * (long long)(d.Tv > 0 ? d.Tv : 1) * 256;
    for (int i = RTFS_WS_TF_N; i <= RTFS_WS_TT_HP; ++i) sz[i] = 0;
    if (train_pass) {
        sz[RTFS_WS_SPEC] = sz[RTFS_WS_A0] = sz[RTFS_WS_A1] = sz[RTFS_WS_XA] = sz[RTFS_WS_XB] = sz[RTFS_WS_Q18] = 0;
        sz[RTFS_WS_VK] = sz[RTFS_WS_ATT] = 0;
        for (int path = 0; path < 2; ++path) {
            const int base = path == 0 ? RTFS_WS_TF_N : RTFS_WS_TT_N;
            sz[base + 0] = G + 8 * 64;                                       // N
            sz[base + 1] = B * d.Pc * 256 + 8 * 256;                         // U0
            for (int l = 1; l <= 3; ++l) sz[base + 1 + l] = B * d.Pc * 192 + 8 * 192;  // U1..U3
            for (int l = 0; l <= 3; ++l) sz[base + 5 + l] = G;               // C0..C3
            for (int l = 0; l <= 2; ++l) sz[base + 9 + l] = G + 8 * 64;      // H0..H2
            sz[base + 12] = B * (path == 0 ? hp_f : hp_t) * 64 + 8 * 64;     // HP
        }
    }
    Plan p;
    long long o = 0;
    for (int i = 0; i < RTFS_WS_COUNT; ++i) {
        p.off[i] = o;
        o += ((sz[i] * 4 + 255) / 256) * 256;
    }
    p.total = o;
    return p;
}

struct Ctx {
    const float* const* P;
    Dims d;
    Plan pl;
    char* ws;
    cudaStream_t st;
    bool train = false;  // training forward: the dual-path RNNs run the taped kernel chain (train_api.cuh)
    // rtfs_avnet_forward_av: the VP block + the CAF video branch run on a side stream behind the first block pass's
    // gateway/projection launch (fork) and are joined in front of the fused CAF epilogue
    mutable bool fork_armed = false, join_pending = false;
    cudaStream_t side = nullptr;
    const float* mouth = nullptr;
    float* video = nullptr;
    mutable cudaEvent_t ev_join = nullptr;
    float* buf(int i) const { return reinterpret_cast<float*>(ws + pl.off[i]); }
    double* stat(int slot) const { return reinterpret_cast<double*>(ws + pl.off[RTFS_WS_STATS]) + (long long)slot * d.B * 2; }
    GlnRef gln(int slot, int pg, int pb, long long n) const {
        GlnRef r;
        r.sums = stat(slot);
        r.gamma = P[pg];
        r.beta = P[pb];
        r.inv_n = 1.0 / (double)n;
        return r;
    }
};

int fail(const char* what, cudaError_t e) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return -1;
}
int fail_msg(const std::string& m) {
    g_err = m;
    return -2;
}

// CK: a kernel launch (counted) ; CKN: any other runtime call
#define CK(call)                                   \
    do {                                           \
        cudaError_t e__ = (call);                  \
        ++g_launches;                              \
        if (e__ != cudaSuccess) return fail(#call, e__); \
    } while (0)
#define CKN(call)                                  \
    do {                                           \
        cudaError_t e__ = (call);                  \
        if (e__ != cudaSuccess) return fail(#call, e__); \
    } while (0)
#define RUN(call)                 \
    do {                          \
        int r__ = (call);         \
        if (r__ != 0) return r__; \
    } while (0)

// ------------------------------------------------------------------ generic gLN statistics (module-level calls only)
__global__ void __launch_bounds__(256) gln_stats_kernel(const float* __restrict__ x, long long n_per_sample, double* sums) {
    __shared__ float scratch[16];
    const int b = blockIdx.y;
    const float4* p = reinterpret_cast<const float4*>(x + (long long)b * n_per_sample);
    const long long n4 = n_per_sample / 4;
    float s = 0.f, q = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(p + i);
        s += v.x + v.y + v.z + v.w;
        q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    block_stats_atomic(s, q, sums + 2 * b, scratch);
}

// ------------------------------------------------------------------ stages
int run_encoder(const Ctx& c, const float* wav, float* a0, int L) {
    const Dims& d = c.d;
    StftArgs sa{wav, c.P[RTFS_P_WINDOW], c.P[RTFS_P_COSTAB], c.P[RTFS_P_SINTAB], c.buf(RTFS_WS_SPEC), L, d.T};
    {
        STAGE(RTFS_SG_STFT);
        static const bool legacy_fe = env_flag("RTFS_LEGACY_FRONTEND");  // one frame per CTA (first generation)
        if (legacy_fe) stft_kernel<<<dim3(d.T, d.B), 288, 0, c.st>>>(sa);
        else stft16_kernel<<<dim3((d.T + FE_FR - 1) / FE_FR, d.B), 160, 0, c.st>>>(sa);
        CK(cudaGetLastError());
    }
    CKN(cudaMemsetAsync(c.stat(RTFS_ST_A0), 0, sizeof(double) * 2 * d.B, c.st));
    Im2colLoader al{c.buf(RTFS_WS_SPEC), d.T, d.F};
    STAGE(RTFS_SG_ENC_CONV);
    if (use_tc() && !env_flag("RTFS_LEGACY_ENC")) {
        Im2colLoader3x al3{al};
        StatsEpi4 ep{a0, 256, nullptr, c.stat(RTFS_ST_A0), (int)d.P, d.B, 0, 0, 0.f, 0.f, 0.f, 0.f};
        CK((launch_gemm_tcp<256, 96, 4, 1, true, 3, 0, 256>(al3, c.P[RTFS_P_ENC_WI3], ep, (int)(d.B * d.P), c.st)));
    } else {
        StatsEpi ep{a0, 256, nullptr, c.stat(RTFS_ST_A0), (int)d.P, d.B};
        CK((launch_gemm<128, 32, true>(al, c.P[RTFS_P_ENC_W], ep, (int)(d.B * d.P), 256, c.st)));
    }
    return 0;
}

int run_bottleneck(const Ctx& c, const float* a0, float* a1, bool compute_stats) {
    const Dims& d = c.d;
    if (compute_stats) {
        CKN(cudaMemsetAsync(c.stat(RTFS_ST_A0), 0, sizeof(double) * 2 * d.B, c.st));
        gln_stats_kernel<<<dim3(64, d.B), 256, 0, c.st>>>(a0, d.P * 256, c.stat(RTFS_ST_A0));
        CK(cudaGetLastError());
    }
    GlnActLoader<256, 1> al{a0, c.gln(RTFS_ST_A0, RTFS_P_BN_GAMMA, RTFS_P_BN_BETA, d.P * 256), (int)d.P, d.B};
    STAGE(RTFS_SG_BOTTLENECK);
    if (use_tc()) {
        StoreEpi4 ep{a1, 256, c.P[RTFS_P_BN_B]};
        if (use_persistent(0)) CK((launch_gemm_tcp<256, 256, 3, 4, false, 4, 2, 512>(al, c.P[RTFS_P_BN_WI], ep, (int)(d.B * d.P), c.st)));
        else CK((launch_gemm_tc<256, 256, 3, 1, 4, 256>(al, c.P[RTFS_P_BN_WI], ep, (int)(d.B * d.P), c.st)));
    } else {
        StoreEpi ep{a1, 256, c.P[RTFS_P_BN_B]};
        CK((launch_gemm<128, 256, false>(al, c.P[RTFS_P_BN_W], ep, (int)(d.B * d.P), 256, c.st)));
    }
    return 0;
}

int run_dprnn_train(const Ctx& c, int which, bool first, const float* g_in, float* g_first, float* g_out);  // train_api.cuh

// DualPathRNN.  first: g = gLN(d1_pre) + pool is formed here and written to g_first.
int run_dprnn(const Ctx& c, int which, bool first, const float* g_in, float* g_first, float* g_out) {
    const Dims& d = c.d;
    const int base = which == 0 ? RTFS_P_RF_LNG : RTFS_P_RT_LNG;
    const int basei = which == 0 ? RTFS_P_RF_WI0 : RTFS_P_RT_WI0;
    const int S = which == 0 ? d.Fc : d.Tc;
    const int n_other = which == 0 ? d.Tc : d.Fc;
    const int nseq = d.B * n_other;
    const int L = S - 7;
    if (L < 1) return fail_msg("dual-path RNN needs at least 8 steps along the scanned axis");
    if (use_tc() && use_fused_dprnn() && S <= DF_NP) {
        DfArgs a;
        memset(&a, 0, sizeof(a));
        a.g_in = g_in;
        a.d1_pre = c.buf(RTFS_WS_D1_PRE);
        a.pool = c.buf(RTFS_WS_POOL);
        a.gln = c.gln(RTFS_ST_D1, RTFS_P_D1_GAMMA, RTFS_P_D1_BETA, d.Pc * 64);
        a.g_first = g_first;
        a.ln_gamma = c.P[base + 0];
        a.ln_beta = c.P[base + 1];
        a.wimg = c.P[which == 0 ? RTFS_P_RF_FUSED : RTFS_P_RT_FUSED];
        for (int l = 0; l < 4; ++l) {
            a.wc[l] = c.P[base + 3 + 3 * l];
            a.bias[l] = c.P[base + 4 + 3 * l];
        }
        a.ct_bias = c.P[base + 15];
        a.g_out = g_out;
        a.B = d.B;
        a.Tc = d.Tc;
        a.Fc = d.Fc;
        a.time_path = which;
        a.first = first ? 1 : 0;
        a.S = S;
        a.L = L;
        a.nseq_total = nseq;
        a.n_other = n_other;
        static const bool dbg = env_flag("RTFS_DF_DEBUG");
        static const bool no_yield = env_flag("RTFS_DF_NO_YIELD");  // A/B: persistent CTAs even while the forked VP block is in flight
        a.yield_sms = (!no_yield && c.join_pending) ? 1 : 0;  // the forked VP block needs SMs (11.77 vs 11.82 ms per forward)
        const int tiles = dprnn_fused_dbg_rows(S, nseq, a.yield_sms != 0);  // timeline rows; nseq_tile is set by the launcher for the tile size it picks
        if (dbg) {
            CKN(cudaMalloc(&a.dbg, sizeof(long long) * 32 * tiles));
            CKN(cudaMemset(a.dbg, 0, sizeof(long long) * 32 * tiles));
        }
        {
            STAGE(RTFS_SG_DPRNN_FUSED);
            CK(launch_dprnn_fused(a, c.st));
        }
        if (dbg) {  // phase timeline of a few tiles (cycles between the stamps of dprnn_fused.cuh)
            std::vector<long long> h(32 * (size_t)tiles);
            CKN(cudaMemcpy(h.data(), a.dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
            cudaFree(a.dbg);
            const int picks[3] = {0, 1, tiles - 1};
            for (int t : picks) {
                fprintf(stderr, "dprnn_fused which=%d row %d/%d (start %+lld):", which, t, tiles, h[16 * t] - h[0]);
                for (int i = 1; i < 16 && h[16 * t + i] != 0; ++i) fprintf(stderr, " %lld", h[16 * t + i] - h[16 * t + i - 1]);
                fprintf(stderr, " | h-warp spans:");
                for (int i = 0; i < 4; ++i) fprintf(stderr, " %lld", h[16 * (tiles + t) + 2 * i + 1] - h[16 * (tiles + t) + 2 * i]);
                fprintf(stderr, "\n");
            }
        }
        return 0;
    }
    const int M = nseq * S;
    float* n = c.buf(RTFS_WS_N);
    float* U = c.buf(RTFS_WS_U);
    float* hA = c.buf(RTFS_WS_HA);
    float* hB = c.buf(RTFS_WS_HB);
    float* hpad = c.buf(RTFS_WS_HPAD);

    PrepArgs pa;
    pa.g_in = g_in;
    pa.d1_pre = c.buf(RTFS_WS_D1_PRE);
    pa.pool = c.buf(RTFS_WS_POOL);
    pa.gln = c.gln(RTFS_ST_D1, RTFS_P_D1_GAMMA, RTFS_P_D1_BETA, d.Pc * 64);
    pa.ln_gamma = c.P[base + 0];
    pa.ln_beta = c.P[base + 1];
    pa.g_out = g_first;
    pa.n_out = n;
    pa.B = d.B;
    pa.Tc = d.Tc;
    pa.Fc = d.Fc;
    pa.time_path = which;
    pa.first = first ? 1 : 0;
    const long long npos = d.B * d.Pc;
    {
        STAGE(RTFS_SG_DPRNN_PREP);
        dprnn_prep_kernel<<<(unsigned)((npos + 15) / 16), 256, 0, c.st>>>(pa);
        CK(cudaGetLastError());
    }
    const float* resid = first ? g_first : g_in;

    // layer 0: unfold(8) + Linear(512 -> 256) as a GEMM over the overlapping row view of n
    {
        PlainLoader al{n, 64, 512};
        {
            STAGE(RTFS_SG_DPRNN_GEMM0);
            if (use_tc()) {
                StoreEpi4 ep{U, 256, nullptr};
                if (use_unfold()) CK((launch_gemm_tc_unfold<256, 3, 1>(n, c.P[basei + 0], ep, M, c.st)));
                else CK((launch_gemm_tc<256, 512, 3, 1, 2, 256>(al, c.P[basei + 0], ep, M, c.st)));
            } else {
                StoreEpi ep{U, 256, nullptr};
                CK((launch_gemm<128, 512, false>(al, c.P[base + 2], ep, M, 256, c.st)));
            }
        }
        STAGE(RTFS_SG_DPRNN_SCAN);
        ScanArgs sa{U, 256, nullptr, c.P[base + 3], c.P[base + 4], hA, nseq, S, L, 4, S, 0, 0};
        sru_scan_kernel<<<(nseq + 3) / 4, 256, 0, c.st>>>(sa);
        CK(cudaGetLastError());
    }
    float* hin = hA;
    for (int l = 1; l <= 3; ++l) {
        const int pw = base + 2 + 3 * l;
        PlainLoader al{hin, 64, 64};
        {
            STAGE(RTFS_SG_DPRNN_GEMML);
            if (use_tc()) {
                StoreEpi4 ep{U, 192, nullptr};
                CK((launch_gemm_tc<192, 64, 2, 2, 2, 256>(al, c.P[basei + l], ep, M, c.st)));
            } else {
                StoreEpi ep{U, 192, nullptr};
                CK((launch_gemm<64, 64, false>(al, c.P[pw], ep, M, 192, c.st)));
            }
        }
        STAGE(RTFS_SG_DPRNN_SCAN);
        const bool last = l == 3;
        float* hout = last ? hpad : (hin == hA ? hB : hA);
        ScanArgs sa{U, 192, hin, c.P[pw + 1], c.P[pw + 2], hout, nseq, S, L, 3, last ? S + 7 : S, last ? 7 : 0, last ? 1 : 0};
        sru_scan_kernel<<<(nseq + 3) / 4, 256, 0, c.st>>>(sa);
        CK(cudaGetLastError());
        hin = hout;
    }
    // ConvTranspose1d(64,64,8) + bias + residual as a GEMM over the overlapping view of the padded h
    {
        PlainLoader al{hpad, 64, 512};
        STAGE(RTFS_SG_DPRNN_CONVT);
        if (use_tc()) {
            ConvTEpi4 ep{g_out, resid, c.P[base + 15], S, n_other, which, d.Tc, d.Fc};
            if (use_unfold()) CK((launch_gemm_tc_unfold<64, 4, 2>(hpad, c.P[basei + 4], ep, nseq * (S + 7), c.st)));
            else CK((launch_gemm_tc<64, 512, 4, 2, 2, 256>(al, c.P[basei + 4], ep, nseq * (S + 7), c.st)));
        } else {
            ConvTEpi ep{g_out, resid, c.P[base + 15], S, n_other, which, d.Tc, d.Fc};
            CK((launch_gemm<64, 512, false>(al, c.P[base + 14], ep, nseq * (S + 7), 64, c.st)));
        }
    }
    return 0;
}

int run_mhsa(const Ctx& c, const float* g_in, float* g_out) {
    const Dims& d = c.d;
    const int H = 4;
    RowblockArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.x = g_in;
    ra.W = c.P[RTFS_P_AT_WQKV];
    ra.bias = c.P[RTFS_P_AT_BQKV];
    ra.slope = c.P[RTFS_P_AT_SLOPE];
    ra.gamma = c.P[RTFS_P_AT_GAMMA];
    ra.beta = c.P[RTFS_P_AT_BETA];
    ra.q = c.buf(RTFS_WS_Q);
    ra.k = c.buf(RTFS_WS_K);
    ra.v = c.buf(RTFS_WS_V);
    ra.B = d.B;
    ra.Tc = d.Tc;
    ra.H = H;
    const bool att_tc = use_tc() && !env_flag("RTFS_LEGACY_ATT");  // tcgen05 conv + PReLU + LN kernels (att_tc.cuh)
    if (att_tc) {
        AttConvArgs aa;
        memset(&aa, 0, sizeof(aa));
        aa.x = g_in;
        aa.wimg = c.P[RTFS_P_AT_WQKVI];
        aa.bias = ra.bias;
        aa.slope = ra.slope;
        aa.gamma = ra.gamma;
        aa.beta = ra.beta;
        aa.q = ra.q;
        aa.k = ra.k;
        aa.v = ra.v;
        aa.B = d.B;
        aa.Tc = d.Tc;
        aa.H = H;
        aa.nframes = d.B * d.Tc;
        STAGE(RTFS_SG_ATT_QKV);
        CK((launch_att_conv_tc<96, 0>(aa, c.st)));
    } else {
        const int smem = rowblock_smem_floats<96>() * 4;
        static SmemCfg cfg;
        CKN(ensure_smem(rowblock_ln_kernel<96, 0>, smem, cfg));
        STAGE(RTFS_SG_ATT_QKV);
        rowblock_ln_kernel<96, 0><<<d.B * d.Tc, 128, smem, c.st>>>(ra);
        CK(cudaGetLastError());
    }
    // attention core: tcgen05 + tensor-map TMA (att_core_tc.cuh) up to 256 frames; RTFS_LEGACY_ATTCORE=1 (A/B switch) and longer
    // sequences run the mma.sync kernel
    static const bool legacy_core = env_flag("RTFS_LEGACY_ATTCORE");
    bool core_done = false;
    if (att_tc && !legacy_core && d.Tc <= 256) {
        STAGE(RTFS_SG_ATT_CORE);
        CK(launch_attn_core_tc(ra.q, ra.k, ra.v, c.buf(RTFS_WS_AO), d.B, H, d.Tc, c.st));
        core_done = true;
    }
    if (!core_done) {
        AttnArgs aa;
        aa.q = ra.q;
        aa.k = ra.k;
        aa.v = ra.v;
        aa.o = c.buf(RTFS_WS_AO);
        aa.Tc = d.Tc;
        aa.H = H;
        aa.tk_pad = ((d.Tc + AT_KR - 1) / AT_KR) * AT_KR;
        aa.scale = 1.f / sqrtf(4.f * 64.f);
        static const bool qt32 = env_flag("RTFS_ATT_QT32");  // 32-query tiles (first schedule of this kernel)
        const int qt = (qt32 || d.Tc <= 32) ? 32 : 64;
        const int smem = attn_smem_floats(aa.tk_pad, qt) * 4;
        if (smem > 227 * 1024) return fail_msg("attention: too many frames for the shared-memory score tile");
        static SmemCfg cfg64, cfg32;
        if (qt == 64) CKN(ensure_smem(attn_core_kernel<64>, smem, cfg64));
        else CKN(ensure_smem(attn_core_kernel<32>, smem, cfg32));
        STAGE(RTFS_SG_ATT_CORE);
        if (qt == 64) attn_core_kernel<64><<<dim3((d.Tc + 63) / 64, d.B * H), 512, smem, c.st>>>(aa);
        else attn_core_kernel<32><<<dim3((d.Tc + 31) / 32, d.B * H), 256, smem, c.st>>>(aa);
        CK(cudaGetLastError());
    }
    if (att_tc) {
        AttConvArgs ab;
        memset(&ab, 0, sizeof(ab));
        ab.x = c.buf(RTFS_WS_AO);
        ab.resid = g_in;
        ab.wimg = c.P[RTFS_P_AT_WOI];
        ab.bias = c.P[RTFS_P_AT_BO];
        ab.slope = c.P[RTFS_P_AT_SLOPEO];
        ab.gamma = c.P[RTFS_P_AT_GAMMAO];
        ab.beta = c.P[RTFS_P_AT_BETAO];
        ab.out = g_out;
        ab.B = d.B;
        ab.Tc = d.Tc;
        ab.H = H;
        ab.nframes = d.B * d.Tc;
        STAGE(RTFS_SG_ATT_PROJ);
        CK((launch_att_conv_tc<64, 1>(ab, c.st)));
    } else {
        RowblockArgs rb;
        memset(&rb, 0, sizeof(rb));
        rb.x = c.buf(RTFS_WS_AO);
        rb.resid = g_in;
        rb.W = c.P[RTFS_P_AT_WO];
        rb.bias = c.P[RTFS_P_AT_BO];
        rb.slope = c.P[RTFS_P_AT_SLOPEO];
        rb.gamma = c.P[RTFS_P_AT_GAMMAO];
        rb.beta = c.P[RTFS_P_AT_BETAO];
        rb.out = g_out;
        rb.B = d.B;
        rb.Tc = d.Tc;
        rb.H = H;
        const int smem = rowblock_smem_floats<64>() * 4;
        static SmemCfg cfg;
        CKN(ensure_smem(rowblock_ln_kernel<64, 1>, smem, cfg));
        STAGE(RTFS_SG_ATT_PROJ);
        rowblock_ln_kernel<64, 1><<<d.B * d.Tc, 128, smem, c.st>>>(rb);
        CK(cudaGetLastError());
    }
    return 0;
}

// caf_fused: apply the CAF fusion (vk/att already produced by caf_video_kernel) to the block output in the
// residual-conv epilogue; addend is then added after the fusion (refinement_module.py:50-56).
int fork_video(const Ctx& c);

int run_block(const Ctx& c, const float* x, const float* addend, float* out, bool caf_fused = false) {
    const Dims& d = c.d;
    const float* const* P = c.P;
    const int M = (int)(d.B * d.P);
    const long long nfull = d.P * 64, ncomp = d.Pc * 64;
    float *p_pre = c.buf(RTFS_WS_P_PRE), *d0_pre = c.buf(RTFS_WS_D0_PRE), *d1_pre = c.buf(RTFS_WS_D1_PRE);
    float *pool = c.buf(RTFS_WS_POOL), *g0 = c.buf(RTFS_WS_G0), *g1 = c.buf(RTFS_WS_G1), *g2 = c.buf(RTFS_WS_G2), *g3 = c.buf(RTFS_WS_G3);
    float *le0 = c.buf(RTFS_WS_LE0_PRE), *lec = c.buf(RTFS_WS_LEC_PRE), *le1 = c.buf(RTFS_WS_LE1);
    float *ge0 = c.buf(RTFS_WS_GE0), *gg0 = c.buf(RTFS_WS_GG0), *ge1 = c.buf(RTFS_WS_GE1), *gg1 = c.buf(RTFS_WS_GG1);
    float *gec = c.buf(RTFS_WS_GEC), *ggc = c.buf(RTFS_WS_GGC);
    const int tseg_full = 32, tseg_comp = 16;

    CKN(cudaMemsetAsync(c.stat(RTFS_ST_PJ), 0, sizeof(double) * 2 * d.B * (RTFS_ST_COUNT - RTFS_ST_PJ), c.st));
    // S1 gateway + projection (+ gLN statistics)                         tdanet.py:34-49,107-108
    {
        GateLoader al{x, P[RTFS_P_GW_W], P[RTFS_P_GW_B], P[RTFS_P_GW_A], 256};
        STAGE(RTFS_SG_GATE_PROJ);
        if (use_tc()) {
            StatsEpi4 ep{p_pre, 64, P[RTFS_P_PJ_B], c.stat(RTFS_ST_PJ), (int)d.P, d.B};
            if (use_persistent(1)) CK((launch_gemm_tcp<64, 256, 4, 1, true, 4, 2, 512>(al, P[RTFS_P_PJ_WI], ep, M, c.st)));
                        else CK((launch_gemm_tc<64, 256, 4, 2, 2, 256>(al, P[RTFS_P_PJ_WI], ep, M, c.st)));
        } else {
            StatsEpi ep{p_pre, 64, P[RTFS_P_PJ_B], c.stat(RTFS_ST_PJ), (int)d.P, d.B};
            CK((launch_gemm<64, 256, false>(al, P[RTFS_P_PJ_W], ep, M, 64, c.st)));
        }
    }
    if (c.fork_armed) RUN(fork_video(c));
    const bool roll = use_roll();
    if (roll) {
        // S2 PReLU(gLN(p)) -> dw4x4 s1 -> d0_pre                              tdanet.py:61-68,113
        {
            XrGln<2> xf{p_pre, d.T, d.F, c.gln(RTFS_ST_PJ, RTFS_P_PJ_GAMMA, RTFS_P_PJ_BETA, nfull), P[RTFS_P_PJ_A]};
            DrArgs<1> a{};
            a.Ti = d.T; a.Fi = d.F;
            a.w[0] = P[RTFS_P_D0_W]; a.bias[0] = P[RTFS_P_D0_B]; a.out[0] = d0_pre; a.sums[0] = c.stat(RTFS_ST_D0);
            STAGE(RTFS_SG_DW_S1);
            CK((launch_dwroll<XrGln<2>, 1, false, 288>(xf, a, d.B, c.st)));
        }
        // S3 one pass over gLN(d0_pre): local conv of fusion_layers.0 (le0), dw4x4 s2 (d1_pre), adaptive pool
        //                                                            tdanet.py:69-76,114-118 ; layers/fusion.py:56
        {
            XrGln<0> xf{d0_pre, d.T, d.F, c.gln(RTFS_ST_D0, RTFS_P_D0_GAMMA, RTFS_P_D0_BETA, nfull), nullptr};
            DrArgs<1> a{};
            a.Ti = d.T; a.Fi = d.F;
            a.w[0] = P[RTFS_P_F0_LW]; a.bias[0] = nullptr; a.out[0] = le0; a.sums[0] = c.stat(RTFS_ST_F0L);
            a.w2 = P[RTFS_P_D1_W]; a.bias2 = P[RTFS_P_D1_B]; a.out2 = d1_pre; a.sums2 = c.stat(RTFS_ST_D1);
            a.pool = pool; a.To2 = d.Tc; a.Fo2 = d.Fc;
            STAGE(RTFS_SG_DW_S2_POOL);
            CK((launch_dwroll<XrGln<0>, 1, true, 288>(xf, a, d.B, c.st)));
        }
    } else {
        // S2 PReLU(gLN(p)) -> dw4x4 s1 -> d0_pre                              tdanet.py:61-68,113
        {
            XfGln<2> xf{p_pre, d.T, d.F, c.gln(RTFS_ST_PJ, RTFS_P_PJ_GAMMA, RTFS_P_PJ_BETA, nfull), P[RTFS_P_PJ_A]};
            DwArgs<1> a{d.T, d.F, d.T, d.F, tseg_full, {P[RTFS_P_D0_W]}, {P[RTFS_P_D0_B]}, {d0_pre}, {c.stat(RTFS_ST_D0)}, nullptr};
            STAGE(RTFS_SG_DW_S1);
            CK((launch_dw<1, 4, 1, false>(xf, a, d.B, c.st)));
        }
        // S3 gLN(d0_pre) -> dw4x4 s2 -> d1_pre ; adaptive_avg_pool2d(d0) -> pool    tdanet.py:69-76,114-118
        {
            XfGln<0> xf{d0_pre, d.T, d.F, c.gln(RTFS_ST_D0, RTFS_P_D0_GAMMA, RTFS_P_D0_BETA, nfull), nullptr};
            DwArgs<1> a{d.T, d.F, d.Tc, d.Fc, tseg_comp, {P[RTFS_P_D1_W]}, {P[RTFS_P_D1_B]}, {d1_pre}, {c.stat(RTFS_ST_D1)}, pool};
            STAGE(RTFS_SG_DW_S2_POOL);
            CK((launch_dw<2, 2, 1, true>(xf, a, d.B, c.st)));
        }
    }
    // S4-S9 global attention stack                                         tdanet.py:121
    if (c.train) {
        RUN(run_dprnn_train(c, 0, true, nullptr, g0, g1));
        RUN(run_dprnn_train(c, 1, false, g1, nullptr, g2));
    } else {
        RUN(run_dprnn(c, 0, true, nullptr, g0, g1));
        RUN(run_dprnn(c, 1, false, g1, nullptr, g2));
    }
    RUN(run_mhsa(c, g2, g3));
    if (roll) {
        // S10-S12 TF-AR units                                                  tdanet.py:124-129, layers/fusion.py:54-69
        {
            // the four "global" convs (embedding + gate of fusion_layers.0 and .1) share their input g3
            XrPlain xf{g3, d.Tc, d.Fc};
            DrArgs<4> a{};
            a.Ti = d.Tc; a.Fi = d.Fc;
            a.w[0] = P[RTFS_P_F0_EW]; a.out[0] = ge0; a.sums[0] = c.stat(RTFS_ST_F0E);
            a.w[1] = P[RTFS_P_F0_GW]; a.out[1] = gg0; a.sums[1] = c.stat(RTFS_ST_F0G);
            a.w[2] = P[RTFS_P_F1_EW]; a.out[2] = ge1; a.sums[2] = c.stat(RTFS_ST_F1E);
            a.w[3] = P[RTFS_P_F1_GW]; a.out[3] = gg1; a.sums[3] = c.stat(RTFS_ST_F1G);
            STAGE(RTFS_SG_TFAR_GLOBAL);
            CK((launch_dwroll<XrPlain, 4, false, 128>(xf, a, d.B, c.st)));
        }
        {
            XrGln<0> xf{d1_pre, d.Tc, d.Fc, c.gln(RTFS_ST_D1, RTFS_P_D1_GAMMA, RTFS_P_D1_BETA, ncomp), nullptr};
            DrArgs<1> a{};
            a.Ti = d.Tc; a.Fi = d.Fc;
            a.w[0] = P[RTFS_P_F1_LW]; a.out[0] = le1; a.sums[0] = c.stat(RTFS_ST_F1L);
            STAGE(RTFS_SG_TFAR_GLOBAL);
            CK((launch_dwroll<XrGln<0>, 1, false, 128>(xf, a, d.B, c.st)));
        }
        {
            // f1 = TFAR_fus1(d1, g) formed on the fly -> the two global convs of concat_layers.0
            XrTfarP xf{le1, gg1, ge1, d.Tc, d.Fc, d.Tc, d.Fc,
                      c.gln(RTFS_ST_F1L, RTFS_P_F1_LG, RTFS_P_F1_LB, ncomp), c.gln(RTFS_ST_F1G, RTFS_P_F1_GG, RTFS_P_F1_GB, ncomp),
                      c.gln(RTFS_ST_F1E, RTFS_P_F1_EG, RTFS_P_F1_EB, ncomp)};
            DrArgs<2> a{};
            a.Ti = d.Tc; a.Fi = d.Fc;
            a.w[0] = P[RTFS_P_C0_EW]; a.out[0] = gec; a.sums[0] = c.stat(RTFS_ST_C0E);
            a.w[1] = P[RTFS_P_C0_GW]; a.out[1] = ggc; a.sums[1] = c.stat(RTFS_ST_C0G);
            STAGE(RTFS_SG_TFAR_CAT_GLOBAL);
            CK((launch_dwroll<XrTfarP, 2, false, 128, 4>(xf, a, d.B, c.st)));
        }
        {
            // f0 = TFAR_fus0(d0, g) formed on the fly -> local conv of concat_layers.0
            DrArgs<1> a{};
            a.Ti = d.T; a.Fi = d.F;
            a.w[0] = P[RTFS_P_C0_LW]; a.out[0] = lec; a.sums[0] = c.stat(RTFS_ST_C0L);
            STAGE(RTFS_SG_TFAR_CAT_LOCAL);
            static const bool scalar = env_flag("RTFS_SCALAR_TFAR");  // first-generation kernel (sigmoid per window column)
            if (scalar) {
                XrTfar xf{le0, gg0, ge0, d.T, d.F, d.Tc, d.Fc,
                          c.gln(RTFS_ST_F0L, RTFS_P_F0_LG, RTFS_P_F0_LB, nfull), c.gln(RTFS_ST_F0G, RTFS_P_F0_GG, RTFS_P_F0_GB, ncomp),
                          c.gln(RTFS_ST_F0E, RTFS_P_F0_EG, RTFS_P_F0_EB, ncomp)};
                CK((launch_dwroll_scalar<XrTfar, 1, false, 288>(xf, a, d.B, c.st)));
            } else {
                XrTfarP xf{le0, gg0, ge0, d.T, d.F, d.Tc, d.Fc,
                           c.gln(RTFS_ST_F0L, RTFS_P_F0_LG, RTFS_P_F0_LB, nfull), c.gln(RTFS_ST_F0G, RTFS_P_F0_GG, RTFS_P_F0_GB, ncomp),
                           c.gln(RTFS_ST_F0E, RTFS_P_F0_EG, RTFS_P_F0_EB, ncomp)};
                CK((launch_dwroll<XrTfarP, 1, false, 288>(xf, a, d.B, c.st)));
            }
        }
    } else {
        // S10-S12 TF-AR units                                                  tdanet.py:124-129, layers/fusion.py:54-69
        {
            XfPlain xf{g3, d.Tc, d.Fc};
            STAGE(RTFS_SG_TFAR_GLOBAL);
            DwArgs<2> a0{d.Tc, d.Fc, d.Tc, d.Fc, tseg_comp, {P[RTFS_P_F0_EW], P[RTFS_P_F0_GW]}, {nullptr, nullptr}, {ge0, gg0}, {c.stat(RTFS_ST_F0E), c.stat(RTFS_ST_F0G)}, nullptr};
            CK((launch_dw<1, 2, 2, false>(xf, a0, d.B, c.st)));
            DwArgs<2> a1{d.Tc, d.Fc, d.Tc, d.Fc, tseg_comp, {P[RTFS_P_F1_EW], P[RTFS_P_F1_GW]}, {nullptr, nullptr}, {ge1, gg1}, {c.stat(RTFS_ST_F1E), c.stat(RTFS_ST_F1G)}, nullptr};
            CK((launch_dw<1, 2, 2, false>(xf, a1, d.B, c.st)));
        }
        {
            XfGln<0> xf{d1_pre, d.Tc, d.Fc, c.gln(RTFS_ST_D1, RTFS_P_D1_GAMMA, RTFS_P_D1_BETA, ncomp), nullptr};
            DwArgs<1> a{d.Tc, d.Fc, d.Tc, d.Fc, tseg_comp, {P[RTFS_P_F1_LW]}, {nullptr}, {le1}, {c.stat(RTFS_ST_F1L)}, nullptr};
            STAGE(RTFS_SG_TFAR_GLOBAL);
            CK((launch_dw<1, 4, 1, false>(xf, a, d.B, c.st)));
        }
        {
            XfGln<0> xf{d0_pre, d.T, d.F, c.gln(RTFS_ST_D0, RTFS_P_D0_GAMMA, RTFS_P_D0_BETA, nfull), nullptr};
            DwArgs<1> a{d.T, d.F, d.T, d.F, tseg_full, {P[RTFS_P_F0_LW]}, {nullptr}, {le0}, {c.stat(RTFS_ST_F0L)}, nullptr};
            STAGE(RTFS_SG_TFAR_LE0);
            CK((launch_dw<1, 4, 1, false>(xf, a, d.B, c.st)));
        }
        {
            // f1 = TFAR_fus1(d1, g) formed on the fly -> the two global convs of concat_layers.0
            XfTfar xf{le1, gg1, ge1, d.Tc, d.Fc, d.Tc, d.Fc,
                      c.gln(RTFS_ST_F1L, RTFS_P_F1_LG, RTFS_P_F1_LB, ncomp), c.gln(RTFS_ST_F1G, RTFS_P_F1_GG, RTFS_P_F1_GB, ncomp),
                      c.gln(RTFS_ST_F1E, RTFS_P_F1_EG, RTFS_P_F1_EB, ncomp)};
            DwArgs<2> a{d.Tc, d.Fc, d.Tc, d.Fc, tseg_comp, {P[RTFS_P_C0_EW], P[RTFS_P_C0_GW]}, {nullptr, nullptr}, {gec, ggc}, {c.stat(RTFS_ST_C0E), c.stat(RTFS_ST_C0G)}, nullptr};
            STAGE(RTFS_SG_TFAR_CAT_GLOBAL);
            CK((launch_dw<1, 2, 2, false>(xf, a, d.B, c.st)));
        }
        {
            // f0 = TFAR_fus0(d0, g) formed on the fly -> local conv of concat_layers.0
            XfTfar xf{le0, gg0, ge0, d.T, d.F, d.Tc, d.Fc,
                      c.gln(RTFS_ST_F0L, RTFS_P_F0_LG, RTFS_P_F0_LB, nfull), c.gln(RTFS_ST_F0G, RTFS_P_F0_GG, RTFS_P_F0_GB, ncomp),
                      c.gln(RTFS_ST_F0E, RTFS_P_F0_EG, RTFS_P_F0_EB, ncomp)};
            DwArgs<1> a{d.T, d.F, d.T, d.F, tseg_full, {P[RTFS_P_C0_LW]}, {nullptr}, {lec}, {c.stat(RTFS_ST_C0L)}, nullptr};
            STAGE(RTFS_SG_TFAR_CAT_LOCAL);
            CK((launch_dw<1, 4, 1, false>(xf, a, d.B, c.st)));
        }
    }
    // S13 e = TFAR_cat0(f0,f1) + d0 on the fly -> residual_conv + gateway(x) [+ addend]     tdanet.py:127-131
    {
        TfarLoader al;
        al.lec = lec;
        al.d0 = d0_pre;
        al.ggc = ggc;
        al.gec = gec;
        al.n_l = c.gln(RTFS_ST_C0L, RTFS_P_C0_LG, RTFS_P_C0_LB, nfull);
        al.n_d = c.gln(RTFS_ST_D0, RTFS_P_D0_GAMMA, RTFS_P_D0_BETA, nfull);
        al.n_g = c.gln(RTFS_ST_C0G, RTFS_P_C0_GG, RTFS_P_C0_GB, ncomp);
        al.n_e = c.gln(RTFS_ST_C0E, RTFS_P_C0_EG, RTFS_P_C0_EB, ncomp);
        al.T = d.T;
        al.F = d.F;
        al.Tc = d.Tc;
        al.Fc = d.Fc;
        al.B = d.B;
        if (c.join_pending) {  // the fused CAF epilogue reads the side stream's key / attention tables
            CKN(cudaStreamWaitEvent(c.st, c.ev_join, 0));
            CKN(cudaEventDestroy(c.ev_join));
            c.join_pending = false;
        }
        STAGE(caf_fused ? RTFS_SG_RESID_OUT_CAF : RTFS_SG_RESID_OUT);
        if (caf_fused) {
            if (addend != nullptr && addend != x) return fail_msg("block (fused CAF pass): the addend must be the block input itself");
            ResidOutCafEpi4 ep{out, P[RTFS_P_RC_B], x, P[RTFS_P_GW_W], P[RTFS_P_GW_B], P[RTFS_P_GW_A], addend,
                               c.buf(RTFS_WS_VK), c.buf(RTFS_WS_ATT), P[RTFS_P_CAF_SK], P[RTFS_P_CAF_TK], P[RTFS_P_CAF_SV], P[RTFS_P_CAF_TV],
                               d.T, d.F, d.Tv, 0.f};
            CK((launch_gemm_tc<256, 64, 2, 2, 2, 256>(al, P[RTFS_P_RC_WI], ep, M, c.st)));  // (the persistent kernel's 96-register cap spills this epilogue: measured 1.18 vs 0.76 ms)
        } else if (use_tc()) {
            ResidOutEpi4 ep{out, P[RTFS_P_RC_B], x, P[RTFS_P_GW_W], P[RTFS_P_GW_B], P[RTFS_P_GW_A], addend, 0.f};
            if (use_persistent(2)) CK((launch_gemm_tcp<256, 64, 4, 1, true, 2, 0, 256>(al, P[RTFS_P_RC_WI], ep, M, c.st)));
            else if (use_wide()) CK((launch_gemm_tc<256, 64, 2, 2, 2, 512>(al, P[RTFS_P_RC_WI], ep, M, c.st)));
            else CK((launch_gemm_tc<256, 64, 2, 2, 2, 256>(al, P[RTFS_P_RC_WI], ep, M, c.st)));
        } else {
            ResidOutEpi ep{out, P[RTFS_P_RC_B], x, P[RTFS_P_GW_W], P[RTFS_P_GW_B], P[RTFS_P_GW_A], addend, 0.f};
            CK((launch_gemm<128, 64, false>(al, P[RTFS_P_RC_W], ep, M, 256, c.st)));
        }
    }
    return 0;
}

int run_caf_video(const Ctx& c, const float* video);

int run_caf(const Ctx& c, const float* audio, const float* video, const float* addend, float* out) {
    const Dims& d = c.d;
    const float* const* P = c.P;
    RUN(run_caf_video(c, video));
    CafApplyArgs aa{audio, addend, c.buf(RTFS_WS_VK), c.buf(RTFS_WS_ATT), P[RTFS_P_CAF_SK], P[RTFS_P_CAF_TK],
                    P[RTFS_P_CAF_SV], P[RTFS_P_CAF_TV], out, d.T, d.F, 256, d.Tv, d.B * d.P * 64};
    long long blocks = (aa.total4 + 255) / 256;
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    STAGE(RTFS_SG_CAF_APPLY);
    caf_apply_kernel<<<(unsigned)blocks, 256, 0, c.st>>>(aa);
    CK(cudaGetLastError());
    return 0;
}

int run_caf_video(const Ctx& c, const float* video) {
    const Dims& d = c.d;
    const float* const* P = c.P;
    CafVideoArgs va{video, P[RTFS_P_CAF_WR], P[RTFS_P_CAF_BR], P[RTFS_P_CAF_GR], P[RTFS_P_CAF_BER],
                    P[RTFS_P_CAF_WA], P[RTFS_P_CAF_BA], P[RTFS_P_CAF_GA], P[RTFS_P_CAF_BEA],
                    c.buf(RTFS_WS_VK), c.buf(RTFS_WS_ATT), 256, d.Tv};
    const int smem = d.Tv * 256 * 4;
    if (smem > 200 * 1024) return fail_msg("CAF: too many video frames for the shared-memory softmax");
    static SmemCfg cfg;
    if (smem > 48 * 1024) CKN(ensure_smem(caf_video_kernel, smem, cfg));
    STAGE(RTFS_SG_CAF_VIDEO);
    caf_video_kernel<<<d.B, 256, smem, c.st>>>(va);
    CK(cudaGetLastError());
    return 0;
}

// VP block of the lip embedding as one kernel (video.cuh): x (B,512,Tv) -> out (B,512,Tv)
int run_video(const float* const* params, const float* x, float* out, int B, int Tv, cudaStream_t st) {
    if (params == nullptr || params[RTFS_P_VIDEO_PACK] == nullptr) return fail_msg("rtfs_video_forward: packed video parameters missing");
    if (B < 1 || Tv < 8 || Tv > 100) return fail_msg("rtfs_video_forward: 8 <= Tv <= 100 frames");
    VideoArgs a;
    a.x = x;
    a.w = params[RTFS_P_VIDEO_PACK];
    a.out = out;
    a.off = vp_offsets();
    a.Tv = Tv;
    int len = Tv, start = 0;
    for (int i = 0; i < VP_DEPTH; ++i) {
        a.len[i] = len;
        a.start[i] = start;
        start += len;
        len = (len - 1) / 2 + 1;  // k = 3, stride 2, padding 1
    }
    a.sumlen = start;
    const int smem = vp_smem_floats(Tv, a.sumlen) * 4;
    if (smem > 227 * 1024) return fail_msg("rtfs_video_forward: shared-memory budget exceeded");
    static SmemCfg cfg;
    CKN(ensure_smem(video_block_kernel, smem, cfg));
    {
        struct { cudaStream_t st; } c{st};
        STAGE(RTFS_SG_VIDEO);
        video_block_kernel<<<B, 256, smem, st>>>(a);
        CK(cudaGetLastError());
    }
    return 0;
}

// Fork: the VP block and the CAF video branch depend on the lip embedding only.  One CTA per utterance cannot fill the
// device, so they run on a (high-priority) side stream next to the full-resolution depthwise / dual-path RNN / attention
// kernels of the first block pass instead of in front of them; ordinary event fork / join, so the forward stays capturable.
int fork_video(const Ctx& c) {
    c.fork_armed = false;
    cudaEvent_t ev_fork = nullptr;
    CKN(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    CKN(cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
    CKN(cudaEventRecord(ev_fork, c.st));
    CKN(cudaStreamWaitEvent(c.side, ev_fork, 0));
    CKN(cudaEventDestroy(ev_fork));
    RUN(run_video(c.P, c.mouth, c.video, c.d.B, c.d.Tv, c.side));
    Ctx cs = c;
    cs.st = c.side;
    RUN(run_caf_video(cs, c.video));
    CKN(cudaEventRecord(c.ev_join, c.side));
    c.join_pending = true;
    return 0;
}

int run_mask(const Ctx& c, const float* refined, const float* a0, float* z, float* m_out = nullptr) {
    const Dims& d = c.d;
    PreluLoader al{refined, c.P[RTFS_P_MK_A], 256};
    STAGE(RTFS_SG_MASK);
    if (m_out != nullptr) {  // training: the mask itself goes to the tape
        MaskEpi4T<true> ep{z, c.P[RTFS_P_MK_B], a0, m_out};
        CK((launch_gemm_tc<256, 256, 3, 1, 4, 256>(al, c.P[RTFS_P_MK_WI], ep, (int)(d.B * d.P), c.st)));
    } else if (use_tc()) {
        MaskEpi4 ep{z, c.P[RTFS_P_MK_B], a0, nullptr};
        if (use_persistent(3)) CK((launch_gemm_tcp<256, 256, 3, 4, false, 2, 2, 512>(al, c.P[RTFS_P_MK_WI], ep, (int)(d.B * d.P), c.st)));  // prefetch depth 2: see run_mask_dec
        else CK((launch_gemm_tc<256, 256, 3, 1, 4, 256>(al, c.P[RTFS_P_MK_WI], ep, (int)(d.B * d.P), c.st)));
    } else {
        MaskEpi ep{z, c.P[RTFS_P_MK_B], a0};
        CK((launch_gemm<128, 256, false>(al, c.P[RTFS_P_MK_W], ep, (int)(d.B * d.P), 256, c.st)));
    }
    return 0;
}

// S^3 mask with the decoder's 256 -> 18 contraction in its epilogue: refined, a0 -> Q18 (workspace); z is never written
int run_mask_dec(const Ctx& c, const float* refined, const float* a0) {
    const Dims& d = c.d;
    PreluLoader al{refined, c.P[RTFS_P_MK_A], 256};
    MaskDecEpi4 ep{c.buf(RTFS_WS_Q18), c.P[RTFS_P_MK_B], a0, c.P[RTFS_P_DEC_WT]};
    STAGE(RTFS_SG_MASK_DEC);
    CK((launch_gemm_tcp<256, 256, 3, 3, false, 2, 2, 512>(al, c.P[RTFS_P_MK_WI], ep, (int)(d.B * d.P), c.st)));  // prefetch depth 2: at 4 the 56-register producers spill
#ifdef RTFS_TCP_TRACE
    {
        long long h[64];
        CKN(cudaStreamSynchronize(c.st));
        CKN(cudaMemcpyFromSymbol(h, g_tcp_trace, sizeof(h)));
        fprintf(stderr, "mask_dec epilogue warp 0, third tile: wait tmem_full %lld |", h[1] - h[0]);
        for (int cb = 0; cb < 4; ++cb)
            fprintf(stderr, " block %d: issue loads %lld, tmem + stage %lld, inputs + transform %lld, reduce %lld |", cb, h[2 + 5 * cb] - (cb ? h[5 * cb] : h[1]),
                    h[3 + 5 * cb] - h[2 + 5 * cb], h[4 + 5 * cb] - h[3 + 5 * cb], h[5 + 5 * cb] - h[4 + 5 * cb]);
        fprintf(stderr, " tile_done %lld, tile total %lld\n", h[30] - h[20], h[30] - h[0]);
    }
#endif
    return 0;
}

// z == nullptr: the partial products Q18 are already in the workspace (run_mask_dec)
int run_decoder(const Ctx& c, const float* z, float* wav_out, int L) {
    const Dims& d = c.d;
    if (z != nullptr) {
        PlainLoader al{z, 256, 256};
        StoreEpi ep{c.buf(RTFS_WS_Q18), 18, nullptr};
        STAGE(RTFS_SG_DEC_GEMM);
        CK((launch_gemm<32, 256, true>(al, c.P[RTFS_P_DEC_W], ep, (int)(d.B * d.P), 18, c.st)));
    }
    IstftArgs ia{c.buf(RTFS_WS_Q18), c.P[RTFS_P_WINDOW], c.P[RTFS_P_COSTAB], c.P[RTFS_P_SINTAB], wav_out, L, d.T};
    STAGE(RTFS_SG_DEC_ISTFT);
    static const bool legacy_fe = env_flag("RTFS_LEGACY_FRONTEND");
    if (legacy_fe) dec_istft_kernel<<<dim3((L + 127) / 128, d.B), 256, 0, c.st>>>(ia);
    else dec_istft16_kernel<<<dim3(((L + 127) / 128 + FE_FR - 1) / FE_FR, d.B), 256, 0, c.st>>>(ia);
    CK(cudaGetLastError());
    return 0;
}

bool make_ctx(Ctx& c, const float* const* params, void* ws, int B, int T, int Tv, void* stream, bool train_pass = false) {
    if (params == nullptr || B < 1 || T < 16) {
        g_err = "bad arguments (params null, B < 1 or fewer than 16 frames)";
        return false;
    }
    c.P = params;
    c.d = make_dims(B, T, Tv);
    if ((long long)B * c.d.P * 256 >= (1ll << 31)) {
        g_err = "batch too large for 32-bit row indexing (B*T*F*256 must be < 2^31 elements per call)";
        return false;
    }
    c.pl = make_plan(c.d, train_pass);
    c.ws = reinterpret_cast<char*>(ws);
    c.st = reinterpret_cast<cudaStream_t>(stream);
    return true;
}

}  // namespace

extern "C" {

int rtfs_abi_version(void) { return RTFS_ABI_VERSION; }
const char* rtfs_last_error(void) { return g_err.c_str(); }
long long rtfs_last_launch_count(void) { return g_launches; }

void rtfs_profile_enable(int on) {
    g_prof.on = on != 0;
    g_prof.used = 0;
    g_prof.stage.clear();
}

int rtfs_profile_collect(float* ms, int* count) {
    for (int i = 0; i < RTFS_SG_COUNT; ++i) {
        ms[i] = 0.f;
        count[i] = 0;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail("cudaDeviceSynchronize", e);
    for (size_t i = 0; i < g_prof.stage.size(); ++i) {
        float t = 0.f;
        e = cudaEventElapsedTime(&t, g_prof.pool[2 * i], g_prof.pool[2 * i + 1]);
        if (e != cudaSuccess) return fail("cudaEventElapsedTime", e);
        ms[g_prof.stage[i]] += t;
        count[g_prof.stage[i]] += 1;
    }
    g_prof.used = 0;
    g_prof.stage.clear();
    return 0;
}

long long rtfs_ws_plan(int B, int L, int Tv, long long* offsets) {
    const Dims d = make_dims(B, L / 128 + 1, Tv);
    const Plan p = make_plan(d);
    if (offsets)
        for (int i = 0; i < RTFS_WS_COUNT; ++i) offsets[i] = p.off[i];
    return p.total;
}

int rtfs_encoder_forward(const float* const* params, const float* wav, float* a0, void* ws, int B, int L, void* stream) {
    Ctx c;
    if (!make_ctx(c, params, ws, B, L / 128 + 1, 0, stream)) return -2;
    return run_encoder(c, wav, a0, L);
}

int rtfs_bottleneck_forward(const float* const* params, const float* a0, float* a1, void* ws, int B, int T, void* stream) {
    Ctx c;
    if (!make_ctx(c, params, ws, B, T, 0, stream)) return -2;
    return run_bottleneck(c, a0, a1, true);
}

int rtfs_block_forward(const float* const* params, const float* x, const float* addend, float* out, void* ws, int B, int T, void* stream) {
    Ctx c;
    if (!make_ctx(c, params, ws, B, T, 0, stream)) return -2;
    if (x == out) return fail_msg("rtfs_block_forward: out must not alias x");
    return run_block(c, x, addend, out);
}

int rtfs_dprnn_forward(const float* const* params, int which, const float* g_in, float* g_out, void* ws, int B, int T, void* stream) {
    Ctx c;
    if (!make_ctx(c, params, ws, B, T, 0, stream)) return -2;
    return run_dprnn(c, which, false, g_in, nullptr, g_out);
}

int rtfs_mhsa_forward(const float* const* params, const float* g_in, float* g_out, void* ws, int B, int T, void* stream) {
    Ctx c;
    if (!make_ctx(c, params, ws, B, T, 0, stream)) return -2;
    return run_mhsa(c, g_in, g_out);
}

int rtfs_caf_forward(const float* const* params, const float* audio, const float* video, const float* addend, float* out, void* ws, int B, int T, int Tv, void* stream) {
    Ctx c;
    if (!make_ctx(c, params, ws, B, T, Tv, stream)) return -2;
    if (Tv < 1) return fail_msg("rtfs_caf_forward: Tv < 1");
    return run_caf(c, audio, video, addend, out);
}

int rtfs_mask_forward(const float* const* params, const float* refined, const float* a0, float* z, int B, int T, void* stream) {
    Ctx c;
    if (!make_ctx(c, params, nullptr, B, T, 0, stream)) return -2;
    return run_mask(c, refined, a0, z);
}

int rtfs_decoder_forward(const float* const* params, const float* z, float* wav_out, void* ws, int B, int L, void* stream) {
    Ctx c;
    if (!make_ctx(c, params, ws, B, L / 128 + 1, 0, stream)) return -2;
    return run_decoder(c, z, wav_out, L);
}

int rtfs_video_pack_plan(int* offsets, int* n_fields) {
    const VpOffsets o = vp_offsets();
    if (offsets)
        for (int i = 0; i < VP_COUNT; ++i) offsets[i] = o.o[i];
    if (n_fields) *n_fields = VP_COUNT;
    return o.total;
}

int rtfs_mouth_preprocess(const unsigned char* roi, float* out, const int* off_y, const int* off_x, const int* flip, int B, int T, int H, int W, int crop,
                          float mean, float std, void* stream) {
    if (roi == nullptr || out == nullptr) return fail_msg("rtfs_mouth_preprocess: null buffer");
    if (B < 1 || T < 1 || crop < 4 || (crop & 3) || crop > H || crop > W) return fail_msg("rtfs_mouth_preprocess: crop must be a multiple of 4 inside the ROI");
    if ((off_y == nullptr) != (off_x == nullptr)) return fail_msg("rtfs_mouth_preprocess: off_y and off_x go together");
    if (!(std > 0.f)) return fail_msg("rtfs_mouth_preprocess: std must be positive");
    MouthPrepArgs a{roi, out, off_y, off_x, flip, B, T, H, W, crop, mean, std};
    const long long total = (long long)B * T * crop * (crop / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    mouth_preprocess_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    CK(cudaGetLastError());
    return 0;
}

int rtfs_wav_normalize(const float* mix, const float* src, float* mix_out, float* src_out, int B, int L, int n_src, float eps, void* stream) {
    if (mix == nullptr || mix_out == nullptr) return fail_msg("rtfs_wav_normalize: null mixture buffer");
    if (n_src > 0 && (src == nullptr || src_out == nullptr)) return fail_msg("rtfs_wav_normalize: null source buffer");
    if (B < 1 || L < 2 || n_src < 0) return fail_msg("rtfs_wav_normalize: B >= 1, L >= 2, n_src >= 0");
    WavNormArgs a{mix, src, mix_out, src_out, B, L, n_src, eps};
    wav_normalize_kernel<<<dim3(1 + n_src, B), 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    CK(cudaGetLastError());
    return 0;
}

int rtfs_video_forward(const float* const* params, const float* x, float* out, int B, int Tv, void* stream) {
    return run_video(params, x, out, B, Tv, reinterpret_cast<cudaStream_t>(stream));
}

static int avnet_forward_impl(const float* const* params, const float* wav, const float* mouth, float* video, float* out, void* ws, int B, int L,
                              int Tv, int repeats, void* stream, void* side_stream) {
    Ctx c;
    if (!make_ctx(c, params, ws, B, L / 128 + 1, Tv, stream)) return -2;
    if (repeats < 1 || Tv < 1) return fail_msg("rtfs_avnet_forward: repeats < 1 or Tv < 1");
    g_launches = 0;
    float *a0 = c.buf(RTFS_WS_A0), *a1 = c.buf(RTFS_WS_A1), *xa = c.buf(RTFS_WS_XA), *xb = c.buf(RTFS_WS_XB);
    const bool fused_caf = use_tc() && !env_flag("RTFS_UNFUSED_CAF") && c.d.F >= 32;  // the fused epilogue assumes <= 2 frames per 32 rows
    bool forked = false;
    if (mouth != nullptr) {
        // per-stage timing (rtfs_profile_enable) keeps everything on one stream so that stage times add up
        if (side_stream != nullptr && side_stream != stream && fused_caf && !g_prof.on) {
            c.fork_armed = forked = true;
            c.side = reinterpret_cast<cudaStream_t>(side_stream);
            c.mouth = mouth;
            c.video = video;
        } else {
            RUN(run_video(params, mouth, video, B, Tv, c.st));
        }
    }
    RUN(run_encoder(c, wav, a0, L));
    RUN(run_bottleneck(c, a0, a1, false));
    // refinement_module.py:45-62 with fusion_repeats = 1
    float *cur = xb, *other = xa;
    if (fused_caf) {
        if (!forked) RUN(run_caf_video(c, video));
        RUN(run_block(c, a1, repeats > 1 ? a1 : nullptr, xb, true));
    } else {
        RUN(run_block(c, a1, nullptr, xa));
        RUN(run_caf(c, xa, video, repeats > 1 ? a1 : nullptr, xb));
    }
    for (int i = 1; i < repeats; ++i) {
        RUN(run_block(c, cur, (i + 1 < repeats) ? a1 : nullptr, other));
        float* t = cur;
        cur = other;
        other = t;
    }
    static const bool unfused_md = env_flag("RTFS_UNFUSED_MASK_DEC");  // A/B: mask writes z, the decoder GEMM reads it back
    if (use_tc() && use_persistent(3) && !unfused_md) {
        RUN(run_mask_dec(c, cur, a0));
        RUN(run_decoder(c, nullptr, out, L));
    } else {
        RUN(run_mask(c, cur, a0, a1));  // a1 is dead by now: reuse it for z
        RUN(run_decoder(c, a1, out, L));
    }
    return 0;
}

int rtfs_avnet_forward(const float* const* params, const float* wav, const float* video, float* out, void* ws, int B, int L, int Tv, int repeats, void* stream) {
    return avnet_forward_impl(params, wav, nullptr, const_cast<float*>(video), out, ws, B, L, Tv, repeats, stream, nullptr);
}

int rtfs_avnet_forward_av(const float* const* params, const float* wav, const float* mouth, float* video, float* out, void* ws, int B, int L, int Tv,
                          int repeats, void* stream, void* side_stream) {
    if (mouth == nullptr || video == nullptr) return fail_msg("rtfs_avnet_forward_av: mouth / video buffer missing");
    return avnet_forward_impl(params, wav, mouth, video, out, ws, B, L, Tv, repeats, stream, side_stream);
}

}  // extern "C"

#include "train_api.cuh"
