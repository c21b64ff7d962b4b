/* rtfs_b200.h -- C ABI of librtfs_b200.so: the RTFS-Net model-forward hot path on B200 (sm_100a).
 *
 * The reference (spkgyk/RTFS-Net) is pure Python: its "FFI" for this path is the set of
 * torch/cuDNN/cuBLAS/cuFFT/`sru` kernels its nn.Modules launch.  Each entry point below replaces
 * the launches of one reference nn.Module.forward (cited as file:line under
 * /root/reference/src/models/), takes raw device pointers + sizes + a CUDA stream, allocates
 * nothing, and returns 0 on success or a negative code (rtfs_last_error() gives the text).
 * The Python binding a maintainer adds (ctypes) is shown in INTEGRATION.md and implemented in
 * rtfs_net_b200/_lib.py.
 *
 * Layout.  Every activation is fp32 channels-last: a logical (B,C,T,F) tensor is stored
 * (B,T,F,C) (torch: memory_format=channels_last).  F = 129 (n_fft 256), Fc = 64, hidden = 64,
 * bottleneck = 256 are fixed by the RTFS-Net configs (config/lrs2_RTFSNet_*_layer.yaml).
 *   T  = L/128 + 1           frames          Tc = (T-2)/2 + 1      compressed frames
 * Weights are passed as an array of RTFS_P_COUNT device pointers prepared by the host
 * (rtfs_net_b200/weights.py) in the order of enum rtfs_param.
 */
#ifndef RTFS_B200_H
#define RTFS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define RTFS_ABI_VERSION 6
#define RTFS_F 129
#define RTFS_FC 64

/* Prepared-parameter slots (all fp32, contiguous, device memory). */
enum rtfs_param {
    RTFS_P_WINDOW = 0, /* [256] hann periodic (encoder.py:159 / decoder.py:108) */
    RTFS_P_COSTAB,     /* [256] cos(2 pi k/256) */
    RTFS_P_SINTAB,     /* [256] sin(2 pi k/256) */
    RTFS_P_ENC_W,      /* [256][32]  encoder conv, k=(i*3+j)*2+ci, zero-padded K */
    RTFS_P_BN_GAMMA, RTFS_P_BN_BETA, /* [256] audio_bottleneck gLN */
    RTFS_P_BN_W,       /* [256][256] tf32 */
    RTFS_P_BN_B,       /* [256] */
    RTFS_P_GW_W, RTFS_P_GW_B, RTFS_P_GW_A,   /* gateway dw1x1 [256],[256], PReLU [1] */
    RTFS_P_PJ_W, RTFS_P_PJ_B, RTFS_P_PJ_GAMMA, RTFS_P_PJ_BETA, RTFS_P_PJ_A, /* projection [64][256] tf32 ... */
    RTFS_P_D0_W, RTFS_P_D0_B, RTFS_P_D0_GAMMA, RTFS_P_D0_BETA, /* downsample 0: [16][64] tap-major, [64]x3 */
    RTFS_P_D1_W, RTFS_P_D1_B, RTFS_P_D1_GAMMA, RTFS_P_D1_BETA,
    /* dual-path RNN, frequency path (globalatt.0) */
    RTFS_P_RF_LNG, RTFS_P_RF_LNB,
    RTFS_P_RF_W0, RTFS_P_RF_WC0, RTFS_P_RF_B0, /* [256][512] tf32 (rows m*64+col, K = tap*64+c), [128], [128] */
    RTFS_P_RF_W1, RTFS_P_RF_WC1, RTFS_P_RF_B1, /* [192][64] tf32 */
    RTFS_P_RF_W2, RTFS_P_RF_WC2, RTFS_P_RF_B2,
    RTFS_P_RF_W3, RTFS_P_RF_WC3, RTFS_P_RF_B3,
    RTFS_P_RF_CTW, RTFS_P_RF_CTB,              /* [64][512] tf32 (K = (7-tap)*64+ci), [64] */
    /* dual-path RNN, time path (globalatt.1) */
    RTFS_P_RT_LNG, RTFS_P_RT_LNB,
    RTFS_P_RT_W0, RTFS_P_RT_WC0, RTFS_P_RT_B0,
    RTFS_P_RT_W1, RTFS_P_RT_WC1, RTFS_P_RT_B1,
    RTFS_P_RT_W2, RTFS_P_RT_WC2, RTFS_P_RT_B2,
    RTFS_P_RT_W3, RTFS_P_RT_WC3, RTFS_P_RT_B3,
    RTFS_P_RT_CTW, RTFS_P_RT_CTB,
    /* TF self-attention (globalatt.2) */
    RTFS_P_AT_WQKV, RTFS_P_AT_BQKV, RTFS_P_AT_SLOPE, RTFS_P_AT_GAMMA, RTFS_P_AT_BETA, /* [96][64] tf32,[96],[12],[6144],[6144] */
    RTFS_P_AT_WO, RTFS_P_AT_BO, RTFS_P_AT_SLOPEO, RTFS_P_AT_GAMMAO, RTFS_P_AT_BETAO,  /* [64][64] tf32,[64],[1],[4096],[4096] */
    /* TF-AR units: fusion_layers.0, fusion_layers.1, concat_layers.0 : {local, embedding, gate} x {w[16][64], gamma, beta} */
    RTFS_P_F0_LW, RTFS_P_F0_LG, RTFS_P_F0_LB, RTFS_P_F0_EW, RTFS_P_F0_EG, RTFS_P_F0_EB, RTFS_P_F0_GW, RTFS_P_F0_GG, RTFS_P_F0_GB,
    RTFS_P_F1_LW, RTFS_P_F1_LG, RTFS_P_F1_LB, RTFS_P_F1_EW, RTFS_P_F1_EG, RTFS_P_F1_EB, RTFS_P_F1_GW, RTFS_P_F1_GG, RTFS_P_F1_GB,
    RTFS_P_C0_LW, RTFS_P_C0_LG, RTFS_P_C0_LB, RTFS_P_C0_EW, RTFS_P_C0_EG, RTFS_P_C0_EB, RTFS_P_C0_GW, RTFS_P_C0_GG, RTFS_P_C0_GB,
    RTFS_P_RC_W, RTFS_P_RC_B, /* residual_conv [256][64] tf32, [256] */
    /* CAF */
    RTFS_P_CAF_WR, RTFS_P_CAF_BR, RTFS_P_CAF_GR, RTFS_P_CAF_BER, /* resize: [256][2],[256],[256],[256] */
    RTFS_P_CAF_WA, RTFS_P_CAF_BA, RTFS_P_CAF_GA, RTFS_P_CAF_BEA, /* attention_embed: [1024][2],[1024]x3 */
    RTFS_P_CAF_SK, RTFS_P_CAF_TK, RTFS_P_CAF_SV, RTFS_P_CAF_TV,  /* folded dw1x1+BatchNorm2d(eval) of key/value_embed */
    /* S3 mask + decoder */
    RTFS_P_MK_A, RTFS_P_MK_W, RTFS_P_MK_B, /* PReLU [1]; [256][256] tf32 rows interleaved (2c: real c, 2c+1: imag c+128); bias likewise */
    RTFS_P_DEC_W,                          /* [18][256]: row o*9+i*3+j = ConvTranspose2d weight[:, o, i, j] */
    /* tcgen05 operand images of the GEMM weights above: W[N][K] (tf32) stored [K/4][N][4], i.e. every
     * 32-wide K chunk is one contiguous N*128-byte slab in the UMMA K-major no-swizzle core-matrix layout */
    RTFS_P_BN_WI, RTFS_P_PJ_WI, RTFS_P_RC_WI, RTFS_P_MK_WI,
    RTFS_P_RF_WI0, RTFS_P_RF_WI1, RTFS_P_RF_WI2, RTFS_P_RF_WI3, RTFS_P_RF_CTWI,
    RTFS_P_RT_WI0, RTFS_P_RT_WI1, RTFS_P_RT_WI2, RTFS_P_RT_WI3, RTFS_P_RT_CTWI,
    /* fused dual-path RNN kernel: 52 slabs of 16 KB = SRU layer 0 (32), layers 1-3 (4 each), ConvTranspose1d (8),
     * SRU slabs = [acc 2][K piece 4][TMEM lane 128][4] with lanes (candidate | reset) and (forget | highway) */
    RTFS_P_RF_FUSED, RTFS_P_RT_FUSED,
    RTFS_P_ENC_WI3, /* encoder conv, 3xTF32 split: image of [W_hi | W_hi | W_lo] (K = 96) */
    RTFS_P_AT_WQKVI, RTFS_P_AT_WOI, /* tcgen05 images of the attention conv weights: [16][96][4], [16][64][4] */
    RTFS_P_COUNT
};

/* Workspace buffers (offsets in bytes from the workspace base, filled by rtfs_ws_plan). */
enum rtfs_ws {
    RTFS_WS_SPEC = 0, RTFS_WS_A0, RTFS_WS_A1, RTFS_WS_XA, RTFS_WS_XB,
    RTFS_WS_P_PRE, RTFS_WS_D0_PRE, RTFS_WS_LE0_PRE, RTFS_WS_LEC_PRE,
    RTFS_WS_D1_PRE, RTFS_WS_POOL, RTFS_WS_G0, RTFS_WS_G1, RTFS_WS_G2, RTFS_WS_G3,
    RTFS_WS_N, RTFS_WS_HA, RTFS_WS_HB, RTFS_WS_HPAD, RTFS_WS_U, RTFS_WS_AO,
    RTFS_WS_Q, RTFS_WS_K, RTFS_WS_V,
    RTFS_WS_GE0, RTFS_WS_GG0, RTFS_WS_GE1, RTFS_WS_GG1, RTFS_WS_LE1, RTFS_WS_GEC, RTFS_WS_GGC,
    RTFS_WS_Q18, RTFS_WS_STATS, RTFS_WS_VK, RTFS_WS_ATT,
    RTFS_WS_COUNT
};

/* gLN statistic slots inside RTFS_WS_STATS: doubles [slot][B][2] = (sum, sum of squares). */
enum rtfs_stat {
    RTFS_ST_A0 = 0, RTFS_ST_PJ, RTFS_ST_D0, RTFS_ST_D1,
    RTFS_ST_F0L, RTFS_ST_F0E, RTFS_ST_F0G, RTFS_ST_F1L, RTFS_ST_F1E, RTFS_ST_F1G,
    RTFS_ST_C0L, RTFS_ST_C0E, RTFS_ST_C0G,
    RTFS_ST_COUNT
};

/* Stages of the forward, for the optional per-stage device timing (rtfs_profile_*). */
enum rtfs_stage {
    RTFS_SG_STFT = 0, RTFS_SG_ENC_CONV, RTFS_SG_BOTTLENECK,
    RTFS_SG_GATE_PROJ, RTFS_SG_DW_S1, RTFS_SG_DW_S2_POOL,
    RTFS_SG_DPRNN_PREP, RTFS_SG_DPRNN_GEMM0, RTFS_SG_DPRNN_SCAN, RTFS_SG_DPRNN_GEMML, RTFS_SG_DPRNN_CONVT,
    RTFS_SG_ATT_QKV, RTFS_SG_ATT_CORE, RTFS_SG_ATT_PROJ,
    RTFS_SG_TFAR_GLOBAL, RTFS_SG_TFAR_LE0, RTFS_SG_TFAR_CAT_GLOBAL, RTFS_SG_TFAR_CAT_LOCAL, RTFS_SG_RESID_OUT,
    RTFS_SG_CAF_VIDEO, RTFS_SG_CAF_APPLY, RTFS_SG_MASK, RTFS_SG_DEC_GEMM, RTFS_SG_DEC_ISTFT,
    RTFS_SG_DPRNN_FUSED, /* one launch per dual-path RNN (dprnn_fused.cuh) instead of PREP..CONVT */
    RTFS_SG_RESID_OUT_CAF, /* residual conv of the first block pass with the CAF fusion in its epilogue (addend aliases x) */
    RTFS_SG_COUNT
};

int rtfs_abi_version(void);
const char* rtfs_last_error(void);

/* Size (bytes) of the workspace for B utterances of L samples with Tv video frames; offsets
 * (RTFS_WS_COUNT entries, bytes) may be NULL. */
long long rtfs_ws_plan(int B, int L, int Tv, long long* offsets);

/* STFTEncoder.forward (TDAVNet/encoder.py:161-175): wav (B,L) -> a0 (B,T,F,256); also leaves the
 * gLN statistics of a0 in stat slot RTFS_ST_A0 for the bottleneck. */
int rtfs_encoder_forward(const float* const* params, const float* wav, float* a0, void* ws, int B, int L, void* stream);

/* ConvNormAct audio_bottleneck (layers/conv_layers.py:65-129, built tdavnet.py:59):
 * a1 = conv1x1(ReLU(gLN(a0))) + b.  Computes the statistics of a0 itself. */
int rtfs_bottleneck_forward(const float* const* params, const float* a0, float* a1, void* ws, int B, int T, void* stream);

/* TDANetBlock.forward, is2d (separators/tdanet.py:106-133): out = Blk(x) [+ addend].
 * x, out, addend: (B,T,F,256); out must not alias x.  addend may be NULL. */
int rtfs_block_forward(const float* const* params, const float* x, const float* addend, float* out, void* ws, int B, int T, void* stream);

/* DualPathRNN.forward (layers/rnn_layers.py:136-162) on g (B,Tc,64,64): which = 0 frequency
 * path (dim=4), 1 time path (dim=3).  T = full-resolution frame count the workspace was planned
 * for (Tc = (T-2)/2+1). */
int rtfs_dprnn_forward(const float* const* params, int which, const float* g_in, float* g_out, void* ws, int B, int T, void* stream);

/* MultiHeadSelfAttention2D.forward (layers/attention.py:149-189) on g (B,Tc,64,64). */
int rtfs_mhsa_forward(const float* const* params, const float* g_in, float* g_out, void* ws, int B, int T, void* stream);

/* ATTNFusionCell.forward (layers/fusion.py:252-274): audio (B,T,F,256), video (B,512,Tv) ->
 * out (B,T,F,256) [+ addend]; eval-mode BatchNorm. */
int rtfs_caf_forward(const float* const* params, const float* audio, const float* video, const float* addend, float* out, void* ws, int B, int T, int Tv, void* stream);

/* MaskGenerator.forward (TDAVNet/mask_generator.py:67-99): z (B,T,F,256) = S3 mask applied to a0. */
int rtfs_mask_forward(const float* const* params, const float* refined, const float* a0, float* z, int B, int T, void* stream);

/* STFTDecoder.forward (TDAVNet/decoder.py:110-132): z (B,T,F,256) -> wav (B,L). */
int rtfs_decoder_forward(const float* const* params, const float* z, float* wav_out, void* ws, int B, int L, void* stream);

/* AVNet.forward (tdavnet.py:86-97) + RefinementModule.forward (TDAVNet/refinement_module.py:45-62)
 * for fusion_repeats = 1: wav (B,L), video = output of the video block (B,512,Tv) -> out (B,L).
 * repeats = audio_params.repeats (4 / 6 / 12). */
int rtfs_avnet_forward(const float* const* params, const float* wav, const float* video, float* out, void* ws, int B, int L, int Tv, int repeats, void* stream);

/* Number of kernels the last rtfs_avnet_forward on this thread launched (bench gpu_launches). */
long long rtfs_last_launch_count(void);

/* Per-stage device timing: when enabled, every stage is bracketed by cudaEvents on the launch
 * stream.  rtfs_profile_collect synchronises the device, adds the elapsed milliseconds and launch
 * counts per stage into ms[RTFS_SG_COUNT] / count[RTFS_SG_COUNT], and resets the recorder. */
void rtfs_profile_enable(int on);
int rtfs_profile_collect(float* ms, int* count);

#ifdef __cplusplus
}
#endif
#endif
