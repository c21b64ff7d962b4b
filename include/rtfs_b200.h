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

#define RTFS_ABI_VERSION 8
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
    /* training only (may be NULL for inference): transposed images for the data-gradient GEMMs dX = dY * W, W'[N = K_fwd][K = N_fwd] */
    RTFS_P_BN_WT,    /* [256][256] */
    RTFS_P_PJ_WT,    /* [256][64]  */
    RTFS_P_RC_WT,    /* [64][256]  */
    RTFS_P_MK_WT,    /* [256][256], output channels (K) in NATURAL order (0..127 real, 128..255 imag) */
    RTFS_P_RF_W0T, RTFS_P_RF_W1T, RTFS_P_RF_W2T, RTFS_P_RF_W3T, /* [512][256], [64][192] x3 */
    RTFS_P_RF_CTWB,  /* [64 ci][tap*64 + co] = ConvTranspose1d weight[ci][co][tap] */
    RTFS_P_RT_W0T, RTFS_P_RT_W1T, RTFS_P_RT_W2T, RTFS_P_RT_W3T,
    RTFS_P_RT_CTWB,
    RTFS_P_AT_WQKVT, /* [64][96] */
    RTFS_P_AT_WOT,   /* [64][64] */
    RTFS_P_DEC_WE,   /* [256][32]: decoder weight in the encoder-conv layout, k = (i*3+j)*2 + o (dz = conv2d(dy, W_dec)) */
    /* fused S^3 mask + decoder epilogue: [256][20], row = interleaved GEMM column (2c: channel c, 2c+1: channel c+128),
     * entries 0..17 = RTFS_P_DEC_W[:, channel], 18..19 = 0 */
    RTFS_P_DEC_WT,
    /* VP (video) block, all parameters packed in the order of rtfs_video_pack_plan (eval BatchNorm folded; may be NULL when the
     * host runs the video block with torch modules: training, Tv > 100, non-RTFS video configurations) */
    RTFS_P_VIDEO_PACK,
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
    /* training tape of the two dual-path RNNs of a block pass (size 0 in the inference plan): LN output, and per SRU layer
     * the gate pre-activations U, the cell states C and the layer outputs H (the last one zero-padded: HP) */
    RTFS_WS_TF_N, RTFS_WS_TF_U0, RTFS_WS_TF_U1, RTFS_WS_TF_U2, RTFS_WS_TF_U3,
    RTFS_WS_TF_C0, RTFS_WS_TF_C1, RTFS_WS_TF_C2, RTFS_WS_TF_C3, RTFS_WS_TF_H0, RTFS_WS_TF_H1, RTFS_WS_TF_H2, RTFS_WS_TF_HP,
    RTFS_WS_TT_N, RTFS_WS_TT_U0, RTFS_WS_TT_U1, RTFS_WS_TT_U2, RTFS_WS_TT_U3,
    RTFS_WS_TT_C0, RTFS_WS_TT_C1, RTFS_WS_TT_C2, RTFS_WS_TT_C3, RTFS_WS_TT_H0, RTFS_WS_TT_H1, RTFS_WS_TT_H2, RTFS_WS_TT_HP,
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
    RTFS_SG_MASK_DEC,      /* S^3 mask with the decoder's 256 -> 18 contraction in its epilogue (z never written) */
    RTFS_SG_VIDEO,         /* VP block kernel (module-level call; timed by the host around rtfs_video_forward) */
    RTFS_SG_COUNT
};

/* Training tape, global part (one per step; the per-pass parts follow, each laid out by rtfs_train_plan's pass offsets). */
enum rtfs_tape {
    RTFS_TP_SPEC = 0,  /* (B,T,F,2) STFT of the mixture */
    RTFS_TP_A0, RTFS_TP_A1, /* encoder / bottleneck outputs (B,T,F,256) */
    RTFS_TP_BLK0,      /* output of the first block pass = CAF audio input */
    RTFS_TP_X,         /* block inputs of passes 1..R-1 (R-1 tensors back to back); X_1 = CAF(BLK0) + a1 */
    RTFS_TP_REFINED,   /* output of the last pass */
    RTFS_TP_M,         /* S^3 mask m = ReLU(conv) in natural channel order */
    RTFS_TP_Z,         /* masked embedding (decoder input) */
    RTFS_TP_Q18, RTFS_TP_VK, RTFS_TP_ATT,
    RTFS_TP_CAFSUM,    /* doubles [256][2]: per-channel (sum, sum of squares) of BLK0 over this rank's batch */
    RTFS_TP_PASS0,     /* first per-pass region; region i starts at offsets[RTFS_TP_PASS0] + i * pass_bytes */
    RTFS_TP_COUNT
};

/* Backward scratch (gradients of activations; caller-provided, contents undefined between calls).  A = B*T*F*256 floats,
 * H = B*T*F*64, G = B*Tc*Fc*64. */
enum rtfs_bwd {
    RTFS_BW_DA = 0, RTFS_BW_DB, /* ping/pong: gradient w.r.t. a block output / input (A) */
    RTFS_BW_DA1,       /* accumulated gradient w.r.t. a1 (A) */
    RTFS_BW_DM,        /* gradient w.r.t. the mask pre-activation; earlier dz (A) */
    RTFS_BW_DA0,       /* gradient w.r.t. a0 (A) */
    RTFS_BW_DSPEC,     /* (B,T,F,2) */
    RTFS_BW_HE, RTFS_BW_HDE, RTFS_BW_HF0, RTFS_BW_HDF0, RTFS_BW_HT, /* H-sized */
    RTFS_BW_GF1, RTFS_BW_GDF1, RTFS_BW_GT1, RTFS_BW_GT2, RTFS_BW_GT3, RTFS_BW_GT4, RTFS_BW_GT5, RTFS_BW_GD1N, RTFS_BW_GDD1N,
    RTFS_BW_GDG3, RTFS_BW_GDG2, RTFS_BW_GDG1, RTFS_BW_GDG0, /* G-sized */
    RTFS_BW_GA1, RTFS_BW_GA2, RTFS_BW_GA3, RTFS_BW_GS, RTFS_BW_GDP, RTFS_BW_GDQ, RTFS_BW_GDK, RTFS_BW_GDPRE, /* attention */
    RTFS_BW_DZP, RTFS_BW_DHA, RTFS_BW_DHB, RTFS_BW_DU, RTFS_BW_DXIN, RTFS_BW_DXUNF, /* dual-path RNN */
    RTFS_BW_RED,       /* doubles: gLN backward sums [32][B][2] */
    RTFS_BW_DVK, RTFS_BW_DATT, /* (B,Tv,256) */
    RTFS_BW_CSUM,      /* doubles [256][4]: CAF BatchNorm backward channel sums (this rank) */
    RTFS_BW_COUNT
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

/* ---- input pipeline on the device (reference: the per-utterance numpy work of the DataLoader workers).
 * Lip-ROI transforms of get_preprocessing_pipelines() (src/datas/transform.py:151-167): Normalize(0,255) -> CenterCrop /
 * RandomCrop(crop) -> HorizontalFlip -> Normalize(mean, std).  roi (B,T,H,W) uint8 -> out (B,1,T,crop,crop) fp32.
 * off_y / off_x / flip: per-utterance crop offsets and flip decisions drawn by the host (NULL = centre crop / no flip,
 * transform.py:86-102); crop % 4 == 0. */
int rtfs_mouth_preprocess(const unsigned char* roi, float* out, const int* off_y, const int* off_x, const int* flip, int B, int T, int H, int W,
                          int crop, float mean, float std, void* stream);
/* normalize_tensor_wav as AVSpeechDataset.__getitem__ applies it (src/datas/avspeech_dataset.py:10-14,128-131,167-170):
 * mix (B,L) -> (x - mean) / (std + eps) with the unbiased std of the mixture; src (B,n_src,L) -> (s - mean_s) / (std_mix + eps).
 * n_src = 0: mixture only (src / src_out may be NULL). */
int rtfs_wav_normalize(const float* mix, const float* src, float* mix_out, float* src_out, int B, int L, int n_src, float eps, void* stream);

/* TDANetBlock.forward with is2d = False (separators/tdanet.py:106-133; GlobalAttention layers/attention.py:28-73,192-220): the VP
 * block on the lip embedding, x (B,512,Tv) -> out (B,512,Tv), one kernel, inference (eval BatchNorm).  8 <= Tv <= 100. */
int rtfs_video_forward(const float* const* params, const float* x, float* out, int B, int Tv, void* stream);
/* Float offsets of the fields of RTFS_P_VIDEO_PACK (offsets[n_fields]); returns the total float count, *n_fields the field count. */
int rtfs_video_pack_plan(int* offsets, int* n_fields);

/* AVNet.forward (tdavnet.py:86-97) + RefinementModule.forward (TDAVNet/refinement_module.py:45-62)
 * for fusion_repeats = 1: wav (B,L), video = output of the video block (B,512,Tv) -> out (B,L).
 * repeats = audio_params.repeats (4 / 6 / 12). */
int rtfs_avnet_forward(const float* const* params, const float* wav, const float* video, float* out, void* ws, int B, int L, int Tv, int repeats, void* stream);

/* The same forward from the raw lip embedding: mouth (B,512,Tv) -> VP block (rtfs_video_forward) -> video (B,512,Tv, caller's
 * buffer) -> the rest of rtfs_avnet_forward.  With a second stream the VP block and the CAF video branch (which depend on the
 * lip embedding only and fill 32 of 148 SMs) run on side_stream next to the first RTFS block pass (event fork / join, graph
 * capturable); side_stream = NULL runs everything on `stream`.  8 <= Tv <= 100. */
int rtfs_avnet_forward_av(const float* const* params, const float* wav, const float* mouth, float* video, float* out, void* ws, int B, int L,
                          int Tv, int repeats, void* stream, void* side_stream);

/* ---- training step (BASELINE configs[2]; reference call chain src/system/core.py:94-117 -> AVNet.forward -> loss.backward(),
 * train.py:135-146).  The forward keeps a tape, the backward consumes it.
 *
 * rtfs_train_plan: sizes (bytes) of the tape and of the backward scratch for B utterances of L samples, Tv video frames and
 * `repeats` block passes; tape_offsets[RTFS_TP_COUNT], bwd_offsets[RTFS_BW_COUNT] (bytes; may be NULL); *pass_bytes = size of
 * one per-pass region, laid out like rtfs_ws_plan's workspace (query it with rtfs_train_pass_plan).  Returns the tape size. */
long long rtfs_train_plan(int B, int L, int Tv, int repeats, long long* tape_offsets, long long* pass_bytes, long long* bwd_bytes, long long* bwd_offsets);
long long rtfs_train_pass_plan(int B, int L, long long* offsets /* RTFS_WS_COUNT */);

/* AVNet.forward in train() mode, split around the cross-rank reduction of the CAF BatchNorm statistics:
 *   phase 0: encoder, bottleneck, first block pass, CAF video vectors, per-channel sums of the CAF audio input (RTFS_TP_CAFSUM)
 *   -- the host forms the batch-statistics scale/shift of key_embed / value_embed (all-reducing the sums under SyncBatchNorm)
 *      and stores them in the RTFS_P_CAF_SK/TK/SV/TV parameter slots --
 *   phase 1: CAF, remaining passes, S^3 mask, decoder -> out (B,L).
 * video = output of the video block (B,512,Tv).  Phase 0 may be called with video = NULL: it then runs the audio-only part, and
 * the CAF video vectors are submitted later as phase 2 (before phase 1) -- the host can enqueue the audio kernels first and
 * prepare the video block's output while they run. */
int rtfs_avnet_train_forward(const float* const* params, const float* wav, const float* video, float* out, void* tape,
                             int B, int L, int Tv, int repeats, int phase, void* stream);

/* Backward of the above.  grads: RTFS_P_COUNT pointers to ZEROED buffers shaped like the parameter slots (NULL where no
 * gradient is wanted / the slot is a derived image); gradients are accumulated into them.  Exceptions to "shaped like the
 * slot": RTFS_P_MK_W / RTFS_P_MK_B gradients are written in NATURAL output-channel order.
 *   phase 0: d_out (B,L) -> decoder, mask, passes R-1..1, CAF reductions (RTFS_BW_CSUM, d_video (B,512,Tv), CAF video-side grads)
 *   -- the host all-reduces the channel sums under SyncBatchNorm and passes caf_mu / caf_c0 / caf_c1 ([256] each:
 *      batch mean of the CAF input, sk*m1k + sv*m1v, sk*m2k*wk/sigk + sv*m2v*wv/sigv) --
 *   phase 1: CAF input gradient, first block pass, bottleneck, encoder. */
int rtfs_avnet_backward(const float* const* params, float* const* grads, const float* wav, const float* video, const float* d_out,
                        float* d_video, const float* caf_mu, const float* caf_c0, const float* caf_c1, void* tape, void* scratch,
                        int B, int L, int Tv, int repeats, int phase, void* stream);

/* Module-level training entries (tests, partial use): the tape of a call is the per-pass region `pass_ws`. */
int rtfs_block_train_forward(const float* const* params, const float* x, const float* addend, float* out, void* pass_ws, int B, int T, void* stream);
int rtfs_block_backward(const float* const* params, float* const* grads, const float* x, const float* d_out, float* d_x,
                        void* pass_ws, void* scratch, int B, int T, void* stream);
int rtfs_dprnn_train_forward(const float* const* params, int which, const float* g_in, float* g_out, void* pass_ws, int B, int T, void* stream);
int rtfs_dprnn_backward(const float* const* params, float* const* grads, int which, const float* g_in, const float* d_out, float* d_in,
                        void* pass_ws, void* scratch, int B, int T, void* stream);
int rtfs_mhsa_train_forward(const float* const* params, const float* g_in, float* g_out, void* pass_ws, int B, int T, void* stream);
int rtfs_mhsa_backward(const float* const* params, float* const* grads, const float* g_in, const float* d_out, float* d_in,
                       void* pass_ws, void* scratch, int B, int T, void* stream);

/* PairwiseNegSDR("snr") for n_src = 1 (src/losses/matrix.py:22-53): loss[b] = -10 log10(sum t'^2 / (sum (e'-t')^2 + eps) + eps)
 * with zero-meaned signals; d_est (may be NULL) = scale * d loss[b] / d est.  sums: B*5 doubles of scratch. */
int rtfs_snr_loss(const float* est, const float* target, float* loss, float* d_est, double* sums, int B, int L, float scale, void* stream);

/* clip_grad_norm_(max_norm) + torch.optim.AdamW step over flat buffers (src/system/optimizers.py:58-75, train.py:143):
 * gnorm_sq (one double of scratch) receives sum(g^2) first; grad_scale multiplies the gradient (1/world after a sum all-reduce). */
int rtfs_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, float max_norm, float grad_scale, double* gnorm_sq, void* stream);

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
