"""Host-side preparation of the reference's parameters into the kernel layouts of
include/rtfs_b200.h (enum rtfs_param).  Input: a state_dict with the reference's key names
(SURVEY.md App. B); output: one contiguous fp32 tensor per slot on the target device.

Done once per parameter version (not on the hot path): TF32 rounding of GEMM weights (cvt.rna),
K/N permutations that turn nn.Unfold / ConvTranspose1d / channel concatenation into plain
row-major GEMMs, tap-major depthwise filters, folded eval-mode BatchNorm.
"""
import ctypes
import math

import torch

from . import _lib

BLK = "refinement_module.audio_net.blocks."
CAF = "refinement_module.crossmodal_fusion.fusion_module.audio_lstm."


def tf32_round(x: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32: round to nearest (ties away from zero) to a 10-bit mantissa.  Under autograd the rounding is a
    straight-through estimator (the gradient of the rounded image is the gradient of the fp32 master weight)."""
    i = x.detach().to(torch.float32).contiguous().view(torch.int32)
    i = (i + 0x1000) & -8192
    r = i.view(torch.float32)
    if x.requires_grad and torch.is_grad_enabled():
        return x + (r - x).detach()
    return r


def umma_image(w: torch.Tensor) -> torch.Tensor:
    """W[N][K] -> [K/4][N][4]: the tcgen05 K-major no-swizzle operand image (gemm_tc.cuh); each 32-wide
    K chunk is then one contiguous N*128-byte slab a single bulk copy drops into shared memory."""
    N, K = w.shape
    assert K % 32 == 0
    return w.reshape(N, K // 4, 4).permute(1, 0, 2).contiguous()


def dprnn_fused_image(w0, wl, ctw):
    """Weight slabs of the fused dual-path RNN kernel (dprnn_fused.cuh), 52 x 4096 floats.
    w0 [256][512] and wl[i] [192][64] have rows m*64+col (m: 0 candidate, 1 forget, 2 reset, 3 highway);
    ctw [64][512] has K = kk*64+ci.  SRU slab (16 K values): [acc 2][K piece 4][lane 128][4] with
    acc0 lanes = (m0 | m2), acc1 lanes = (m1 | m3 or zeros); conv slab kk: [K piece 16][co 64][4]."""
    def lanes(w, k):
        m = [w[i * 64:(i + 1) * 64] for i in range(k)]
        z = torch.zeros_like(m[0])
        return torch.cat([m[0], m[2]], 0), torch.cat([m[1], m[3] if k == 4 else z], 0)

    slabs = []
    for w, k in [(w0, 4)] + [(x, 3) for x in wl]:
        a0, a1 = lanes(w, k)
        for g in range(w.shape[1] // 16):
            for acc in (a0, a1):
                blk = acc[:, 16 * g:16 * g + 16]  # [128][16]
                slabs.append(blk.reshape(128, 4, 4).permute(1, 0, 2).reshape(-1))
    for kk in range(8):
        blk = ctw[:, 64 * kk:64 * kk + 64]  # [64 co][64]
        slabs.append(blk.reshape(64, 16, 4).permute(1, 0, 2).reshape(-1))
    out = torch.cat(slabs).contiguous()
    assert out.numel() == 52 * 4096
    return out


def _tapmajor(w):
    """depthwise (C,1,4,4) -> [16][C]"""
    C = w.shape[0]
    return w.reshape(C, -1).t().contiguous()


# slots that only the training step reads (transposed images for the data-gradient GEMMs): NULL in the inference table
TRAIN_ONLY = ("RTFS_P_BN_WT", "RTFS_P_PJ_WT", "RTFS_P_RC_WT", "RTFS_P_MK_WT", "RTFS_P_RF_W0T", "RTFS_P_RF_W1T", "RTFS_P_RF_W2T",
              "RTFS_P_RF_W3T", "RTFS_P_RF_CTWB", "RTFS_P_RT_W0T", "RTFS_P_RT_W1T", "RTFS_P_RT_W2T", "RTFS_P_RT_W3T", "RTFS_P_RT_CTWB",
              "RTFS_P_AT_WQKVT", "RTFS_P_AT_WOT", "RTFS_P_DEC_WE")
# slots that are images / transposes of another slot: the backward returns no gradient for them (the base slot carries it)
DERIVED = TRAIN_ONLY[:-1] + ("RTFS_P_BN_WI", "RTFS_P_PJ_WI", "RTFS_P_RC_WI", "RTFS_P_MK_WI", "RTFS_P_RF_WI0", "RTFS_P_RF_WI1", "RTFS_P_RF_WI2",
                             "RTFS_P_RF_WI3", "RTFS_P_RF_CTWI", "RTFS_P_RT_WI0", "RTFS_P_RT_WI1", "RTFS_P_RT_WI2", "RTFS_P_RT_WI3", "RTFS_P_RT_CTWI",
                             "RTFS_P_RF_FUSED", "RTFS_P_RT_FUSED", "RTFS_P_ENC_WI3", "RTFS_P_AT_WQKVI", "RTFS_P_AT_WOI", "RTFS_P_DEC_W", "RTFS_P_DEC_WT", "RTFS_P_VIDEO_PACK",
                             "RTFS_P_WINDOW", "RTFS_P_COSTAB", "RTFS_P_SINTAB", "RTFS_P_CAF_SK", "RTFS_P_CAF_TK", "RTFS_P_CAF_SV", "RTFS_P_CAF_TV")


VID = "refinement_module.video_net.blocks."


def video_pack_supported(sd):
    """The VP-block kernel (csrc/video.cuh) is built for the RTFS-Net video configuration: 512 -> 64 channels, four k=3 scales,
    BatchNorm1d, GlobalAttention with 8 heads and a 128-channel FFN."""
    try:
        return (tuple(sd[VID + "projection.full_layer.2.weight"].shape) == (64, 512, 1)
                and tuple(sd[VID + "downsample_layers.3.full_layer.2.weight"].shape) == (64, 1, 3)
                and (VID + "downsample_layers.4.full_layer.2.weight") not in sd
                and (VID + "downsample_layers.0.full_layer.3.running_mean") in sd
                and tuple(sd[VID + "globalatt.0.MHSA.attention.in_proj_weight"].shape) == (192, 64)
                and tuple(sd[VID + "globalatt.0.FFN.encoder.full_layer.2.weight"].shape) == (128, 64, 1)
                and tuple(sd[VID + "globalatt.0.FFN.refiner.full_layer.2.weight"].shape) == (128, 1, 3)
                and (VID + "globalatt.1.MHSA.norm1.weight") not in sd)
    except KeyError:
        return False


def pack_video(sd, device, n_head=8):
    """All parameters of the VP block in the layout of csrc/video.cuh (enum vp_field): transposed 1x1-conv weights, tap-major
    depthwise filters, eval BatchNorm1d (and the conv bias in front of it) folded into per-channel scale / shift."""
    g = lambda k: sd[VID + k].detach().to(device=device, dtype=torch.float32)
    has = lambda k: (VID + k) in sd
    offs, total = _lib.video_pack_plan()
    buf = torch.zeros(total, device=device, dtype=torch.float32)

    def put(field, t, at=0):
        t = t.reshape(-1)
        buf[offs[field] + at: offs[field] + at + t.numel()] = t

    def bn_fold(q, conv_bias=None):
        s = g(q + "weight") / torch.sqrt(g(q + "running_var") + 1e-5)
        b = conv_bias if conv_bias is not None else 0.0
        return s, (b - g(q + "running_mean")) * s + g(q + "bias")

    put("GW_W", g("gateway.full_layer.2.weight"))
    put("GW_B", g("gateway.full_layer.2.bias"))
    put("GW_A", g("gateway.full_layer.4.weight"))
    put("PJ_WT", g("projection.full_layer.2.weight").reshape(64, 512).t().contiguous())
    s, t = bn_fold("projection.full_layer.3.", g("projection.full_layer.2.bias"))
    put("PJ_S", s)
    put("PJ_T", t)
    put("PJ_A", g("projection.full_layer.4.weight"))
    for i in range(4):
        q = f"downsample_layers.{i}.full_layer."
        put("DS_W", g(q + "2.weight").reshape(64, 3).t().contiguous(), i * 192)
        s, t = bn_fold(q + "3.", g(q + "2.bias"))
        put("DS_S", s, i * 64)
        put("DS_T", t, i * 64)
    a = "globalatt.0.MHSA."
    put("LN1_G", g(a + "norm1.weight"))
    put("LN1_B", g(a + "norm1.bias"))
    if has(a + "pos_enc.pe"):
        put("PE", g(a + "pos_enc.pe")[0, :16])
    put("IN_WT", g(a + "attention.in_proj_weight").t().contiguous())
    put("IN_B", g(a + "attention.in_proj_bias"))
    put("OUT_WT", g(a + "attention.out_proj.weight").t().contiguous())
    put("OUT_B", g(a + "attention.out_proj.bias"))
    put("LN2_G", g(a + "norm2.weight"))
    put("LN2_B", g(a + "norm2.bias"))
    f = "globalatt.0.FFN."
    put("F1_WT", g(f + "encoder.full_layer.2.weight").reshape(128, 64).t().contiguous())
    put("F1_G", g(f + "encoder.full_layer.3.norm.weight"))
    put("F1_B", g(f + "encoder.full_layer.3.norm.bias"))
    put("FR_W", g(f + "refiner.full_layer.2.weight").reshape(128, 3).t().contiguous())
    put("FR_B", g(f + "refiner.full_layer.2.bias"))
    put("F2_WT", g(f + "decoder.full_layer.2.weight").reshape(64, 128).t().contiguous())
    put("F2_G", g(f + "decoder.full_layer.3.norm.weight"))
    put("F2_B", g(f + "decoder.full_layer.3.norm.bias"))
    units = [f"fusion_layers.{i}." for i in range(4)] + [f"concat_layers.{i}." for i in range(3)]
    for u, key in enumerate(units):
        for j, sub in enumerate(("local_embedding.", "global_embedding.", "global_gate.")):
            q = key + sub + "full_layer."
            put("TF", g(q + "2.weight").reshape(64, 3).t().contiguous(), u * 960 + j * 320)
            s, t = bn_fold(q + "3.")
            put("TF", s, u * 960 + j * 320 + 192)
            put("TF", t, u * 960 + j * 320 + 256)
    put("RC_WT", g("residual_conv.full_layer.2.weight").reshape(512, 64).t().contiguous())
    put("RC_B", g("residual_conv.full_layer.2.bias"))
    return buf


_CONSTS = {}


def _device_constants(device):
    """Parameter-independent tables, built once per device.  They used to be rebuilt on the host and copied in every prepare() call:
    three pageable host -> device copies, each of which makes the host wait for everything queued before it -- in training that
    is the whole backward of the previous step (tools/prof_train_dp.py: three cudaStreamSynchronize per step, 14 ms each)."""
    key = str(device)
    c = _CONSTS.get(key)
    if c is None:
        k = torch.arange(256, dtype=torch.float64)
        c = {"window": torch.hann_window(256, periodic=True, dtype=torch.float32, device=device),
             "cos": torch.cos(2.0 * math.pi * k / 256).to(torch.float32).to(device),
             "sin": torch.sin(2.0 * math.pi * k / 256).to(torch.float32).to(device),
             "mask_perm": torch.stack([torch.arange(128), torch.arange(128) + 128], 1).reshape(-1).to(device)}
        _CONSTS[key] = c
    return c


def prepare(sd, device, train=False):
    """{reference key: tensor} -> {slot name: tensor}.  train=True: the inputs are the live parameters and every slot is a
    differentiable function of them (reshapes / permutes / straight-through TF32 rounding), so autograd carries the slot
    gradients the CUDA backward produces back to the parameters; the training-only transposed images are added and the
    CAF BatchNorm scale / shift slots are plain buffers that the runtime fills from the batch statistics."""
    if train:
        g = lambda k: sd[k].to(device=device, dtype=torch.float32)
    else:
        g = lambda k: sd[k].detach().to(device=device, dtype=torch.float32)
    out = {}
    consts = _device_constants(device)
    out["RTFS_P_WINDOW"], out["RTFS_P_COSTAB"], out["RTFS_P_SINTAB"] = consts["window"], consts["cos"], consts["sin"]

    w = g("encoder.conv.full_layer.2.weight")  # (256,2,3,3) [co,ci,i,j] -> k = (i*3+j)*2+ci
    out["RTFS_P_ENC_W"] = torch.cat([w.permute(0, 2, 3, 1).reshape(256, 18), torch.zeros(256, 14, device=device)], 1)

    out["RTFS_P_BN_GAMMA"] = g("audio_bottleneck.full_layer.0.norm.weight")
    out["RTFS_P_BN_BETA"] = g("audio_bottleneck.full_layer.0.norm.bias")
    out["RTFS_P_BN_W"] = tf32_round(g("audio_bottleneck.full_layer.2.weight").reshape(256, 256))
    out["RTFS_P_BN_B"] = g("audio_bottleneck.full_layer.2.bias")

    out["RTFS_P_GW_W"] = g(BLK + "gateway.full_layer.2.weight").reshape(-1)
    out["RTFS_P_GW_B"] = g(BLK + "gateway.full_layer.2.bias")
    out["RTFS_P_GW_A"] = g(BLK + "gateway.full_layer.4.weight").reshape(1)
    out["RTFS_P_PJ_W"] = tf32_round(g(BLK + "projection.full_layer.2.weight").reshape(64, 256))
    out["RTFS_P_PJ_B"] = g(BLK + "projection.full_layer.2.bias")
    out["RTFS_P_PJ_GAMMA"] = g(BLK + "projection.full_layer.3.norm.weight")
    out["RTFS_P_PJ_BETA"] = g(BLK + "projection.full_layer.3.norm.bias")
    out["RTFS_P_PJ_A"] = g(BLK + "projection.full_layer.4.weight").reshape(1)
    for i in (0, 1):
        q = BLK + f"downsample_layers.{i}.full_layer."
        out[f"RTFS_P_D{i}_W"] = _tapmajor(g(q + "2.weight"))
        out[f"RTFS_P_D{i}_B"] = g(q + "2.bias")
        out[f"RTFS_P_D{i}_GAMMA"] = g(q + "3.norm.weight")
        out[f"RTFS_P_D{i}_BETA"] = g(q + "3.norm.bias")

    for r, tag in ((0, "RF"), (1, "RT")):
        q = BLK + f"globalatt.{r}."
        out[f"RTFS_P_{tag}_LNG"] = g(q + "norm.gamma").reshape(64)
        out[f"RTFS_P_{tag}_LNB"] = g(q + "norm.beta").reshape(64)
        w0 = g(q + "rnn.rnn_lst.0.weight")  # (512,256): rows c*8+tap, cols col*4+m
        # -> [m*64+col][tap*64+c]
        out[f"RTFS_P_{tag}_W0"] = tf32_round(w0.view(64, 8, 64, 4).permute(3, 2, 1, 0).reshape(256, 512))
        out[f"RTFS_P_{tag}_WC0"] = g(q + "rnn.rnn_lst.0.weight_c")
        out[f"RTFS_P_{tag}_B0"] = g(q + "rnn.rnn_lst.0.bias")
        for l in (1, 2, 3):
            wl = g(q + f"rnn.rnn_lst.{l}.weight")  # (64,192): rows ci, cols col*3+m -> [m*64+col][ci]
            out[f"RTFS_P_{tag}_W{l}"] = tf32_round(wl.view(64, 64, 3).permute(2, 1, 0).reshape(192, 64))
            out[f"RTFS_P_{tag}_WC{l}"] = g(q + f"rnn.rnn_lst.{l}.weight_c")
            out[f"RTFS_P_{tag}_B{l}"] = g(q + f"rnn.rnn_lst.{l}.bias")
        wct = g(q + "linear.weight")  # ConvTranspose1d (ci, co, tap) -> [co][(7-tap)*64+ci]
        out[f"RTFS_P_{tag}_CTW"] = tf32_round(wct.flip(2).permute(1, 2, 0).reshape(64, 512))
        out[f"RTFS_P_{tag}_CTB"] = g(q + "linear.bias")

    a = BLK + "globalatt.2."
    n_head = 0
    while (a + f"Queries.{n_head}.conv.weight") in sd:
        n_head += 1
    if n_head != 4:
        raise NotImplementedError("the attention kernels are built for n_head = 4 (RTFS-Net configs)")
    ws, bs, sl, gm, bt = [], [], [], [], []
    for kind in ("Queries", "Keys", "Values"):
        for h in range(n_head):
            p = a + f"{kind}.{h}."
            wk = g(p + "conv.weight")
            E = wk.shape[0]
            ws.append(wk.reshape(E, 64))
            bs.append(g(p + "conv.bias"))
            sl.append(g(p + "act.weight").reshape(1))
            gm.append(g(p + "norm.gamma").reshape(E, 64).t().reshape(-1))  # [f*E+e]
            bt.append(g(p + "norm.beta").reshape(E, 64).t().reshape(-1))
    out["RTFS_P_AT_WQKV"] = tf32_round(torch.cat(ws, 0))
    out["RTFS_P_AT_BQKV"] = torch.cat(bs)
    out["RTFS_P_AT_SLOPE"] = torch.cat(sl)
    out["RTFS_P_AT_GAMMA"] = torch.cat(gm)
    out["RTFS_P_AT_BETA"] = torch.cat(bt)
    p = a + "attn_concat_proj."
    out["RTFS_P_AT_WO"] = tf32_round(g(p + "conv.weight").reshape(64, 64))
    out["RTFS_P_AT_BO"] = g(p + "conv.bias")
    out["RTFS_P_AT_SLOPEO"] = g(p + "act.weight").reshape(1)
    out["RTFS_P_AT_GAMMAO"] = g(p + "norm.gamma").reshape(64, 64).t().reshape(-1)  # [f*64+c]
    out["RTFS_P_AT_BETAO"] = g(p + "norm.beta").reshape(64, 64).t().reshape(-1)

    for tag, key in (("F0", "fusion_layers.0."), ("F1", "fusion_layers.1."), ("C0", "concat_layers.0.")):
        for x, sub in (("L", "local_embedding."), ("E", "global_embedding."), ("G", "global_gate.")):
            q = BLK + key + sub + "full_layer."
            out[f"RTFS_P_{tag}_{x}W"] = _tapmajor(g(q + "2.weight"))
            out[f"RTFS_P_{tag}_{x}G"] = g(q + "3.norm.weight")
            out[f"RTFS_P_{tag}_{x}B"] = g(q + "3.norm.bias")
    out["RTFS_P_RC_W"] = tf32_round(g(BLK + "residual_conv.full_layer.2.weight").reshape(256, 64))
    out["RTFS_P_RC_B"] = g(BLK + "residual_conv.full_layer.2.bias")

    out["RTFS_P_CAF_WR"] = g(CAF + "resize.full_layer.2.weight").reshape(-1)
    out["RTFS_P_CAF_BR"] = g(CAF + "resize.full_layer.2.bias")
    out["RTFS_P_CAF_GR"] = g(CAF + "resize.full_layer.3.norm.weight")
    out["RTFS_P_CAF_BER"] = g(CAF + "resize.full_layer.3.norm.bias")
    out["RTFS_P_CAF_WA"] = g(CAF + "attention_embed.full_layer.2.weight").reshape(-1)
    out["RTFS_P_CAF_BA"] = g(CAF + "attention_embed.full_layer.2.bias")
    out["RTFS_P_CAF_GA"] = g(CAF + "attention_embed.full_layer.3.norm.weight")
    out["RTFS_P_CAF_BEA"] = g(CAF + "attention_embed.full_layer.3.norm.bias")
    for name, s_slot, t_slot in (("key_embed", "SK", "TK"), ("value_embed", "SV", "TV")):
        q = CAF + name + ".full_layer."
        if train:  # filled by the runtime (batch statistics in train(), running statistics in eval())
            out[f"RTFS_P_CAF_{s_slot}"] = torch.zeros(256, device=device)
            out[f"RTFS_P_CAF_{t_slot}"] = torch.zeros(256, device=device)
            continue
        wdw = g(q + "2.weight").reshape(-1)
        inv = g(q + "3.weight") / torch.sqrt(g(q + "3.running_var") + 1e-5)
        out[f"RTFS_P_CAF_{s_slot}"] = wdw * inv
        out[f"RTFS_P_CAF_{t_slot}"] = g(q + "3.bias") - g(q + "3.running_mean") * inv

    out["RTFS_P_MK_A"] = g("mask_generator.mask_generator.0.weight").reshape(1)
    perm = consts["mask_perm"]
    out["RTFS_P_MK_W"] = tf32_round(g("mask_generator.mask_generator.1.full_layer.2.weight").reshape(256, 256)[perm])
    out["RTFS_P_MK_B"] = g("mask_generator.mask_generator.1.full_layer.2.bias")[perm].contiguous()
    out["RTFS_P_DEC_W"] = g("decoder.decoder.weight").permute(1, 2, 3, 0).reshape(18, 256).contiguous()
    out["RTFS_P_DEC_WT"] = torch.cat([out["RTFS_P_DEC_W"].t()[perm], torch.zeros(256, 2, device=device)], 1).contiguous()  # [interleaved col][20]
    if train:
        # W'[N = K_fwd][K = N_fwd] images for dX = dY * W
        out["RTFS_P_BN_WT"] = out["RTFS_P_BN_W"].t()
        out["RTFS_P_PJ_WT"] = out["RTFS_P_PJ_W"].t()
        out["RTFS_P_RC_WT"] = out["RTFS_P_RC_W"].t()
        out["RTFS_P_MK_WT"] = tf32_round(g("mask_generator.mask_generator.1.full_layer.2.weight").reshape(256, 256)).t()  # natural order
        for tag in ("RF", "RT"):
            for l in range(4):
                out[f"RTFS_P_{tag}_W{l}T"] = out[f"RTFS_P_{tag}_W{l}"].t()
            q = BLK + f"globalatt.{0 if tag == 'RF' else 1}."
            out[f"RTFS_P_{tag}_CTWB"] = tf32_round(g(q + "linear.weight").permute(0, 2, 1).reshape(64, 512))  # [ci][tap*64+co]
        out["RTFS_P_AT_WQKVT"] = out["RTFS_P_AT_WQKV"].t()
        out["RTFS_P_AT_WOT"] = out["RTFS_P_AT_WO"].t()
        dw = torch.zeros(256, 32, device=device)
        out["RTFS_P_DEC_WE"] = torch.cat([g("decoder.decoder.weight").permute(0, 2, 3, 1).reshape(256, 18), dw[:, 18:]], 1)  # k = (i*3+j)*2+o

    out["RTFS_P_AT_WQKVI"] = umma_image(out["RTFS_P_AT_WQKV"])
    out["RTFS_P_AT_WOI"] = umma_image(out["RTFS_P_AT_WO"])
    enc_hi = tf32_round(out["RTFS_P_ENC_W"])
    enc_lo = tf32_round(out["RTFS_P_ENC_W"] - enc_hi)
    out["RTFS_P_ENC_WI3"] = umma_image(torch.cat([enc_hi, enc_hi, enc_lo], 1))
    for src, dst in (("BN_W", "BN_WI"), ("PJ_W", "PJ_WI"), ("RC_W", "RC_WI"), ("MK_W", "MK_WI")):
        out["RTFS_P_" + dst] = umma_image(out["RTFS_P_" + src])
    for tag in ("RF", "RT"):
        for l in range(4):
            out[f"RTFS_P_{tag}_WI{l}"] = umma_image(out[f"RTFS_P_{tag}_W{l}"])
        out[f"RTFS_P_{tag}_CTWI"] = umma_image(out[f"RTFS_P_{tag}_CTW"])
        out[f"RTFS_P_{tag}_FUSED"] = dprnn_fused_image(
            out[f"RTFS_P_{tag}_W0"], [out[f"RTFS_P_{tag}_W{l}"] for l in (1, 2, 3)], out[f"RTFS_P_{tag}_CTW"])

    if not train and video_pack_supported(sd):
        out["RTFS_P_VIDEO_PACK"] = pack_video(sd, device)
    missing = [n for n in _lib.PARAM_NAMES if n not in out and not (n in TRAIN_ONLY and not train) and n != "RTFS_P_VIDEO_PACK"]
    if missing:
        raise RuntimeError(f"unprepared parameter slots: {missing}")
    return {n: (out[n].contiguous() if n in out else None) for n in _lib.PARAM_NAMES}


class PackedParams:
    """The pointer table passed as `params` to every C-ABI call (keeps the tensors alive)."""

    def __init__(self, sd, device, train=False):
        self.device = torch.device(device)
        self.tensors = prepare(sd, self.device, train)
        self.table = (ctypes.c_void_p * len(_lib.PARAM_NAMES))(
            *[(self.tensors[n].data_ptr() if self.tensors[n] is not None else None) for n in _lib.PARAM_NAMES])

    @property
    def ptr(self):
        return ctypes.cast(self.table, ctypes.c_void_p)
