"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the RTFS-Net model-forward hot path.

A from-scratch functional restatement (plain torch CPU ops, fp32 or fp64) of
`AVNet.forward` of spkgyk/RTFS-Net for the RTFS configurations
(config/lrs2_RTFSNet_{4,6,12}_layer.yaml).  It takes a *state_dict* with the reference's key
names and never imports the reference, so it travels to the GPU box.

Nothing in the product path (`rtfs_net_b200/`) may import this module: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs do, and only
as the checker / the timed CPU baseline.

Pinning: `oracle/make_golden.py` imports the reference's own `src/models` (with the shims in
`oracle/shims/`) in the authoring container and checks this restatement against it tensor by
tensor; the resulting golden vectors live in `tests/golden/`.  The SRU recurrence comes from an
un-vendored third-party package and is PARITY UNPINNED (see oracle/sru_ref.py).

Every function cites the reference file:line it follows (paths relative to
/root/reference/src/models/).
"""
import math

import torch
import torch.nn.functional as F

from .sru_ref import sru_layer_forward

EPS = 1e-5  # layers/normalizations.py:5


# ----------------------------------------------------------------------------- primitives
def gln(x, w, b):
    """GlobalLayerNorm = GroupNorm(1, C): stats over (C, spatial) per sample, biased var.
    layers/normalizations.py:8-17"""
    dims = tuple(range(1, x.ndim))
    mu = x.mean(dim=dims, keepdim=True)
    var = x.var(dim=dims, unbiased=False, keepdim=True)
    shape = (1, -1) + (1,) * (x.ndim - 2)
    return (x - mu) / torch.sqrt(var + EPS) * w.view(shape) + b.view(shape)


def ln4d(x, gamma, beta):
    """LayerNormalization4D: over dim 1 when gamma is (1,C,1,1), over dims (1,3) when (1,C,1,Q).
    layers/normalizations.py:20-37"""
    dim = (1, 3) if gamma.shape[-1] > 1 else (1,)
    mu = x.mean(dim=dim, keepdim=True)
    std = torch.sqrt(x.var(dim=dim, unbiased=False, keepdim=True) + EPS)
    return (x - mu) / std * gamma + beta


def prelu(x, a):
    """nn.PReLU() with a single shared slope."""
    return torch.where(x >= 0, x, a.view(()) * x)


BN_TRAIN = False  # True: BatchNorm layers normalise with the statistics of the batch (nn.Module.train() semantics)


def batchnorm_eval(x, w, b, rm, rv, eps=1e-5):
    shape = (1, -1) + (1,) * (x.ndim - 2)
    if BN_TRAIN:  # torch.nn.functional.batch_norm(training=True): biased variance over (batch, spatial)
        dims = (0,) + tuple(range(2, x.ndim))
        rm = x.mean(dim=dims)
        rv = x.var(dim=dims, unbiased=False)
    return (x - rm.view(shape)) / torch.sqrt(rv.view(shape) + eps) * w.view(shape) + b.view(shape)


def dwconv2d(x, w, b, stride):
    """Depthwise KxK conv as ConvNormAct builds it (layers/conv_layers.py:100-112):
    stride 1 -> padding='same' (k=4: 1 before, 2 after); stride>1 -> pad (k-1)//2."""
    k = w.shape[-1]
    C = x.shape[1]
    if stride == 1:
        tot = k - 1
        lo = tot // 2
        x = F.pad(x, (lo, tot - lo, lo, tot - lo))
        return F.conv2d(x, w, b, stride=1, groups=C)
    return F.conv2d(x, w, b, stride=stride, padding=(k - 1) // 2, groups=C)


def dwconv1d(x, w, b, stride):
    k = w.shape[-1]
    C = x.shape[1]
    if stride == 1:
        tot = k - 1
        lo = tot // 2
        x = F.pad(x, (lo, tot - lo))
        return F.conv1d(x, w, b, stride=1, groups=C)
    return F.conv1d(x, w, b, stride=stride, padding=(k - 1) // 2, groups=C)


def nearest_idx(n_out, n_in, device):
    """F.interpolate(mode='nearest') source index: min(floor(dst*in/out), in-1)."""
    return torch.clamp((torch.arange(n_out, device=device) * n_in) // n_out, max=n_in - 1)


def nearest2d(x, size):
    it = nearest_idx(size[0], x.shape[-2], x.device)
    jf = nearest_idx(size[1], x.shape[-1], x.device)
    return x[..., it, :][..., jf]


def nearest1d(x, size):
    return x[..., nearest_idx(size, x.shape[-1], x.device)]


# ----------------------------------------------------------------------------- encoder / decoder
def hann_periodic(n, dtype):
    """The reference registers `torch.hann_window(win)` as an fp32 buffer (encoder.py:159,
    decoder.py:108): the window values are fp32-rounded whatever the compute dtype."""
    return torch.hann_window(n, periodic=True, dtype=torch.float32).to(dtype)


def stft_spec(wav, win=256, hop=128):
    """torch.stft(center=True, reflect, onesided, hann periodic) restated with a real DFT.
    TDAVNet/encoder.py:161-172.  wav (B,L) -> spec (B,2,T,F) with channel 0 = Re, 1 = Im."""
    B, L = wav.shape
    x = F.pad(wav[:, None, :], (win // 2, win // 2), mode="reflect")[:, 0]
    frames = x.unfold(-1, win, hop) * hann_periodic(win, wav.dtype).to(wav.device)  # (B,T,win)
    n = torch.arange(win, dtype=torch.float64)
    f = torch.arange(win // 2 + 1, dtype=torch.float64)
    ang = 2.0 * math.pi * (n[:, None] * f[None, :] % win) / win
    cosm = torch.cos(ang).to(wav.dtype).to(wav.device)
    sinm = (-torch.sin(ang)).to(wav.dtype).to(wav.device)
    return torch.stack([frames @ cosm, frames @ sinm], 1)  # (B,2,T,F)


def encoder(sd, wav):
    """STFTEncoder.forward, TDAVNet/encoder.py:161-175 (conv 3x3 2->C, no bias/norm/act)."""
    if wav.ndim == 1:
        wav = wav[None]
    elif wav.ndim == 3:
        wav = wav[:, 0]
    spec = stft_spec(wav)
    return F.conv2d(spec, sd["encoder.conv.full_layer.2.weight"], None, padding=1)


def istft(re, im, length, win=256, hop=128):
    """torch.istft(center=True, length=L) restated: irDFT, window, overlap-add, /sum(window^2),
    trim win//2 each side, cut/pad to L.  TDAVNet/decoder.py:122-128.  re, im: (B,T,F)."""
    B, T, Fq = re.shape
    dt = re.dtype
    n = torch.arange(win, dtype=torch.float64)
    f = torch.arange(Fq, dtype=torch.float64)
    ang = 2.0 * math.pi * (f[:, None] * n[None, :] % win) / win
    wgt = torch.full((Fq, 1), 2.0, dtype=torch.float64)
    wgt[0] = 1.0
    wgt[-1] = 1.0
    cosm = (wgt * torch.cos(ang) / win).to(dt).to(re.device)
    sinm = (-wgt * torch.sin(ang) / win)
    sinm[0] = 0.0  # imaginary parts of DC / Nyquist are ignored by a C2R transform
    sinm[-1] = 0.0
    sinm = sinm.to(dt).to(re.device)
    w = hann_periodic(win, dt).to(re.device)
    frames = (re @ cosm + im @ sinm) * w  # (B,T,win)
    n_out = win + hop * (T - 1)
    y = re.new_zeros(B, n_out)
    env = re.new_zeros(n_out)
    for t in range(T):
        y[:, t * hop : t * hop + win] += frames[:, t]
        env[t * hop : t * hop + win] += w * w
    y = y[:, win // 2 :]
    env = env[win // 2 :]
    end = min(length, y.shape[1])
    out = re.new_zeros(B, length)
    out[:, :end] = y[:, :end] / env[:end]
    return out


def decoder(sd, z, length):
    """STFTDecoder.forward, TDAVNet/decoder.py:110-132.  z (B,1,C,T,F) -> (B,1,L)."""
    B = z.shape[0]
    z = z.reshape(B, z.shape[-3], z.shape[-2], z.shape[-1])
    w = sd["decoder.decoder.weight"]  # (C,2,3,3) ConvTranspose2d weight
    y = F.conv2d(z, w.transpose(0, 1).flip(-1, -2), None, padding=1)
    return istft(y[:, 0], y[:, 1], length)[:, None, :]


# ----------------------------------------------------------------------------- RTFS block parts
def dual_path_rnn(sd, p, z, dim):
    """DualPathRNN.forward, layers/rnn_layers.py:136-162 (kernel 8, stride 1, SRU 4 layers bidir)."""
    if dim == 4:
        z = z.transpose(-2, -1).contiguous()
    B, C, S, O = z.shape
    ks = 8
    assert S >= ks
    n = ln4d(z, sd[p + "norm.gamma"], sd[p + "norm.beta"])
    seq = n.permute(0, 3, 1, 2).reshape(B * O, C, S)  # (B*O, C, S)
    X = seq.unfold(2, ks, 1)  # (B*O, C, L, ks)
    L = X.shape[2]
    X = X.permute(2, 0, 1, 3).reshape(L, B * O, C * ks)  # channel-major, tap-minor (nn.Unfold)
    i = 0
    while (p + f"rnn.rnn_lst.{i}.weight") in sd:
        q = p + f"rnn.rnn_lst.{i}."
        hid = sd[q + "weight_c"].shape[0] // 4
        X, _ = sru_layer_forward(X, sd[q + "weight"], sd[q + "weight_c"], sd[q + "bias"], hid, True)
        i += 1
    Y = X.permute(1, 2, 0)  # (B*O, 64, L)
    zt = F.conv_transpose1d(Y, sd[p + "linear.weight"], sd[p + "linear.bias"])  # (B*O, C, S)
    out = zt.view(B, O, C, S).permute(0, 2, 3, 1) + z
    if dim == 4:
        out = out.transpose(-2, -1).contiguous()
    return out


def conv_act_norm(sd, p, x):
    """ConvActNorm (1x1 conv -> PReLU -> LN4D), layers/conv_layers.py:201-205."""
    y = F.conv2d(x, sd[p + "conv.weight"], sd[p + "conv.bias"])
    y = prelu(y, sd[p + "act.weight"])
    return ln4d(y, sd[p + "norm.gamma"], sd[p + "norm.beta"])


def mhsa2d(sd, p, x):
    """MultiHeadSelfAttention2D.forward, layers/attention.py:149-189."""
    B, C, T, Fq = x.shape
    n_head = 0
    while (p + f"Queries.{n_head}.conv.weight") in sd:
        n_head += 1
    outs = []
    for h in range(n_head):
        Q = conv_act_norm(sd, p + f"Queries.{h}.", x).transpose(1, 2).flatten(2)  # (B,T,E*F)
        K = conv_act_norm(sd, p + f"Keys.{h}.", x).transpose(1, 2).flatten(2)
        V = conv_act_norm(sd, p + f"Values.{h}.", x).transpose(1, 2)  # (B,T,Cv,F)
        shp = V.shape
        att = torch.softmax(Q @ K.transpose(1, 2) / math.sqrt(Q.shape[-1]), dim=2)
        outs.append((att @ V.flatten(2)).reshape(shp).transpose(1, 2))  # (B,Cv,T,F)
    y = torch.cat(outs, 1)  # head-major channels (attention.py:178-181)
    return conv_act_norm(sd, p + "attn_concat_proj.", y) + x


def tfar(sd, p, local, glob, taps=None, tag=""):
    """InjectionMultiSum.forward (TF-AR unit), layers/fusion.py:54-69.  1-D or 2-D by rank.
    `taps` (test instrumentation only) receives the pre-norm conv outputs."""
    two_d = local.ndim == 4
    dw = dwconv2d if two_d else dwconv1d
    up = nearest2d if two_d else nearest1d
    norm_p = lambda q, t: _norm(sd, q + "full_layer.3.", t)
    loc_shape = tuple(local.shape[2:]) if two_d else local.shape[-1]
    n_loc = math.prod(local.shape[2:])
    n_glob = math.prod(glob.shape[2:])
    lw = sd[p + "local_embedding.full_layer.2.weight"]
    l_pre = dw(local, lw, None, 1)
    local_emb = norm_p(p + "local_embedding.", l_pre)
    gi = glob if n_loc > n_glob else up(glob, loc_shape)
    e_pre = dw(gi, sd[p + "global_embedding.full_layer.2.weight"], None, 1)
    g_pre = dw(gi, sd[p + "global_gate.full_layer.2.weight"], None, 1)
    if n_loc > n_glob:
        ge = up(norm_p(p + "global_embedding.", e_pre), loc_shape)
        gate = up(torch.sigmoid(norm_p(p + "global_gate.", g_pre)), loc_shape)
    else:
        ge = norm_p(p + "global_embedding.", e_pre)
        gate = torch.sigmoid(norm_p(p + "global_gate.", g_pre))
    if taps is not None:
        taps.update({tag + "_l_pre": l_pre, tag + "_e_pre": e_pre, tag + "_g_pre": g_pre})
    return local_emb * gate + ge


def _norm(sd, q, x):
    """Norm stored at `q` = '...full_layer.N.': gLN (key q+'norm.weight') or eval-mode BatchNorm
    (keys q+'weight', q+'running_mean'); identity if absent.  layers/normalizations.py:44-58."""
    if (q + "norm.weight") in sd:
        return gln(x, sd[q + "norm.weight"], sd[q + "norm.bias"])
    if (q + "running_mean") in sd:
        return batchnorm_eval(x, sd[q + "weight"], sd[q + "bias"], sd[q + "running_mean"], sd[q + "running_var"])
    return x


def rtfs_block(sd, p, x, taps=None):
    """TDANetBlock.forward with is2d=True, upsampling_depth=2, layers = [DPRNN(4), DPRNN(3), MHSA2D].
    separators/tdanet.py:106-133."""
    r = prelu(x * sd[p + "gateway.full_layer.2.weight"].view(1, -1, 1, 1) + sd[p + "gateway.full_layer.2.bias"].view(1, -1, 1, 1),
              sd[p + "gateway.full_layer.4.weight"])
    p_pre = F.conv2d(r, sd[p + "projection.full_layer.2.weight"], sd[p + "projection.full_layer.2.bias"])
    pp = prelu(gln(p_pre, sd[p + "projection.full_layer.3.norm.weight"], sd[p + "projection.full_layer.3.norm.bias"]),
               sd[p + "projection.full_layer.4.weight"])
    q = p + "downsample_layers.0.full_layer."
    d0_pre = dwconv2d(pp, sd[q + "2.weight"], sd[q + "2.bias"], 1)
    d0 = gln(d0_pre, sd[q + "3.norm.weight"], sd[q + "3.norm.bias"])
    q = p + "downsample_layers.1.full_layer."
    d1_pre = dwconv2d(d0, sd[q + "2.weight"], sd[q + "2.bias"], 2)
    d1 = gln(d1_pre, sd[q + "3.norm.weight"], sd[q + "3.norm.bias"])
    pool = F.adaptive_avg_pool2d(d0, d1.shape[-2:])
    g = pool + d1  # tdanet.py:117-118 (pool of d1 to its own size = id)
    g0 = g
    g = dual_path_rnn(sd, p + "globalatt.0.", g, 4)
    g1 = g
    g = dual_path_rnn(sd, p + "globalatt.1.", g, 3)
    g2 = g
    g = mhsa2d(sd, p + "globalatt.2.", g)
    f0 = tfar(sd, p + "fusion_layers.0.", d0, g, taps, "f0")
    f1 = tfar(sd, p + "fusion_layers.1.", d1, g, taps, "f1")
    e = tfar(sd, p + "concat_layers.0.", f0, f1, taps, "c0") + d0
    out = F.conv2d(e, sd[p + "residual_conv.full_layer.2.weight"], sd[p + "residual_conv.full_layer.2.bias"]) + r
    if taps is not None:
        taps.update(dict(d0=d0, d1=d1, g0=g0, g1=g1, g2=g2, g3=g, f0=f0, f1=f1, e=e,
                         p_pre=p_pre, d0_pre=d0_pre, d1_pre=d1_pre, pool=pool))
    return out


# ----------------------------------------------------------------------------- video block (1-D)
def positional_encoding(n_pos, channels, dtype, max_len=10000):
    """layers/attention.py:9-25"""
    pe = torch.zeros(n_pos, channels)
    position = torch.arange(0, n_pos).unsqueeze(1).float()
    div_term = torch.exp(torch.arange(0, channels, 2).float() * -(torch.log(torch.tensor(max_len).float()) / channels))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.to(dtype)


def video_mhsa(sd, p, x, n_head=8):
    """MultiHeadSelfAttention.forward (eval mode), layers/attention.py:57-73.  x (B,C,T)."""
    res = x
    y = x.transpose(1, 2)
    C = y.shape[-1]
    y = F.layer_norm(y, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    pe = sd[p + "pos_enc.pe"][:, : y.shape[1]] if (p + "pos_enc.pe") in sd else positional_encoding(y.shape[1], C, y.dtype)[None]
    y = y + pe
    residual = y
    qkv = y @ sd[p + "attention.in_proj_weight"].T + sd[p + "attention.in_proj_bias"]
    q, k, v = qkv.chunk(3, dim=-1)
    B, T, _ = q.shape
    hd = C // n_head
    q = q.view(B, T, n_head, hd).transpose(1, 2)
    k = k.view(B, T, n_head, hd).transpose(1, 2)
    v = v.view(B, T, n_head, hd).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, T, C)
    o = o @ sd[p + "attention.out_proj.weight"].T + sd[p + "attention.out_proj.bias"]
    y = o + residual
    y = F.layer_norm(y, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    return y.transpose(1, 2) + res


def video_ffn(sd, p, x):
    """FeedForwardNetwork.forward (eval mode), layers/conv_layers.py:252-259."""
    y = F.conv1d(x, sd[p + "encoder.full_layer.2.weight"])
    y = gln(y, sd[p + "encoder.full_layer.3.norm.weight"], sd[p + "encoder.full_layer.3.norm.bias"])
    y = torch.relu(dwconv1d(y, sd[p + "refiner.full_layer.2.weight"], sd[p + "refiner.full_layer.2.bias"], 1))
    y = F.conv1d(y, sd[p + "decoder.full_layer.2.weight"])
    y = gln(y, sd[p + "decoder.full_layer.3.norm.weight"], sd[p + "decoder.full_layer.3.norm.bias"])
    return y + x


def video_block(sd, p, x, stride=2):
    """TDANetBlock.forward with is2d=False (VP block), separators/tdanet.py:106-133; eval mode."""
    r = prelu(x * sd[p + "gateway.full_layer.2.weight"].view(1, -1, 1) + sd[p + "gateway.full_layer.2.bias"].view(1, -1, 1),
              sd[p + "gateway.full_layer.4.weight"])
    y = F.conv1d(r, sd[p + "projection.full_layer.2.weight"], sd[p + "projection.full_layer.2.bias"])
    y = prelu(_norm(sd, p + "projection.full_layer.3.", y), sd[p + "projection.full_layer.4.weight"])
    depth = 0
    while (p + f"downsample_layers.{depth}.full_layer.2.weight") in sd:
        depth += 1
    ds = []
    cur = y
    for i in range(depth):
        q = p + f"downsample_layers.{i}.full_layer."
        cur = _norm(sd, q + "3.", dwconv1d(cur, sd[q + "2.weight"], sd[q + "2.bias"], 1 if i == 0 else stride))
        ds.append(cur)
    size = ds[-1].shape[-1]
    g = sum(F.adaptive_avg_pool1d(d, size) for d in ds)
    g = video_mhsa(sd, p + "globalatt.0.MHSA.", g)
    g = video_ffn(sd, p + "globalatt.0.FFN.", g)
    fused = [tfar(sd, p + f"fusion_layers.{i}.", ds[i], g) for i in range(depth)]
    expanded = tfar(sd, p + f"concat_layers.{depth - 2}.", fused[-2], fused[-1]) + ds[-2]
    for i in range(depth - 3, -1, -1):
        expanded = tfar(sd, p + f"concat_layers.{i}.", fused[i], expanded) + ds[i]
    return F.conv1d(expanded, sd[p + "residual_conv.full_layer.2.weight"], sd[p + "residual_conv.full_layer.2.bias"]) + r


# ----------------------------------------------------------------------------- CAF, S3 mask
def caf(sd, p, a, v):
    """ATTNFusionCell.forward (CAF), layers/fusion.py:252-274; eval-mode BatchNorm2d; is2d=True.
    a (B,Ca,T,F), v (B,Cv,Tv) -> (B,Ca,T,F)."""
    B, Ca, T, _ = a.shape
    groups = Ca
    vr = F.conv1d(v, sd[p + "resize.full_layer.2.weight"], sd[p + "resize.full_layer.2.bias"], groups=groups)
    vr = gln(vr, sd[p + "resize.full_layer.3.norm.weight"], sd[p + "resize.full_layer.3.norm.bias"])
    vk = nearest1d(vr, T)[..., None]
    k1 = torch.relu(_norm(sd, p + "key_embed.full_layer.3.", a * sd[p + "key_embed.full_layer.2.weight"].view(1, -1, 1, 1))) * vk
    val = _norm(sd, p + "value_embed.full_layer.3.", a * sd[p + "value_embed.full_layer.2.weight"].view(1, -1, 1, 1))
    att = F.conv1d(v, sd[p + "attention_embed.full_layer.2.weight"], sd[p + "attention_embed.full_layer.2.bias"], groups=groups)
    att = gln(att, sd[p + "attention_embed.full_layer.3.norm.weight"], sd[p + "attention_embed.full_layer.3.norm.bias"])
    ksz = att.shape[1] // Ca
    att = att.reshape(B, Ca, ksz, -1).mean(2)
    att = nearest1d(torch.softmax(att, -1), T)[..., None]
    return k1 + att * val


def s3_mask(sd, refined, a0):
    """MaskGenerator.forward with RI_split, n_src=1.  TDAVNet/mask_generator.py:67-99."""
    m = prelu(refined, sd["mask_generator.mask_generator.0.weight"])
    m = torch.relu(F.conv2d(m, sd["mask_generator.mask_generator.1.full_layer.2.weight"], sd["mask_generator.mask_generator.1.full_layer.2.bias"]))
    C = a0.shape[1]
    mr, mi = m[:, : C // 2], m[:, C // 2 :]
    er, ei = a0[:, : C // 2], a0[:, C // 2 :]
    z = torch.cat([er * mr - ei * mi, er * mi + ei * mr], 1)
    return z[:, None]  # (B, n_src=1, C, T, F)


def audio_bottleneck(sd, a0):
    """ConvNormAct(pre_norm gLN, pre_act ReLU, conv 1x1 + bias), tdavnet.py:59; config:15-20."""
    y = torch.relu(gln(a0, sd["audio_bottleneck.full_layer.0.norm.weight"], sd["audio_bottleneck.full_layer.0.norm.bias"]))
    return F.conv2d(y, sd["audio_bottleneck.full_layer.2.weight"], sd["audio_bottleneck.full_layer.2.bias"])


# ----------------------------------------------------------------------------- full forward
def avnet_forward(sd, audio_mixture, mouth_embedding, repeats, taps=None):
    """AVNet.forward (tdavnet.py:86-97) + RefinementModule.forward (TDAVNet/refinement_module.py:45-62)
    for fusion_repeats = 1 (video_params.repeats: 1), shared audio block, `repeats` audio passes."""
    wav = audio_mixture
    if wav.ndim == 1:
        wav = wav[None]
    elif wav.ndim == 3:
        wav = wav[:, 0]
    L = wav.shape[-1]
    a0 = encoder(sd, wav)
    a1 = audio_bottleneck(sd, a0)
    v0 = mouth_embedding  # video_bottleneck: kernel_size -1 -> identity (conv_layers.py:87,118-124)
    ap = "refinement_module.audio_net.blocks."
    A = rtfs_block(sd, ap, a1)
    blk0 = A
    V = video_block(sd, "refinement_module.video_net.blocks.", v0)
    A = caf(sd, "refinement_module.crossmodal_fusion.fusion_module.audio_lstm.", A, V)
    caf_out = A
    for _ in range(1, repeats):
        A = rtfs_block(sd, ap, A + a1)
    z = s3_mask(sd, A, a0)
    out = decoder(sd, z, L)
    if taps is not None:
        taps.update(dict(a0=a0, a1=a1, blk0=blk0, video=V, caf=caf_out, refined=A, masked=z))
    return out


def neg_sisdr(est, target, eps=1e-8):
    """PairwiseNegSDR('sisdr') for n_src = 1 (zero-mean), /root/reference/src/losses/matrix.py:13-53.
    est, target (B,1,L) -> (B,) negative SI-SDR in dB."""
    t = target - target.mean(dim=-1, keepdim=True)
    e = est - est.mean(dim=-1, keepdim=True)
    dot = (e * t).sum(-1, keepdim=True)
    energy = (t**2).sum(-1, keepdim=True) + eps
    proj = dot * t / energy
    noise = e - proj
    ratio = (proj**2).sum(-1) / ((noise**2).sum(-1) + eps)
    return (-10.0 * torch.log10(ratio + eps))[:, 0]


def neg_snr(est, target, eps=1e-8):
    """PairwiseNegSDR('snr') for n_src = 1 (zero-mean, noise = est - target), /root/reference/src/losses/matrix.py:22-53;
    the training loss of the reference (train.py:99).  est, target (B,1,L) -> (B,) negative SNR in dB."""
    t = target - target.mean(dim=-1, keepdim=True)
    e = est - est.mean(dim=-1, keepdim=True)
    noise = e - t
    ratio = (t**2).sum(-1) / ((noise**2).sum(-1) + eps)
    return (-10.0 * torch.log10(ratio + eps))[:, 0]
