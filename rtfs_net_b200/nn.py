"""Host-side mirror of the reference's `src/models` interface for the RTFS-Net configurations.

Same class names, constructor keywords, `forward` signatures and `state_dict` keys as
/root/reference/src/models (AVNet: tdavnet.py:14-108; BaseAVModel: TDAVNet/base_av_model.py:9-118),
so `AVNet(**conf["audionet"])`, `load_state_dict(..., strict=True)` and `from_pretrain` behave as
in the reference -- but the modules are *parameter holders*: the arithmetic of the audio path runs
in the hand-written CUDA kernels of librtfs_b200.so (no eager/CPU fallback; calling the audio
path without the library, on a CPU tensor, or with autograd enabled raises).

Physical layout: tensors crossing module boundaries are logical (B,C,T,F) views of channels-last
storage (B,T,F,C); inputs in any other stride order are converted once at the boundary.

Only the RTFS-Net family (STFTEncoder/STFTDecoder, 2-D TDANet with [DualPathRNN(dim 4),
DualPathRNN(dim 3), MultiHeadSelfAttention2D], ATTNFusion, MaskGenerator with RI_split) is
accelerated; other reference configurations raise NotImplementedError at construction.
"""
import math
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .weights import PackedParams

EPS = 1e-5


# =============================================================================== parameter holders
class GlobalLayerNorm(nn.Module):
    """layers/normalizations.py:8-17 (GroupNorm(1, C))."""

    def __init__(self, num_channels=1, eps=EPS):
        super().__init__()
        self.norm = nn.GroupNorm(1, num_channels, eps=eps)

    def forward(self, x):
        return self.norm(x)


class LayerNormalization4D(nn.Module):
    """layers/normalizations.py:20-37; parameters (1,C,1,Q)."""

    def __init__(self, input_dimension, eps=EPS):
        super().__init__()
        c, q = input_dimension
        self.gamma = nn.Parameter(torch.ones(1, c, 1, q))
        self.beta = nn.Parameter(torch.zeros(1, c, 1, q))
        self.eps = eps


gLN = GlobalLayerNorm


def _norm_cls(name):
    """normalizations.get (layers/normalizations.py:44-58)."""
    from .generic import normalizations_get

    return normalizations_get(name)


def _act_cls(name):
    """activations.get (layers/activations.py:4-18)."""
    from .generic import activations_get

    return activations_get(name)


class DropPath(nn.Module):
    """timm DropPath (stochastic depth); identity in eval."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


class ConvNormAct(nn.Module):
    """pre_norm -> pre_act -> conv -> norm -> act (layers/conv_layers.py:65-129).

    1-D instances (video block, launch-bound and < 0.1 % of the bytes) execute their torch layers;
    2-D instances are parameter holders whose arithmetic is fused into the CUDA kernels."""

    def __init__(self, in_chan=1, out_chan=1, kernel_size=-1, stride=1, groups=1, dilation=1, padding=None,
                 pre_norm_type=None, pre_act_type=None, norm_type=None, act_type=None, xavier_init=False,
                 bias=True, is2d=False, *args, **kwargs):
        super().__init__()
        self.in_chan = in_chan
        self.out_chan = out_chan if kernel_size > 0 else in_chan
        self.kernel_size = kernel_size
        self.stride = stride
        self.groups = groups
        self.is2d = is2d
        if padding is None:
            padding = dilation * (kernel_size - 1) // 2 if stride > 1 else "same"
        if kernel_size > 0:
            conv = (nn.Conv2d if is2d else nn.Conv1d)(in_chan, self.out_chan, kernel_size, stride=stride, padding=padding,
                                                       dilation=dilation, groups=groups, bias=bias)
            if xavier_init:
                nn.init.xavier_uniform_(conv.weight)
        else:
            conv = nn.Identity()
        self.full_layer = nn.Sequential(_norm_cls(pre_norm_type)(in_chan), _act_cls(pre_act_type)(), conv,
                                        _norm_cls(norm_type)(self.out_chan), _act_cls(act_type)())
        self._fused = None  # set by AVNet for layers that are entry points of a fused kernel chain

    def forward(self, x):
        if self._fused is not None:
            return self._fused(x)
        if self.kernel_size <= 0 and all(isinstance(m, nn.Identity) for m in self.full_layer):
            return x
        if self.is2d:
            raise NotImplementedError("2-D ConvNormAct layers are fused into the RTFS block kernels; call the enclosing block")
        return self.full_layer(x)


class ConvActNorm(nn.Module):
    """conv -> act -> norm holder (layers/conv_layers.py:142-205)."""

    def __init__(self, in_chan=1, out_chan=1, kernel_size=-1, act_type=None, norm_type=None, n_freqs=-1, is2d=False, bias=True, *args, **kwargs):
        super().__init__()
        self.conv = (nn.Conv2d if is2d else nn.Conv1d)(in_chan, out_chan, kernel_size, bias=bias)
        self.act = _act_cls(act_type)()
        self.norm = LayerNormalization4D((out_chan, n_freqs)) if norm_type == "LayerNormalization4D" else _norm_cls(norm_type)(out_chan)


class SRUCell(nn.Module):
    """Parameter holder with the layout of `sru.SRUCell`."""

    def __init__(self, input_size, hidden_size, bidirectional):
        super().__init__()
        d = hidden_size * (2 if bidirectional else 1)
        k = 3 if input_size == d else 4
        self.weight = nn.Parameter(torch.empty(input_size, d * k))
        self.weight_c = nn.Parameter(torch.empty(2 * d))
        self.bias = nn.Parameter(torch.zeros(2 * d))
        val = math.sqrt(3.0 / input_size)
        nn.init.uniform_(self.weight, -val, val)
        nn.init.uniform_(self.weight_c, -math.sqrt(3.0), math.sqrt(3.0))
        with torch.no_grad():
            self.weight_c.mul_(math.sqrt(0.5))
        # upstream `sru` cells carry a `scale_x` buffer (only read when rescale=True, which SRU(...) does not enable:
        # SURVEY.md App. B/C, from memory -- the package is not vendored).  It is registered so that checkpoints written
        # here load strictly into the real package, and it is optional on load so that checkpoints without it load too.
        self.register_buffer("scale_x", torch.zeros(1))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        key = prefix + "scale_x"
        if key not in state_dict:
            state_dict = dict(state_dict)
            state_dict[key] = self.scale_x.detach().clone()
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


class SRU(nn.Module):
    def __init__(self, input_size, hidden_size, num_layers=2, bidirectional=False, **kwargs):
        super().__init__()
        d = hidden_size * (2 if bidirectional else 1)
        self.rnn_lst = nn.ModuleList([SRUCell(input_size if i == 0 else d, hidden_size, bidirectional) for i in range(num_layers)])


class DualPathRNN(nn.Module):
    """layers/rnn_layers.py:62-162."""

    def __init__(self, in_chan, hid_chan, dim, kernel_size=8, stride=1, rnn_type="LSTM", num_layers=1,
                 norm_type="LayerNormalization4D", act_type="Tanh", bidirectional=True, apply_ffn=False, *args, **kwargs):
        super().__init__()
        if not (rnn_type == "SRU" and kernel_size == 8 and stride == 1 and num_layers == 4 and bidirectional
                and in_chan == 64 and hid_chan == 32 and norm_type == "LayerNormalization4D" and not apply_ffn and dim in (3, 4)):
            raise NotImplementedError("DualPathRNN kernels are built for the RTFS-Net configuration (SRU, 4 layers, bidirectional, k=8)")
        self.dim = dim
        self.norm = LayerNormalization4D((in_chan, 1))
        self.rnn = SRU(in_chan * kernel_size, hid_chan, num_layers=num_layers, bidirectional=True)
        self.linear = nn.ConvTranspose1d(hid_chan * 2, in_chan, kernel_size, stride=stride)
        self._rt = None

    def forward(self, x):
        return self._rt().dprnn(x, 0 if self.dim == 4 else 1)


class MultiHeadSelfAttention2D(nn.Module):
    """layers/attention.py:76-189."""

    def __init__(self, in_chan, n_freqs, n_head=4, hid_chan=4, act_type="PReLU", norm_type="LayerNormalization4D", dim=3, *args, **kwargs):
        super().__init__()
        if not (in_chan == 64 and n_freqs == 64 and n_head == 4 and hid_chan == 4 and act_type == "PReLU" and norm_type == "LayerNormalization4D" and dim == 3):
            raise NotImplementedError("MultiHeadSelfAttention2D kernels are built for the RTFS-Net configuration")
        mk = lambda oc: ConvActNorm(in_chan, oc, 1, act_type=act_type, norm_type=norm_type, n_freqs=n_freqs, is2d=True)
        self.Queries, self.Keys, self.Values = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for _ in range(n_head):
            self.Queries.append(mk(hid_chan))
            self.Keys.append(mk(hid_chan))
            self.Values.append(mk(in_chan // n_head))
        self.attn_concat_proj = mk(in_chan)
        self._rt = None

    def forward(self, x):
        return self._rt().mhsa(x)


class PositionalEncoding(nn.Module):
    """layers/attention.py:9-25."""

    def __init__(self, channels, max_len=10000):
        super().__init__()
        pe = torch.zeros(max_len, channels)
        position = torch.arange(0, max_len).unsqueeze(1).float()
        div_term = torch.exp(torch.arange(0, channels, 2).float() * -(torch.log(torch.tensor(max_len).float()) / channels))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))

    def forward(self, x):
        return x + self.pe[:, : x.size(1)]


class MultiHeadSelfAttention(nn.Module):
    """Video-block attention over <= 7 frames (layers/attention.py:28-73); torch library ops."""

    def __init__(self, in_chan, n_head=8, dropout=0.1, positional_encoding=True, batch_first=True, *args, **kwargs):
        super().__init__()
        assert in_chan % n_head == 0
        self.norm1 = nn.LayerNorm(in_chan)
        self.pos_enc = PositionalEncoding(in_chan) if positional_encoding else nn.Identity()
        self.attention = nn.MultiheadAttention(in_chan, n_head, dropout, batch_first=True)
        self.dropout_layer = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(in_chan)
        self.drop_path_layer = DropPath(dropout)

    def forward(self, x):
        y = self.pos_enc(self.norm1(x.transpose(1, 2)))
        y = self.norm2(self.dropout_layer(self.attention(y, y, y, need_weights=False)[0]) + y)
        return self.drop_path_layer(y.transpose(1, 2)) + x


class FeedForwardNetwork(nn.Module):
    """layers/conv_layers.py:218-259 (1-D use in the video block)."""

    def __init__(self, in_chan, hid_chan, kernel_size=5, norm_type="gLN", act_type="ReLU", dropout=0, is2d=False, *args, **kwargs):
        super().__init__()
        self.encoder = ConvNormAct(in_chan, hid_chan, 1, norm_type=norm_type, bias=False, is2d=is2d)
        self.refiner = ConvNormAct(hid_chan, hid_chan, kernel_size, groups=hid_chan, act_type=act_type, is2d=is2d)
        self.decoder = ConvNormAct(hid_chan, in_chan, 1, norm_type=norm_type, bias=False, is2d=is2d)
        self.dropout_layer = DropPath(dropout)

    def forward(self, x):
        y = self.dropout_layer(self.refiner(self.encoder(x)))
        return self.dropout_layer(self.decoder(y)) + x


class GlobalAttention(nn.Module):
    """layers/attention.py:192-220."""

    def __init__(self, in_chan, hid_chan=None, ffn_name="FeedForwardNetwork", kernel_size=5, n_head=8, dropout=0.1, pos_enc=True, *args, **kwargs):
        super().__init__()
        hid_chan = hid_chan if hid_chan is not None else 2 * in_chan
        if ffn_name != "FeedForwardNetwork":
            raise NotImplementedError(ffn_name)
        self.MHSA = MultiHeadSelfAttention(in_chan, n_head, dropout, pos_enc)
        self.FFN = FeedForwardNetwork(in_chan, hid_chan, kernel_size, dropout=dropout)

    def forward(self, x):
        return self.FFN(self.MHSA(x))


class InjectionMultiSum(nn.Module):
    """TF-AR unit (layers/fusion.py:9-69).  2-D: holder (fused); 1-D: torch ops (video block)."""

    def __init__(self, in_chan, kernel_size, norm_type="gLN", is2d=False, *args, **kwargs):
        super().__init__()
        mk = lambda act: ConvNormAct(in_chan, in_chan, kernel_size, groups=in_chan, norm_type=norm_type, act_type=act, bias=False, is2d=is2d)
        self.local_embedding = mk(None)
        self.global_embedding = mk(None)
        self.global_gate = mk("Sigmoid")
        self.is2d = is2d

    def forward(self, local_features, global_features):
        if self.is2d:
            raise NotImplementedError("2-D TF-AR units are fused into the RTFS block kernels; call the enclosing block")
        size = local_features.shape[-1]
        local_emb = self.local_embedding(local_features)
        if local_features.shape[-1] > global_features.shape[-1]:
            g_emb = F.interpolate(self.global_embedding(global_features), size=size, mode="nearest")
            gate = F.interpolate(self.global_gate(global_features), size=size, mode="nearest")
        else:
            gi = F.interpolate(global_features, size=size, mode="nearest")
            g_emb, gate = self.global_embedding(gi), self.global_gate(gi)
        return local_emb * gate + g_emb


class TDANetBlock(nn.Module):
    """RTFS block / VP block (separators/tdanet.py:8-133)."""

    def __init__(self, in_chan, hid_chan, kernel_size=5, stride=2, norm_type="gLN", act_type="PReLU", upsampling_depth=4, layers=dict(), is2d=False):
        super().__init__()
        self.in_chan, self.hid_chan, self.upsampling_depth, self.is2d = in_chan, hid_chan, upsampling_depth, is2d
        self.gateway = ConvNormAct(in_chan, in_chan, 1, groups=in_chan, act_type=act_type, is2d=is2d)
        self.projection = ConvNormAct(in_chan, hid_chan, 1, norm_type=norm_type, act_type=act_type, is2d=is2d)
        self.downsample_layers = nn.ModuleList(
            [ConvNormAct(hid_chan, hid_chan, kernel_size, stride=1 if i == 0 else stride, groups=hid_chan, norm_type=norm_type, is2d=is2d)
             for i in range(upsampling_depth)])
        from .generic import layers_get

        mods = [layers_get(layer["layer_type"])(in_chan=hid_chan, **layer) for _, layer in layers.items()]
        self.globalatt = nn.Sequential(*mods)
        self.fusion_layers = nn.ModuleList([InjectionMultiSum(hid_chan, kernel_size, norm_type, is2d) for _ in range(upsampling_depth)])
        self.concat_layers = nn.ModuleList([InjectionMultiSum(hid_chan, kernel_size, norm_type, is2d) for _ in range(upsampling_depth - 1)])
        self.residual_conv = ConvNormAct(hid_chan, in_chan, 1, is2d=is2d)
        if is2d:
            ok = (in_chan == 256 and hid_chan == 64 and kernel_size == 4 and stride == 2 and norm_type == "gLN" and act_type == "PReLU"
                  and upsampling_depth == 2 and [type(m) for m in self.globalatt] == [DualPathRNN, DualPathRNN, MultiHeadSelfAttention2D]
                  and self.globalatt[0].dim == 4 and self.globalatt[1].dim == 3)
            if not ok:
                raise NotImplementedError("the 2-D block kernels are built for the RTFS-Net configuration (256/64 channels, k=4, depth 2)")
        self._rt = None

    def forward(self, x):
        if self.is2d:
            return self._rt().block(x)
        # 1-D VP block: (B,512,Tv) with Tv ~ 50 -- launch-bound torch library ops
        residual = self.gateway(x)
        ds = [self.downsample_layers[0](self.projection(residual))]
        for i in range(1, self.upsampling_depth):
            ds.append(self.downsample_layers[i](ds[-1]))
        g = sum(F.adaptive_avg_pool1d(d, ds[-1].shape[-1]) for d in ds)
        g = self.globalatt(g)
        fused = [self.fusion_layers[i](ds[i], g) for i in range(self.upsampling_depth)]
        expanded = self.concat_layers[-1](fused[-2], fused[-1]) + ds[-2]
        for i in range(self.upsampling_depth - 3, -1, -1):
            expanded = self.concat_layers[i](fused[i], expanded) + ds[i]
        return self.residual_conv(expanded) + residual


class TDANet(nn.Module):
    """separators/tdanet.py:136-211 (shared-block form only, as in every RTFS config)."""

    def __init__(self, in_chan=-1, hid_chan=-1, kernel_size=5, stride=2, norm_type="gLN", act_type="PReLU", upsampling_depth=4,
                 layers=dict(), repeats=4, shared=False, is2d=False, *args, **kwargs):
        super().__init__()
        if not shared and is2d:
            raise NotImplementedError("the 2-D block kernels serve the RTFS-Net configurations, which share one block across repeats (shared: true)")
        self.repeats, self.shared, self.is2d = repeats, shared, is2d
        mk = lambda: TDANetBlock(in_chan, hid_chan, kernel_size, stride, norm_type, act_type, upsampling_depth, layers, is2d)
        self.blocks = mk() if shared else nn.ModuleList(mk() for _ in range(repeats))  # separators/tdanet.py:168-205

    def get_block(self, i):
        return self.blocks if self.shared else self.blocks[i]

    def forward(self, x):
        residual = x
        for i in range(self.repeats):
            x = self.get_block(i)((x + residual) if i > 0 else x)
        return x


class ATTNFusionCell(nn.Module):
    """CAF cell (layers/fusion.py:194-274)."""

    def __init__(self, in_chan_a, in_chan_b, kernel_size=1, is2d=False, *args, **kwargs):
        super().__init__()
        if not (in_chan_a == 256 and in_chan_b == 512 and kernel_size == 4 and is2d):
            raise NotImplementedError("the CAF kernels are built for 256 audio / 512 video channels, 4 heads")
        self.key_embed = ConvNormAct(in_chan_a, in_chan_a, 1, groups=in_chan_a, norm_type="BatchNorm2d", act_type="ReLU", bias=False, is2d=True)
        self.value_embed = ConvNormAct(in_chan_a, in_chan_a, 1, groups=in_chan_a, norm_type="BatchNorm2d", bias=False, is2d=True)
        self.attention_embed = ConvNormAct(in_chan_b, kernel_size * in_chan_a, 1, groups=in_chan_a, norm_type="gLN")
        self.resize = ConvNormAct(in_chan_b, in_chan_a, 1, groups=in_chan_a, norm_type="gLN")
        self._rt = None

    def forward(self, audio, video):
        return self._rt().caf(audio, video)


class ATTNFusion(nn.Module):
    """TDAVNet/fusion.py:187-212 (video_fusion is False whenever fusion_repeats <= 1)."""

    def __init__(self, ain_chan, vin_chan, kernel_size, video_fusion=True, is2d=True, *args, **kwargs):
        super().__init__()
        if video_fusion:
            raise NotImplementedError("video-side fusion (fusion_repeats > 1) is outside the RTFS-Net configurations")
        self.video_fusion = False
        self.audio_lstm = ATTNFusionCell(ain_chan, vin_chan, kernel_size, is2d)

    def forward(self, audio, video):
        return self.audio_lstm(audio, video), video


class MultiModalFusion(nn.Module):
    """TDAVNet/fusion.py:215-281."""

    def __init__(self, audio_bn_chan, video_bn_chan, kernel_size=1, fusion_repeats=3, fusion_type="ConcatFusion", fusion_shared=False, is2d=False, **kwargs):
        super().__init__()
        from .generic import fusion_get

        self.fusion_repeats, self.fusion_type, self.fusion_shared, self.is2d = fusion_repeats, fusion_type, fusion_shared, is2d
        cls = fusion_get(fusion_type) if fusion_repeats > 0 else nn.Identity
        if cls is ATTNFusion and (not fusion_shared or fusion_repeats != 1):
            raise NotImplementedError("the CAF kernels serve the RTFS-Net form of ATTNFusion (shared, fusion_repeats = 1)")
        mk = lambda vf: cls(ain_chan=audio_bn_chan, vin_chan=video_bn_chan, kernel_size=kernel_size, video_fusion=vf, is2d=is2d, **kwargs)
        if fusion_shared:
            self.fusion_module = mk(fusion_repeats > 1)
        else:  # TDAVNet/fusion.py:250-262: the last fusion does not feed the video stream
            self.fusion_module = nn.ModuleList(mk(i != fusion_repeats - 1) for i in range(fusion_repeats))

    def get_fusion_block(self, i):
        return self.fusion_module if self.fusion_shared else self.fusion_module[i]

    def forward(self, audio, video):
        a_res, v_res = audio, video
        a = v = None
        for i in range(self.fusion_repeats):  # TDAVNet/fusion.py:270-281
            a, v = self.get_fusion_block(i)(audio, video) if i == 0 else self.get_fusion_block(i)(a + a_res, v + v_res)
        return a


class RefinementModule(nn.Module):
    """TDAVNet/refinement_module.py:10-83."""

    def __init__(self, audio_params, video_params, audio_bn_chan, video_bn_chan, fusion_params):
        super().__init__()
        self.audio_params, self.video_params, self.fusion_params = audio_params, video_params, fusion_params
        self.fusion_repeats = video_params.get("repeats", 0)
        self.audio_repeats = audio_params["repeats"] - self.fusion_repeats
        from .generic import separators_get

        self.audio_net = separators_get(audio_params.get("audio_net", None))(**audio_params, in_chan=audio_bn_chan)
        self.video_net = separators_get(video_params.get("video_net", None))(**video_params, in_chan=video_bn_chan)
        self.crossmodal_fusion = MultiModalFusion(**fusion_params, audio_bn_chan=audio_bn_chan, video_bn_chan=video_bn_chan, fusion_repeats=self.fusion_repeats)
        # CUDA path: the RTFS-Net form (2-D shared TDANet audio block, CAF fusion once); anything else runs the eager schedule below
        self.fast = bool(getattr(self.audio_net, "is2d", False)) and isinstance(self.audio_net, TDANet) and \
            isinstance(self.crossmodal_fusion.get_fusion_block(0) if self.fusion_repeats > 0 else None, ATTNFusion)
        if bool(getattr(self.audio_net, "is2d", False)) and not self.fast:
            raise NotImplementedError("2-D audio blocks are served by the CUDA path only in the RTFS-Net form (TDANet is2d + ATTNFusion, fusion_repeats = 1)")
        self._rt = None

    def forward(self, audio, video):
        if self.fast:
            return self._rt().refine(audio, video)
        a_res, v_res = audio, video  # TDAVNet/refinement_module.py:45-62, eager (1-D configurations)
        for i in range(self.fusion_repeats):
            audio = self.audio_net.get_block(i)(audio + a_res if i > 0 else audio)
            video = self.video_net.get_block(i)(video + v_res if i > 0 else video)
            audio, video = self.crossmodal_fusion.get_fusion_block(i)(audio, video)
        for j in range(self.audio_repeats):
            i = j + self.fusion_repeats
            audio = self.audio_net.get_block(i)(audio + a_res if i > 0 else audio)
        return audio

    def get_MACs(self, bn_audio, bn_video):
        """[MACs (M), params (K)] of audio_net, video_net and crossmodal_fusion, in the reference's order
        (TDAVNet/refinement_module.py:74-83).  The fused audio path is counted in closed form, eager
        modules with forward hooks (generic.count_macs) instead of thop."""
        from .generic import count_macs

        out = []
        if self.fast:
            T, Fq = bn_audio.shape[-2], bn_audio.shape[-1]
            out += [int(self.audio_params["repeats"] * rtfs_block_macs(T, Fq) / 1e6), int(sum(p.numel() for p in self.audio_net.parameters()) / 1e3)]
            m, p = count_macs(self.video_net.get_block(0), (bn_video.float().cpu(),)) if next(self.video_net.parameters()).device.type == "cpu" else \
                count_macs(self.video_net.get_block(0), (bn_video,))
            out += [int(m * self.fusion_repeats / 1e6), int(sum(q.numel() for q in self.video_net.parameters()) / 1e3)]
            Tv = bn_video.shape[-1]
            caf = T * Fq * 256 * 2 + Tv * (256 * 2 + 1024 * 2)
            out += [int(caf / 1e6), int(sum(q.numel() for q in self.crossmodal_fusion.parameters()) / 1e3)]
            return out
        for mod, inp in ((self.audio_net, (bn_audio,)), (self.video_net, (bn_video,)), (self.crossmodal_fusion, (bn_audio, bn_video))):
            m, p = count_macs(mod, inp)
            out += [int(m / 1e6), int(p / 1e3)]
        return out


class STFTEncoder(nn.Module):
    """TDAVNet/encoder.py:122-175."""

    def __init__(self, win, hop_length, out_chan=2, kernel_size=-1, stride=1, act_type="ReLU", norm_type="gLN", bias=False, *args, **kwargs):
        super().__init__()
        if not (win == 256 and hop_length == 128 and out_chan == 256 and kernel_size == 3 and stride == 1 and act_type is None and norm_type is None and not bias):
            raise NotImplementedError("the STFT encoder kernel is built for win 256 / hop 128 / 3x3 conv to 256 channels, no norm/act/bias")
        self.win, self.hop_length, self.out_chan = win, hop_length, out_chan
        self.conv = ConvNormAct(2, out_chan, kernel_size, stride=stride, act_type=act_type, norm_type=norm_type, xavier_init=True, bias=bias, is2d=True)
        self.register_buffer("window", torch.hann_window(win), False)
        self._rt = None

    def get_out_chan(self):
        return self.out_chan

    def forward(self, x):
        return self._rt().encode(x)


class STFTDecoder(nn.Module):
    """TDAVNet/decoder.py:72-132."""

    def __init__(self, win, hop_length, in_chan, n_src, kernel_size=-1, stride=1, bias=False, *args, **kwargs):
        super().__init__()
        if not (win == 256 and hop_length == 128 and in_chan == 256 and n_src == 1 and kernel_size == 3 and stride == 1 and not bias):
            raise NotImplementedError("the iSTFT decoder kernel is built for win 256 / hop 128 / 3x3 transposed conv from 256 channels, n_src 1")
        self.decoder = nn.ConvTranspose2d(in_chan, 2, kernel_size, stride=stride, padding=(kernel_size - 1) // 2, bias=bias)
        nn.init.xavier_uniform_(self.decoder.weight)
        self.register_buffer("window", torch.hann_window(win), False)
        self._rt = None

    def forward(self, x, input_shape):
        return self._rt().decode(x, input_shape)


class MaskGenerator(nn.Module):
    """TDAVNet/mask_generator.py:20-99 (RI_split S^3 form)."""

    def __init__(self, n_src, audio_emb_dim, bottleneck_chan, kernel_size=1, mask_act="ReLU", RI_split=False, output_gate=False,
                 dw_gate=False, direct=False, is2d=False, *args, **kwargs):
        super().__init__()
        self.fast = bool(n_src == 1 and audio_emb_dim == 256 and bottleneck_chan == 256 and kernel_size == 1 and mask_act == "ReLU" and RI_split
                         and not output_gate and not direct and is2d)
        if is2d and not self.fast:
            raise NotImplementedError("the 2-D mask kernel is built for the RTFS-Net S^3 head (RI_split, ReLU, n_src 1, 256 channels)")
        self.n_src, self.in_chan, self.RI_split, self.output_gate, self.direct = n_src, audio_emb_dim, RI_split, output_gate, direct
        if not direct:
            chan = n_src * audio_emb_dim
            self.mask_generator = nn.Sequential(nn.PReLU(), ConvNormAct(bottleneck_chan, chan, kernel_size, act_type=mask_act, is2d=is2d))
            if output_gate:
                groups = chan if dw_gate else 1
                self.output = ConvNormAct(chan, chan, 1, act_type="Tanh", is2d=is2d, groups=groups)
                self.gate = ConvNormAct(chan, chan, 1, act_type="Sigmoid", is2d=is2d, groups=groups)
        self._rt = None

    def forward(self, refined_features, audio_mixture_embedding):
        if self.fast:
            return self._rt().mask(refined_features, audio_mixture_embedding)
        if self.direct:  # 1-D configurations (CTCNet), eager: mask_generator.py:84-99
            return refined_features
        m = self.mask_generator(refined_features)
        if self.output_gate:
            m = self.output(m) * self.gate(m)
        e = audio_mixture_embedding
        B, dims = e.shape[0], tuple(e.shape[-(e.ndim // 2):])
        if self.RI_split:
            m = m.view(B, self.n_src, 2, self.in_chan // 2, *dims)
            e = e.view(B, 2, self.in_chan // 2, *dims)
            mr, mi, er, ei = m[:, :, 0], m[:, :, 1], e[:, 0].unsqueeze(1), e[:, 1].unsqueeze(1)
            return torch.cat([er * mr - ei * mi, er * mi + ei * mr], 2)
        return m.view(B, self.n_src, self.in_chan, *dims) * e.unsqueeze(1)


def rtfs_block_macs(T, Fq):
    """Multiply-accumulates of one RTFS block pass on a (256, T, Fq) input."""
    Tc, Fc = (T - 2) // 2 + 1, 64
    P, Pc = T * Fq, Tc * Fc
    blk = 2 * P * 256 * 64 + 3 * P * 64 * 16 + 8 * Pc * 64 * 16
    for S, O in ((Fc, Tc), (Tc, Fc)):
        blk += O * (S - 7) * (512 * 256 + 3 * 64 * 192 + 64 * 512)
    return blk + Pc * 64 * 96 + 4 * Tc * Tc * (256 + 1024) + Pc * 64 * 64


# =============================================================================== runtime
def _nhwc(x):
    """Storage view (B,T,F,C) of a logical (B,C,T,F) tensor; converts only if it is not channels-last."""
    y = x.permute(0, 2, 3, 1)
    return y if y.is_contiguous() else y.contiguous()


def _nchw_view(y):
    return y.permute(0, 3, 1, 2)


class _Runtime:
    """Weight packing + workspace + C-ABI calls for one AVNet instance."""

    def __init__(self, model):
        self.model = model
        self._packed = None
        self._key = None
        self._ws = None
        self._ws_key = None
        self._vgraph = {}  # (device, shape, parameter key) -> (CUDAGraph, static input, static output) of the video block
        self._train_bufs = None
        self._side_streams = {}  # device index -> high-priority stream of the forked VP block

    # ---------------------------------------------------------------- plumbing
    def _needs_train_path(self):
        """train() mode (batch-statistics BatchNorm, dropout in the video block) or autograd through the parameters."""
        return self.model.training or (torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters()))

    def train_buffers(self, B, L, Tv, R, device):
        from .train import TrainBuffers

        key = (B, L, Tv, R, str(device))
        if self._train_bufs is None or self._train_bufs.key != key:
            self._train_bufs = None  # release before allocating the new tape
            self._train_bufs = TrainBuffers(B, L, Tv, R, device)
        return self._train_bufs

    def _check(self, *tensors, module_level=True):
        if module_level and self._needs_train_path():
            raise NotImplementedError(
                "module-level calls run the inference kernels: call them under torch.no_grad() in eval() mode "
                "(training goes through AVNet.forward, which keeps the tape the backward kernels need)")
        for t in tensors:
            if t is None:
                continue
            if not t.is_cuda:
                raise RuntimeError("rtfs_net_b200 runs on CUDA tensors only (there is no CPU fallback)")
            if t.dtype != torch.float32:
                raise TypeError("rtfs_net_b200 expects float32 tensors")

    def params(self, device):
        tensors = list(self.model.parameters()) + list(self.model.buffers())
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in tensors)
        if key != self._key:
            with torch.no_grad():
                self._packed = PackedParams(self.model.state_dict(), device)
            self._key = key
        return self._packed

    def workspace(self, B, L, Tv, device):
        key = (B, L // 128 + 1, Tv, str(device))
        if key != self._ws_key:
            total, _ = _lib.ws_plan(B, L, Tv)
            self._ws = None
            self._ws = torch.empty(total, dtype=torch.uint8, device=device)
            self._ws_key = key
        return self._ws

    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    # ---------------------------------------------------------------- module-level ops
    def encode(self, wav):
        if wav.ndim == 1:
            wav = wav[None]
        elif wav.ndim == 3:
            wav = wav[:, 0]
        assert wav.ndim == 2, f"Expected input to be 1D, 2D or 3D tensor, got {wav.ndim}D"
        self._check(wav)
        wav = wav.contiguous()
        B, L = wav.shape
        T = L // 128 + 1
        with torch.cuda.device(wav.device):
            P = self.params(wav.device)
            ws = self.workspace(B, L, 0, wav.device)
            a0 = torch.empty(B, T, 129, 256, device=wav.device, dtype=torch.float32)
            _lib.check(_lib.lib().rtfs_encoder_forward(P.ptr, wav.data_ptr(), a0.data_ptr(), ws.data_ptr(), B, L, self._stream()), "rtfs_encoder_forward")
        return _nchw_view(a0)

    def bottleneck(self, a0):
        self._check(a0)
        x = _nhwc(a0)
        B, T = x.shape[0], x.shape[1]
        with torch.cuda.device(x.device):
            P = self.params(x.device)
            ws = self.workspace(B, (T - 1) * 128, 0, x.device)
            out = torch.empty_like(x)
            _lib.check(_lib.lib().rtfs_bottleneck_forward(P.ptr, x.data_ptr(), out.data_ptr(), ws.data_ptr(), B, T, self._stream()), "rtfs_bottleneck_forward")
        return _nchw_view(out)

    def block(self, x, addend=None):
        self._check(x, addend)
        x = _nhwc(x)
        add = _nhwc(addend) if addend is not None else None
        B, T = x.shape[0], x.shape[1]
        with torch.cuda.device(x.device):
            P = self.params(x.device)
            ws = self.workspace(B, (T - 1) * 128, 0, x.device)
            out = torch.empty_like(x)
            _lib.check(_lib.lib().rtfs_block_forward(P.ptr, x.data_ptr(), add.data_ptr() if add is not None else None, out.data_ptr(),
                                                     ws.data_ptr(), B, T, self._stream()), "rtfs_block_forward")
        return _nchw_view(out)

    def _full_T(self, Tc):
        return 2 * Tc + 1  # any T with (T-2)//2+1 == Tc plans the same compressed buffers

    def dprnn(self, g, which):
        self._check(g)
        x = _nhwc(g)
        B, Tc = x.shape[0], x.shape[1]
        T = self._full_T(Tc)
        with torch.cuda.device(x.device):
            P = self.params(x.device)
            ws = self.workspace(B, (T - 1) * 128, 0, x.device)
            out = torch.empty_like(x)
            _lib.check(_lib.lib().rtfs_dprnn_forward(P.ptr, which, x.data_ptr(), out.data_ptr(), ws.data_ptr(), B, T, self._stream()), "rtfs_dprnn_forward")
        return _nchw_view(out)

    def mhsa(self, g):
        self._check(g)
        x = _nhwc(g)
        B, Tc = x.shape[0], x.shape[1]
        T = self._full_T(Tc)
        with torch.cuda.device(x.device):
            P = self.params(x.device)
            ws = self.workspace(B, (T - 1) * 128, 0, x.device)
            out = torch.empty_like(x)
            _lib.check(_lib.lib().rtfs_mhsa_forward(P.ptr, x.data_ptr(), out.data_ptr(), ws.data_ptr(), B, T, self._stream()), "rtfs_mhsa_forward")
        return _nchw_view(out)

    def caf(self, audio, video, addend=None):
        self._check(audio, video, addend)
        x = _nhwc(audio)
        v = video.contiguous()
        add = _nhwc(addend) if addend is not None else None
        B, T, Tv = x.shape[0], x.shape[1], v.shape[-1]
        with torch.cuda.device(x.device):
            P = self.params(x.device)
            ws = self.workspace(B, (T - 1) * 128, Tv, x.device)
            out = torch.empty_like(x)
            _lib.check(_lib.lib().rtfs_caf_forward(P.ptr, x.data_ptr(), v.data_ptr(), add.data_ptr() if add is not None else None, out.data_ptr(),
                                                   ws.data_ptr(), B, T, Tv, self._stream()), "rtfs_caf_forward")
        return _nchw_view(out)

    def refine(self, audio, video):
        """RefinementModule.forward (refinement_module.py:45-62), fusion_repeats = 1."""
        rm = self.model.refinement_module
        R = rm.audio_params["repeats"]
        a = self.block(audio)
        v = rm.video_net.get_block(0)(video)
        a = self.caf(a, v, audio if R > 1 else None)
        for i in range(1, R):
            a = self.block(a, audio if i + 1 < R else None)
        return a

    def mask(self, refined, a0):
        self._check(refined, a0)
        x, e = _nhwc(refined), _nhwc(a0)
        B, T = x.shape[0], x.shape[1]
        with torch.cuda.device(x.device):
            P = self.params(x.device)
            z = torch.empty_like(x)
            _lib.check(_lib.lib().rtfs_mask_forward(P.ptr, x.data_ptr(), e.data_ptr(), z.data_ptr(), B, T, self._stream()), "rtfs_mask_forward")
        return _nchw_view(z).unsqueeze(1)  # (B, n_src=1, C, T, F)

    def decode(self, z, input_shape):
        self._check(z)
        B = z.shape[0]
        zz = _nhwc(z.reshape(B * z.shape[1], z.shape[-3], z.shape[-2], z.shape[-1]) if z.ndim == 5 else z)
        L = int(input_shape[-1])
        with torch.cuda.device(zz.device):
            P = self.params(zz.device)
            ws = self.workspace(zz.shape[0], L, 0, zz.device)
            out = torch.empty(zz.shape[0], L, device=zz.device, dtype=torch.float32)
            _lib.check(_lib.lib().rtfs_decoder_forward(P.ptr, zz.data_ptr(), out.data_ptr(), ws.data_ptr(), zz.shape[0], L, self._stream()), "rtfs_decoder_forward")
        return out.view(B, -1, L)

    # ---------------------------------------------------------------- video (VP) block
    @staticmethod
    def _video_kernel_ok(P, Tv):
        return P.tensors.get("RTFS_P_VIDEO_PACK") is not None and 8 <= Tv <= 100 and not os.environ.get("RTFS_TORCH_VIDEO")

    def _side_stream(self, device):
        """One high-priority stream per device for the forked VP block (joined before the forward returns)."""
        key = torch.device(device).index
        st = self._side_streams.get(key)
        if st is None:
            st = self._side_streams[key] = torch.cuda.Stream(device=device, priority=-1)
        return st

    def video_block(self, mouth):
        """The VP block (tdanet.py:106-133 in 1-D, attention.py:9-73,192-220): one hand-written kernel per call (csrc/video.cuh)
        for the RTFS-Net video configuration and 8..100 frames.  Other shapes (and RTFS_TORCH_VIDEO=1, the A/B switch) run the
        torch modules: ~150 launch-bound library ops, captured once per (shape, parameter version) into a CUDA graph and
        replayed (RTFS_NO_VIDEO_GRAPH=1: eagerly)."""
        rm = self.model.refinement_module
        P = self.params(mouth.device)
        Tv = mouth.shape[-1]
        if self._video_kernel_ok(P, Tv):
            # the VP block as one kernel (csrc/video.cuh); the video bottleneck of the RTFS-Net configurations is the identity
            x = self.model.video_bottleneck(mouth).contiguous()
            out = torch.empty_like(x)
            _lib.check(_lib.lib().rtfs_video_forward(P.ptr, x.data_ptr(), out.data_ptr(), x.shape[0], Tv, self._stream()), "rtfs_video_forward")
            return out
        if os.environ.get("RTFS_NO_VIDEO_GRAPH") or torch.cuda.is_current_stream_capturing():
            # (a forward that is itself being captured into a CUDA graph records the eager ops directly)
            return rm.video_net.get_block(0)(self.model.video_bottleneck(mouth)).contiguous()
        key = (str(mouth.device), tuple(mouth.shape), self._key)
        ent = self._vgraph.pop(key, None)
        if ent is None:
            # stale parameter versions can never be replayed again: drop them; keep a few shapes (e.g. the ragged last
            # batch of an evaluation set alternating with the regular one) instead of re-capturing on every change
            for k in [k for k in self._vgraph if k[2] != self._key]:
                del self._vgraph[k]
            while len(self._vgraph) >= 4:
                del self._vgraph[next(iter(self._vgraph))]
            static_in = mouth.detach().clone()
            side = torch.cuda.Stream(device=mouth.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up on a side stream (lazy cuDNN / cuBLAS initialisation) before capture
                for _ in range(2):
                    rm.video_net.get_block(0)(self.model.video_bottleneck(static_in))
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = rm.video_net.get_block(0)(self.model.video_bottleneck(static_in)).contiguous()
            ent = (graph, static_in, static_out)
        self._vgraph[key] = ent  # (re-)insert last: the dict order is the LRU order
        graph, static_in, static_out = ent
        static_in.copy_(mouth)
        graph.replay()
        return static_out

    @staticmethod
    def _validate(wav, mouth):
        """Shape limits of the kernels, checked up front with a readable message."""
        if mouth is None:
            raise ValueError("the RTFS-Net configurations are audio-visual: mouth_embedding (B,512,Tv) is required")
        B, L = wav.shape
        T = L // 128 + 1
        if mouth.ndim != 3 or mouth.shape[0] != B or mouth.shape[1] != 512:
            raise ValueError(f"mouth_embedding must be (B={B}, 512, Tv), got {tuple(mouth.shape)}")
        if T < 16:
            raise ValueError(f"mixture too short: {L} samples give {T} STFT frames, the dual-path RNN needs >= 16")
        if not 1 <= mouth.shape[-1] <= 200:
            raise ValueError(f"video length {mouth.shape[-1]} outside 1..200 frames (CAF soft-max tile in shared memory)")
        if (T - 2) // 2 + 1 > 352:
            raise ValueError(f"mixture too long for one call: {(T - 2) // 2 + 1} compressed frames > 352 (attention score tile in shared memory); split the utterance")
        if B * T * 129 * 256 >= 2 ** 31:
            raise ValueError(f"batch too large for one call: B*T*F*256 = {B * T * 129 * 256} must stay below 2^31 elements")

    # ---------------------------------------------------------------- whole forward (one C call)
    def forward(self, wav, mouth):
        if wav.ndim == 1:
            wav = wav[None]
        elif wav.ndim == 3:
            wav = wav[:, 0]
        self._check(wav, mouth, module_level=False)
        self._validate(wav, mouth)
        if self._needs_train_path():
            from .train import forward_train

            return forward_train(self, wav, mouth)
        wav = wav.contiguous()
        B, L = wav.shape
        rm = self.model.refinement_module
        R = rm.audio_params["repeats"]
        with torch.cuda.device(wav.device):
            P = self.params(wav.device)
            Tv = mouth.shape[-1]
            out = torch.empty(B, L, device=wav.device, dtype=torch.float32)
            if self._video_kernel_ok(P, Tv):
                # one C call from the raw lip embedding; the VP block + the CAF video branch run on a high-priority side stream
                # next to the first block pass (RTFS_NO_VIDEO_FORK=1, the A/B switch: everything on the current stream)
                x = self.model.video_bottleneck(mouth).contiguous()
                video = torch.empty_like(x)
                ws = self.workspace(B, L, Tv, wav.device)
                side = 0 if os.environ.get("RTFS_NO_VIDEO_FORK") else self._side_stream(wav.device).cuda_stream
                _lib.check(_lib.lib().rtfs_avnet_forward_av(P.ptr, wav.data_ptr(), x.data_ptr(), video.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                                            B, L, Tv, R, self._stream(), side), "rtfs_avnet_forward_av")
            else:
                video = self.video_block(mouth.contiguous())
                ws = self.workspace(B, L, Tv, wav.device)
                _lib.check(_lib.lib().rtfs_avnet_forward(P.ptr, wav.data_ptr(), video.data_ptr(), out.data_ptr(), ws.data_ptr(), B, L, Tv, R, self._stream()),
                           "rtfs_avnet_forward")
        return out.view(B, 1, L)


# =============================================================================== AVNet
class BaseAVModel(nn.Module):
    """TDAVNet/base_av_model.py:9-118."""

    @staticmethod
    def load_state_dict_in(model, pretrained_dict):
        model_dict = model.state_dict()
        model_dict.update({k[12:]: v for k, v in pretrained_dict.items() if "audio_model" in k})
        model.load_state_dict(model_dict)
        return model

    @staticmethod
    def from_pretrain(pretrained_model_conf_or_path, *args, **kwargs):
        from . import get

        conf = torch.load(pretrained_model_conf_or_path, map_location="cpu")
        model = get(conf["model_name"])(print_macs=False, *args, **kwargs)
        model.load_state_dict(conf["state_dict"])
        return model

    def serialize(self):
        infos = dict(software_versions=dict(torch_version=torch.__version__, python_version=sys.version))
        return dict(model_name=self.__class__.__name__, state_dict=self.get_state_dict(), model_args=self.get_config(), infos=infos)

    def get_state_dict(self):
        return self.state_dict()


class AVNet(BaseAVModel):
    """Drop-in for the reference's AVNet (tdavnet.py:14-108) restricted to the RTFS-Net family."""

    def __init__(self, n_src, enc_dec_params, audio_bn_params, audio_params, mask_generation_params, pretrained_vout_chan=-1,
                 video_bn_params=dict(), video_params=dict(), fusion_params=dict(), print_macs=True, *args, **kwargs):
        super().__init__()
        self.n_src = n_src
        self.pretrained_vout_chan = pretrained_vout_chan
        self.audio_bn_params, self.video_bn_params = dict(audio_bn_params), dict(video_bn_params)
        self.enc_dec_params, self.audio_params, self.video_params = dict(enc_dec_params), dict(audio_params), dict(video_params)
        self.fusion_params, self.mask_generation_params = dict(fusion_params), dict(mask_generation_params)
        self.print_macs = print_macs
        from .generic import decoder_get, encoder_get, mask_generator_get

        # tdavnet.py:37-84, string-keyed as in the reference.  STFT encoder => the RTFS-Net CUDA path (every component must then
        # be in its RTFS-Net form: the constructors below raise otherwise -- there is no eager fallback on that path); any other
        # encoder => the 1-D family (CTCNet) as plain PyTorch modules (generic.py), an API-compatibility path.
        self.fast = self.enc_dec_params.get("encoder_type") == "STFTEncoder"
        self.encoder = encoder_get(self.enc_dec_params["encoder_type"])(**self.enc_dec_params, in_chan=1, upsampling_depth=self.audio_params.get("upsampling_depth", 1))
        self.enc_out_chan = self.encoder.get_out_chan()
        self.mask_generation_params["mask_generator_type"] = self.mask_generation_params.get("mask_generator_type", "MaskGenerator")
        self.audio_bn_chan = self.audio_bn_params.get("out_chan", self.enc_out_chan)
        self.audio_bn_params["out_chan"] = self.audio_bn_chan
        self.video_bn_chan = self.video_bn_params.get("out_chan", self.pretrained_vout_chan)
        self.audio_bottleneck = ConvNormAct(**self.audio_bn_params, in_chan=self.enc_out_chan)
        self.video_bottleneck = ConvNormAct(**self.video_bn_params, in_chan=self.pretrained_vout_chan)
        if self.fast:
            bn = self.audio_bn_params
            if not (bn.get("pre_norm_type") == "gLN" and bn.get("pre_act_type") == "ReLU" and bn.get("kernel_size") == 1 and bn.get("is2d")
                    and self.audio_bn_chan == 256 and bn.get("norm_type") is None and bn.get("act_type") is None):
                raise NotImplementedError("the bottleneck kernel is built for gLN -> ReLU -> 1x1 conv 256 -> 256")
            if self.video_bn_params.get("kernel_size", -1) > 0:
                raise NotImplementedError("a non-identity video bottleneck is outside the RTFS-Net configurations")
            if self.enc_dec_params.get("decoder_type") != "STFTDecoder":
                raise NotImplementedError("the STFT encoder pairs with the STFT decoder on the CUDA path")
        self.refinement_module = RefinementModule(fusion_params=self.fusion_params, audio_params=self.audio_params, video_params=self.video_params,
                                                  audio_bn_chan=self.audio_bn_chan, video_bn_chan=self.video_bn_chan)
        self.mask_generator = mask_generator_get(self.mask_generation_params["mask_generator_type"])(
            **self.mask_generation_params, n_src=self.n_src, audio_emb_dim=self.enc_out_chan, bottleneck_chan=self.audio_bn_chan)
        self.decoder = decoder_get(self.enc_dec_params["decoder_type"])(**self.enc_dec_params, in_chan=self.enc_out_chan * self.n_src, n_src=self.n_src)
        if self.fast and not (self.refinement_module.fast and self.mask_generator.fast):
            raise NotImplementedError("STFT-domain configurations are served by the CUDA path in their RTFS-Net form only")

        rt = _Runtime(self)
        object.__setattr__(self, "_runtime", rt)
        getter = lambda: rt
        for m in self.modules():
            if hasattr(m, "_rt"):
                object.__setattr__(m, "_rt", getter)
        if self.fast:
            object.__setattr__(self.audio_bottleneck, "_fused", rt.bottleneck)
        if self.print_macs:
            self.get_MACs()

    def forward(self, audio_mixture, mouth_embedding=None):
        if self.fast:
            return self._runtime.forward(audio_mixture, mouth_embedding)
        emb = self.encoder(audio_mixture)  # tdavnet.py:86-97, eager (1-D configurations)
        refined = self.refinement_module(self.audio_bottleneck(emb), self.video_bottleneck(mouth_embedding))
        return self.decoder(self.mask_generator(refined, emb), audio_mixture.shape)

    def get_config(self):
        return dict(n_src=self.n_src, pretrained_vout_chan=self.pretrained_vout_chan, enc_dec_params=self.enc_dec_params,
                    audio_bn_params=self.audio_bn_params, video_bn_params=self.video_bn_params, audio_params=self.audio_params,
                    video_params=self.video_params, fusion_params=self.fusion_params, mask_generation_params=self.mask_generation_params)

    def get_MACs(self):
        """MAC / parameter table at the reference's convention (B = 1, 2 s of audio, 50 video frames; base_av_model.py:61-118).
        CUDA path: closed form for the fused audio stages, forward hooks for the eager video block -- no
        CPU forward of the audio path is run; 1-D configurations: forward hooks on the eager modules (generic.count_macs)."""
        from .generic import count_macs

        params = lambda m: sum(p.numel() for p in m.parameters())
        dev = next(self.parameters()).device
        v_chan = self.pretrained_vout_chan if self.pretrained_vout_chan > 0 else 1
        video = torch.rand(1, v_chan, 50, device=dev)
        if self.fast:
            T, Fq = 2 * 16000 // 128 + 1, 129
            P = T * Fq
            R = self.audio_params["repeats"]
            macs = dict(encoder=P * 18 * 256, audio_bn=P * 256 * 256, audio_net=R * rtfs_block_macs(T, Fq), mask=P * 256 * 256, decoder=P * 18 * 256)
            with torch.no_grad():
                rm = self.refinement_module.get_MACs(torch.empty(1, 256, T, Fq, device="meta"), video)
            rows = [("Encoder", macs["encoder"], params(self.encoder)), ("Audio BN", macs["audio_bn"], params(self.audio_bottleneck)),
                    ("Video BN", 0, params(self.video_bottleneck)), ("   AudioNet", rm[0] * 1e6, rm[1] * 1e3), ("   VideoNet", rm[2] * 1e6, rm[3] * 1e3),
                    ("   FusionNet", rm[4] * 1e6, rm[5] * 1e3), ("Mask Generator", macs["mask"], params(self.mask_generator)),
                    ("Decoder", macs["decoder"], params(self.decoder))]
        else:
            audio = torch.rand(1, 2 * 16000, device=dev)
            was_training = self.training
            self.eval()
            with torch.no_grad():
                emb = self.encoder(audio)
                bn_a, bn_v = self.audio_bottleneck(emb), self.video_bottleneck(video)
                sep = self.mask_generator(bn_a, emb)
                rm = self.refinement_module.get_MACs(bn_a, bn_v)
                rows = [("Encoder",) + count_macs(self.encoder, (audio,)), ("Audio BN",) + count_macs(self.audio_bottleneck, (emb,)),
                        ("Video BN",) + count_macs(self.video_bottleneck, (video,)), ("   AudioNet", rm[0] * 1e6, rm[1] * 1e3),
                        ("   VideoNet", rm[2] * 1e6, rm[3] * 1e3), ("   FusionNet", rm[4] * 1e6, rm[5] * 1e3),
                        ("Mask Generator",) + count_macs(self.mask_generator, (bn_a, emb)), ("Decoder",) + count_macs(self.decoder, (sep, audio.shape))]
            self.train(was_training)
            macs = {k.strip().lower().replace(" ", "_"): v for k, v, _ in rows}
        total = sum(v for k, v, _ in rows)
        self.macs_parms = "".join(f"{k + ' ':-<22} MACs: {v / 1e6:>10.1f} M    Params: {p / 1e3:>8.1f} K\n" for k, v, p in rows) + \
            f"{'Total ':-<22} MACs: {total / 1e6:>10.1f} M    Params: {params(self) / 1e3:>8.1f} K\n"
        if self.print_macs:
            print(self.macs_parms)
        return macs
