"""Plain-PyTorch (library-op) modules for the reference configurations OUTSIDE the RTFS-Net hot path -- the 1-D CTCNet family of
config/lrs2_CTCNet_16_layer.yaml (BASELINE configs[4]: API-surface / drop-in compatibility check) -- and the string-keyed
factories the reference builds its models with.

Nothing here is accelerated and nothing here is used by the RTFS-Net configurations: `AVNet` dispatches on the configuration
(`enc_dec_params.encoder_type == "STFTEncoder"` -> the CUDA path of nn.py, which has no eager fallback; anything else -> these
modules).  Same constructor keywords, attribute names and therefore `state_dict` keys as the reference:
  ConvolutionalEncoder   TDAVNet/encoder.py:58-119        ConvolutionalDecoder   TDAVNet/decoder.py:25-69
  FRCNNBlock / FRCNN     separators/frcnn.py:8-237        ConcatFusion / SumFusion   TDAVNet/fusion.py:40-94
  factories              layers/__init__.py:19-31, layers/normalizations.py:44-58, layers/activations.py:4-18,
                         TDAVNet/encoder.py:178-189, TDAVNet/decoder.py:135-146, TDAVNet/mask_generator.py:190-201,
                         separators/__init__.py:8-20
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# =============================================================================== factories
def _resolve(identifier, table, what, fall_back_to_nn=False, known_elsewhere=()):
    """The reference's `get`: None -> nn.Identity, a callable -> itself, a string -> class from the module's namespace
    (optionally torch.nn first), anything else -> ValueError."""
    if identifier is None:
        return nn.Identity
    if callable(identifier):
        return identifier
    if isinstance(identifier, str):
        if fall_back_to_nn and hasattr(nn, identifier):
            return getattr(nn, identifier)
        cls = table.get(identifier)
        if cls is not None:
            return cls
        if identifier in known_elsewhere:
            raise NotImplementedError(f"{what} '{identifier}' exists in the reference but is outside this package's scope "
                                      "(RTFS-Net hot path + CTCNet compatibility)")
    raise ValueError(f"Could not interpret {what} identifier: " + str(identifier))


def normalizations_get(identifier):
    """layers/normalizations.py:44-58 (torch.nn first, then gLN / GlobalLayerNorm / LayerNormalization4D / LN4d)."""
    from . import nn as M

    table = {"gLN": M.GlobalLayerNorm, "GlobalLayerNorm": M.GlobalLayerNorm, "LayerNormalization4D": M.LayerNormalization4D, "LN4d": M.LayerNormalization4D}
    return _resolve(identifier, table, "normalization", fall_back_to_nn=True)


def activations_get(identifier):
    """layers/activations.py:4-18 (torch.nn only)."""
    return _resolve(identifier, {}, "activation", fall_back_to_nn=True)


def layers_get(identifier):
    """layers/__init__.py:19-31."""
    from . import nn as M

    table = {"ConvNormAct": M.ConvNormAct, "ConvActNorm": M.ConvActNorm, "FeedForwardNetwork": M.FeedForwardNetwork, "DualPathRNN": M.DualPathRNN,
             "InjectionMultiSum": M.InjectionMultiSum, "ATTNFusionCell": M.ATTNFusionCell, "GlobalAttention": M.GlobalAttention,
             "MultiHeadSelfAttention": M.MultiHeadSelfAttention, "MultiHeadSelfAttention2D": M.MultiHeadSelfAttention2D}
    other = ("ConvolutionalRNN", "DepthwiseSeparableConvolution", "BiLSTM2D", "RNNProjection", "GlobalGALR", "GlobalAttentionRNN", "ConvLSTMFusionCell",
             "ConvGRUFusionCell", "GlobalAttention2D", "CBAMBlock", "ShuffleAttention", "CoTAttention", "MLP", "Permutator")
    return _resolve(identifier, table, "layer", known_elsewhere=other)


def encoder_get(identifier):
    """TDAVNet/encoder.py:178-189."""
    from . import nn as M

    return _resolve(identifier, {"STFTEncoder": M.STFTEncoder, "ConvolutionalEncoder": ConvolutionalEncoder}, "encoder", known_elsewhere=("BaseEncoder",))


def decoder_get(identifier):
    """TDAVNet/decoder.py:135-146."""
    from . import nn as M

    return _resolve(identifier, {"STFTDecoder": M.STFTDecoder, "ConvolutionalDecoder": ConvolutionalDecoder}, "decoder", known_elsewhere=("BaseDecoder",))


def mask_generator_get(identifier):
    """TDAVNet/mask_generator.py:190-201."""
    from . import nn as M

    return _resolve(identifier, {"MaskGenerator": M.MaskGenerator}, "mask generator", known_elsewhere=("MaskGenerator2Chan",))


def separators_get(identifier):
    """separators/__init__.py:8-20."""
    from . import nn as M

    return _resolve(identifier, {"TDANet": M.TDANet, "FRCNN": FRCNN}, "separator", known_elsewhere=("DPTNet",))


def fusion_get(identifier):
    """MultiModalFusion's `globals().get(fusion_type)` (TDAVNet/fusion.py:238)."""
    from . import nn as M

    return _resolve(identifier, {"ATTNFusion": M.ATTNFusion, "ConcatFusion": ConcatFusion, "SumFusion": SumFusion}, "fusion",
                    known_elsewhere=("InjectionFusion", "LSTMFusion", "GRUFusion"))


# =============================================================================== 1-D encoder / decoder
class ConvolutionalEncoder(nn.Module):
    """Strided Conv1d front end (CTCNet): zero-pad the waveform to the two least common multiples the U-shaped separator
    needs, then sum `layers` dilated ConvNormAct branches."""

    def __init__(self, in_chan, out_chan, kernel_size, stride, act_type=None, norm_type="gLN", bias=False, layers=1, upsampling_depth=4, *args, **kwargs):
        super().__init__()
        from .nn import ConvNormAct

        self.in_chan, self.out_chan, self.kernel_size, self.stride = in_chan, out_chan, kernel_size, stride
        self.act_type, self.norm_type, self.bias, self.layers, self.upsampling_depth = act_type, norm_type, bias, layers, upsampling_depth
        down = 2 ** upsampling_depth
        g = math.gcd(kernel_size // 2, down)
        self.lcm_1 = abs(out_chan // 2 * down) // g
        self.lcm_2 = abs(kernel_size // 2 * down) // g
        self.encoder = nn.ModuleList(
            ConvNormAct(in_chan, out_chan, kernel_size * (i + 1), stride=stride, dilation=i + 1, norm_type=norm_type, act_type=act_type, xavier_init=True, bias=bias)
            for i in range(layers))

    def get_out_chan(self):
        return self.out_chan

    @staticmethod
    def _pad_to_multiple(x, m):
        r = x.shape[-1] % m
        return F.pad(x, (0, m - r)) if r else x

    def forward(self, x):
        if x.ndim == 1:
            x = x.reshape(1, 1, -1)
        elif x.ndim == 2:
            x = x.unsqueeze(1)
        x = self._pad_to_multiple(self._pad_to_multiple(x, self.lcm_1), self.lcm_2)
        out = self.encoder[0](x)
        for branch in list(self.encoder)[1:]:
            out = out + branch(x)
        return out


class ConvolutionalDecoder(nn.Module):
    """ConvTranspose1d back end (CTCNet), output padded to the input length."""

    def __init__(self, in_chan, n_src, kernel_size, stride, bias=False, *args, **kwargs):
        super().__init__()
        self.in_chan, self.n_src, self.kernel_size, self.stride, self.bias = in_chan, n_src, kernel_size, stride, bias
        self.padding = (kernel_size - 1) // 2
        self.output_padding = self.padding - 1
        self.decoder = nn.ConvTranspose1d(in_chan, 1, kernel_size, stride=stride, padding=self.padding, output_padding=self.output_padding, bias=bias)
        nn.init.xavier_uniform_(self.decoder.weight)

    def forward(self, separated_audio_embedding, input_shape):
        batch, length = input_shape[0], input_shape[-1]
        y = self.decoder(separated_audio_embedding.reshape(batch * self.n_src, self.in_chan, -1))
        y = F.pad(y, (0, length - y.shape[-1]))
        return y.view(batch, self.n_src, -1)


# =============================================================================== FRCNN separator (A-FRCNN block of CTCNet)
def _spatial(t):
    return t.shape[-(t.ndim // 2):]


class FRCNNBlock(nn.Module):
    def __init__(self, in_chan, hid_chan, kernel_size=5, stride=2, norm_type="gLN", act_type="PReLU", upsampling_depth=4, is2d=False):
        super().__init__()
        from .nn import ConvNormAct

        self.in_chan, self.hid_chan, self.upsampling_depth, self.is2d = in_chan, hid_chan, upsampling_depth, is2d
        depth = upsampling_depth
        dw = lambda s: ConvNormAct(hid_chan, hid_chan, kernel_size, stride=s, groups=hid_chan, norm_type=norm_type, is2d=is2d)
        self.gateway = ConvNormAct(in_chan, in_chan, 1, groups=in_chan, act_type=act_type, is2d=is2d)
        self.projection = ConvNormAct(in_chan, hid_chan, 1, is2d=is2d)
        self.downsample_layers = nn.ModuleList(dw(1 if i == 0 else stride) for i in range(depth))
        # fusion_layers[i][0]: strided conv that brings scale i-1 down to scale i (i >= 1); scale 0 has none
        self.fusion_layers = nn.ModuleList(nn.ModuleList([dw(stride)] if i > 0 else []) for i in range(depth))
        n_in = lambda i: 2 if i in (0, depth - 1) else 3
        self.concat_layers = nn.ModuleList(ConvNormAct(hid_chan * n_in(i), hid_chan, 1, norm_type=norm_type, act_type=act_type, is2d=is2d) for i in range(depth))
        self.residual_conv = nn.Sequential(ConvNormAct(hid_chan * depth, hid_chan, 1, norm_type=norm_type, act_type=act_type, is2d=is2d),
                                           ConvNormAct(hid_chan, in_chan, 1, is2d=is2d))

    def forward(self, x):
        depth = self.upsampling_depth
        residual = self.gateway(x)
        scales = [self.downsample_layers[0](self.projection(residual))]
        for i in range(1, depth):
            scales.append(self.downsample_layers[i](scales[-1]))
        fused = []
        for i in range(depth):  # lateral connections: the scale above (strided conv), the scale itself, the scale below (nearest)
            parts = []
            if i > 0:
                parts.append(self.fusion_layers[i][0](scales[i - 1]))
            parts.append(scales[i])
            if i + 1 < depth:
                parts.append(F.interpolate(scales[i + 1], size=_spatial(scales[i]), mode="nearest"))
            fused.append(self.concat_layers[i](torch.cat(parts, 1)))
        top = _spatial(scales[0])
        fused = [fused[0]] + [F.interpolate(t, size=top, mode="nearest") for t in fused[1:]]
        return self.residual_conv(torch.cat(fused, 1)) + residual


class FRCNN(nn.Module):
    def __init__(self, in_chan=-1, hid_chan=-1, kernel_size=5, stride=2, norm_type="gLN", act_type="PReLU", upsampling_depth=4, repeats=4,
                 shared=False, is2d=False, *args, **kwargs):
        super().__init__()
        self.repeats, self.shared = repeats, shared
        mk = (lambda: FRCNNBlock(in_chan, hid_chan, kernel_size, stride, norm_type, act_type, upsampling_depth, is2d)) if in_chan > 0 and hid_chan > 0 else nn.Identity
        self.blocks = mk() if shared else nn.ModuleList(mk() for _ in range(repeats))

    def get_block(self, i):
        return self.blocks if self.shared else self.blocks[i]

    def forward(self, x):
        residual = x
        for i in range(self.repeats):
            x = self.get_block(i)((x + residual) if i > 0 else x)
        return x


# =============================================================================== fusion modules
class _FusionBase(nn.Module):
    """TDAVNet/fusion.py:9-37: a 1-D stream is given a trailing singleton axis when the other stream is 2-D."""

    def __init__(self, ain_chan, vin_chan, kernel_size, video_fusion, is2d):
        super().__init__()
        self.ain_chan, self.vin_chan, self.kernel_size, self.video_fusion, self.is2d = ain_chan, vin_chan, kernel_size, video_fusion, is2d

    def _match_rank(self, audio, video):
        self._lift_v = len(_spatial(audio)) > len(_spatial(video))
        self._lift_a = len(_spatial(video)) > len(_spatial(audio))
        return (audio.unsqueeze(-1) if self._lift_a else audio), (video.unsqueeze(-1) if self._lift_v else video)

    def _restore_rank(self, audio, video):
        return (audio.squeeze(-1) if self._lift_a else audio), (video.squeeze(-1) if self._lift_v else video)


class ConcatFusion(_FusionBase):
    def __init__(self, ain_chan, vin_chan, kernel_size, video_fusion=True, is2d=False):
        super().__init__(ain_chan, vin_chan, kernel_size, video_fusion, is2d)
        from .nn import ConvNormAct

        self.audio_conv = ConvNormAct(ain_chan + vin_chan, ain_chan, kernel_size, norm_type="gLN", is2d=is2d)
        if video_fusion:
            self.video_conv = ConvNormAct(ain_chan + vin_chan, vin_chan, kernel_size, norm_type="gLN", is2d=is2d)

    def forward(self, audio, video):
        audio, video = self._match_rank(audio, video)
        a = self.audio_conv(torch.cat([audio, F.interpolate(video, size=_spatial(audio), mode="nearest")], 1))
        v = self.video_conv(torch.cat([F.interpolate(audio, size=_spatial(video), mode="nearest"), video], 1)) if self.video_fusion else video
        return self._restore_rank(a, v)


class SumFusion(_FusionBase):
    def __init__(self, ain_chan, vin_chan, kernel_size, video_fusion=True, is2d=False):
        super().__init__(ain_chan, vin_chan, kernel_size, video_fusion, is2d)
        from .nn import ConvNormAct

        if video_fusion:
            self.audio_conv = ConvNormAct(ain_chan, vin_chan, kernel_size, norm_type="gLN", is2d=is2d)
        self.video_conv = ConvNormAct(vin_chan, ain_chan, kernel_size, norm_type="gLN", is2d=is2d)

    def forward(self, audio, video):
        audio, video = self._match_rank(audio, video)
        v = self.audio_conv(F.interpolate(audio, size=_spatial(video), mode="nearest")) + video if self.video_fusion else video
        a = self.video_conv(F.interpolate(video, size=_spatial(audio), mode="nearest")) + audio
        return self._restore_rank(a, v)


# =============================================================================== MAC counting (thop stand-in)
def count_macs(module, inputs):
    """Multiply-accumulates of one forward of an eager torch module, counted with forward hooks on the contraction layers
    (the reference uses `thop.profile`, src/models/utils/utils.py:5-30).  Returns (MACs, parameters)."""
    total = [0]

    def conv_hook(m, inp, out):
        k = 1
        for s in m.kernel_size:
            k *= s
        if isinstance(m, (nn.ConvTranspose1d, nn.ConvTranspose2d)):  # every input element meets (out_ch / groups) * k weights
            total[0] += inp[0].numel() * (m.out_channels // m.groups) * k
        else:  # every output element is a dot product over (in_ch / groups) * k inputs
            total[0] += out.numel() * (m.in_channels // m.groups) * k

    def linear_hook(m, inp, out):
        total[0] += out.numel() * m.in_features

    def mha_hook(m, inp, out):
        q = inp[0]
        B, T, E = (q.shape if m.batch_first else (q.shape[1], q.shape[0], q.shape[2]))
        total[0] += B * T * E * 3 * E + 2 * B * T * T * E + B * T * E * E

    hooks = []
    for m in module.modules():
        if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.ConvTranspose1d, nn.ConvTranspose2d)):
            hooks.append(m.register_forward_hook(conv_hook))
        elif isinstance(m, nn.MultiheadAttention):
            hooks.append(m.register_forward_hook(mha_hook))
        elif isinstance(m, nn.Linear) and not isinstance(m, nn.modules.linear.NonDynamicallyQuantizableLinear):
            hooks.append(m.register_forward_hook(linear_hook))
    try:
        with torch.no_grad():
            module(*inputs)
    finally:
        for h in hooks:
            h.remove()
    return total[0], sum(p.numel() for p in module.parameters())
