"""TEST INFRASTRUCTURE ONLY -- golden vector of the reference's CTCNet configuration (BASELINE configs[4]).

Run in the authoring container (needs /root/reference):    python -m oracle.make_golden_ctcnet

Imports the reference's own `src.models.AVNet` (with the shims of oracle/shims for the absent third-party packages), builds
it from config/lrs2_CTCNet_16_layer.yaml, loads the SYNTHETIC state_dict of `synthetic_state_dict` below (7 M parameters: too
large to commit, so both sides regenerate it from the key names), runs the reference forward in eval mode on a seeded
1-second mixture and writes tests/golden/ctcnet_b1_1s.npz (inputs, output, and the (key, shape) list of the state_dict).
"""
import json
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")

CTCNET_AUDIONET = dict(
    n_src=1, pretrained_vout_chan=512,
    video_bn_params=dict(out_chan=64, kernel_size=1, is2d=False),
    audio_bn_params=dict(out_chan=512, kernel_size=1, is2d=False),
    enc_dec_params=dict(encoder_type="ConvolutionalEncoder", decoder_type="ConvolutionalDecoder", out_chan=512, kernel_size=21, stride=10, bias=False,
                        act_type="ReLU", norm_type="gLN", layers=1),
    audio_params=dict(audio_net="FRCNN", hid_chan=512, upsampling_depth=5, shared=True, repeats=16, norm_type="gLN", act_type="PReLU", kernel_size=5, stride=2, is2d=False),
    video_params=dict(video_net="FRCNN", hid_chan=64, upsampling_depth=4, shared=False, repeats=3, norm_type="BatchNorm1d", act_type="PReLU", kernel_size=3, stride=2, is2d=False),
    fusion_params=dict(fusion_type="ConcatFusion", fusion_shared=False, is2d=False),
    mask_generation_params=dict(mask_act="ReLU", is2d=False, output_gate=False),
)  # = the `audionet` section of config/lrs2_CTCNet_16_layer.yaml (checked against the yaml below)


def synthetic_state_dict(template):
    """Deterministic parameter values from the key names (CRC32 seeds): weights ~ N(0, 1/fan_in), norm scales around 1,
    offsets small, BatchNorm statistics away from (0, 1), PReLU slopes 0.25."""
    out = {}
    for k in sorted(template):
        v = template[k]
        g = torch.Generator().manual_seed(zlib.crc32(k.encode()))
        if not v.dtype.is_floating_point:
            out[k] = torch.zeros_like(v)
        elif k.endswith("running_var"):
            out[k] = 0.5 + torch.rand(v.shape, generator=g)
        elif k.endswith("running_mean"):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif v.ndim >= 2:
            fan_in = int(np.prod(v.shape[1:]))
            out[k] = torch.randn(v.shape, generator=g) / max(fan_in, 1) ** 0.5
        elif v.numel() == 1:
            out[k] = torch.full(v.shape, 0.25)
        elif k.endswith("weight"):
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:
            out[k] = 0.05 * torch.randn(v.shape, generator=g)
    return out


def inputs(L=16000, Tv=25):
    g = torch.Generator().manual_seed(11)
    return 0.1 * torch.randn(1, L, generator=g), torch.rand(1, 512, Tv, generator=g)


def main():
    import yaml

    from oracle.make_golden import REF, import_reference

    AVNet = import_reference()
    with open(os.path.join(REF, "config", "lrs2_CTCNet_16_layer.yaml")) as f:
        conf = yaml.safe_load(f)["audionet"]
    assert conf == CTCNET_AUDIONET, "CTCNET_AUDIONET drifted from the reference yaml"
    import copy

    model = AVNet(print_macs=False, **copy.deepcopy(conf)).eval()
    sd = synthetic_state_dict(model.state_dict())
    model.load_state_dict(sd, strict=True)
    wav, lip = inputs()
    with torch.no_grad():
        out = model(wav, lip)
    keys = [[k, list(v.shape)] for k, v in model.state_dict().items()]
    np.savez_compressed(os.path.join(GOLD, "ctcnet_b1_1s.npz"), wav=wav.numpy(), lip=lip.numpy(), out_ref_fp32=out.numpy(),
                        keys=np.frombuffer(json.dumps(keys).encode(), dtype=np.uint8), n_params=sum(p.numel() for p in model.parameters()))
    print("wrote ctcnet_b1_1s.npz: out", tuple(out.shape), "rms", float(out.pow(2).mean().sqrt()), "params", sum(p.numel() for p in model.parameters()))


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
