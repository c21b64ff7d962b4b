import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

# the RTFS-Net `audionet` configuration (config/lrs2_RTFSNet_4_layer.yaml:8-104 of the reference)
AUDIONET_CONF = dict(
    n_src=1,
    pretrained_vout_chan=512,
    video_bn_params=dict(kernel_size=-1),
    audio_bn_params=dict(pre_norm_type="gLN", pre_act_type="ReLU", out_chan=256, kernel_size=1, is2d=True),
    enc_dec_params=dict(encoder_type="STFTEncoder", decoder_type="STFTDecoder", win=256, hop_length=128, out_chan=256,
                        kernel_size=3, stride=1, bias=False, act_type=None, norm_type=None),
    audio_params=dict(audio_net="TDANet", hid_chan=64, kernel_size=4, stride=2, norm_type="gLN", act_type="PReLU",
                      upsampling_depth=2, repeats=4, shared=True, is2d=True,
                      layers=dict(
                          layer_1=dict(layer_type="DualPathRNN", hid_chan=32, dim=4, kernel_size=8, stride=1, rnn_type="SRU", num_layers=4, bidirectional=True),
                          layer_2=dict(layer_type="DualPathRNN", hid_chan=32, dim=3, kernel_size=8, stride=1, rnn_type="SRU", num_layers=4, bidirectional=True),
                          layer_3=dict(layer_type="MultiHeadSelfAttention2D", dim=3, n_freqs=64, n_head=4, hid_chan=4, act_type="PReLU", norm_type="LayerNormalization4D"))),
    video_params=dict(video_net="TDANet", hid_chan=64, kernel_size=3, stride=2, norm_type="BatchNorm1d", act_type="PReLU",
                      upsampling_depth=4, repeats=1, shared=True, is2d=False,
                      layers=dict(layer_1=dict(layer_type="GlobalAttention", ffn_name="FeedForwardNetwork", kernel_size=3, n_head=8, dropout=0.1))),
    fusion_params=dict(fusion_type="ATTNFusion", fusion_shared=True, kernel_size=4, is2d=True),
    mask_generation_params=dict(mask_generator_type="MaskGenerator", mask_act="ReLU", RI_split=True, is2d=True),
)


def audionet_conf(repeats=4):
    import copy

    c = copy.deepcopy(AUDIONET_CONF)
    c["audio_params"]["repeats"] = repeats
    return c


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_sd():
    g = np.load(os.path.join(GOLD, "state_dict_rtfs.npz"))
    return {k: torch.from_numpy(g[k]) for k in g.files}


def load_case(tag):
    g = np.load(os.path.join(GOLD, tag + ".npz"))
    return {k: (torch.from_numpy(g[k]) if g[k].ndim > 0 else int(g[k])) for k in g.files}


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def strided(t, n=4096):
    flat = t.reshape(-1)
    step = max(1, flat.numel() // n)
    return flat[::step][:n].contiguous()


def build_model(sd, repeats=4, device="cuda"):
    from rtfs_net_b200 import AVNet

    m = AVNet(print_macs=False, **audionet_conf(repeats))
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval()
