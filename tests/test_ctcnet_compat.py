"""BASELINE configs[4]: lrs2_CTCNet_16_layer.yaml as an API-surface / drop-in compatibility check (CPU test; the 1-D CTCNet family
runs as plain PyTorch modules -- rtfs_net_b200/generic.py -- and is not accelerated).  The golden output was produced by the
reference's own AVNet (oracle/make_golden_ctcnet.py) from a synthetic state_dict that both sides regenerate from the key names."""
import copy
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD, rel_l2


@pytest.fixture(scope="module")
def case():
    g = np.load(os.path.join(GOLD, "ctcnet_b1_1s.npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="module")
def model():
    from oracle.make_golden_ctcnet import CTCNET_AUDIONET, synthetic_state_dict
    from rtfs_net_b200 import AVNet

    m = AVNet(print_macs=False, **copy.deepcopy(CTCNET_AUDIONET)).eval()
    m.load_state_dict(synthetic_state_dict(m.state_dict()), strict=True)
    return m


def test_ctcnet_constructs_with_the_reference_state_dict_layout(case, model):
    keys = json.loads(bytes(case["keys"]).decode())
    ours = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert list(ours.items()) == [(k, s) for k, s in keys]  # same keys, same order, same shapes
    assert sum(p.numel() for p in model.parameters()) == int(case["n_params"]) == 7043482
    assert not model.fast  # plain PyTorch path: nothing of the CUDA library is involved


def test_ctcnet_forward_matches_the_reference(case, model):
    with torch.no_grad():
        out = model(torch.from_numpy(case["wav"]), torch.from_numpy(case["lip"]))
    assert tuple(out.shape) == tuple(case["out_ref_fp32"].shape) == (1, 1, 16000)
    assert rel_l2(out, torch.from_numpy(case["out_ref_fp32"])) < 1e-5


def test_ctcnet_trains_with_stock_autograd(model):
    m = copy.deepcopy(model).train()
    out = m(0.1 * torch.randn(1, 8000), torch.rand(1, 512, 13))
    out.pow(2).mean().backward()
    assert all(p.grad is not None for p in m.parameters())


def test_factories_follow_the_reference_semantics():
    import torch.nn as nn

    from rtfs_net_b200 import generic as G
    from rtfs_net_b200 import nn as M

    for get in (G.layers_get, G.normalizations_get, G.activations_get, G.encoder_get, G.decoder_get, G.mask_generator_get, G.separators_get, G.fusion_get):
        assert get(None) is nn.Identity
        assert get(nn.ReLU) is nn.ReLU  # callables pass through
        with pytest.raises(ValueError):
            get("NoSuchThing")
        with pytest.raises(ValueError):
            get(3)
    assert G.layers_get("DualPathRNN") is M.DualPathRNN and G.layers_get("MultiHeadSelfAttention2D") is M.MultiHeadSelfAttention2D
    assert G.normalizations_get("gLN") is M.GlobalLayerNorm and G.normalizations_get("BatchNorm1d") is nn.BatchNorm1d
    assert G.activations_get("PReLU") is nn.PReLU
    assert G.encoder_get("STFTEncoder") is M.STFTEncoder and G.encoder_get("ConvolutionalEncoder") is G.ConvolutionalEncoder
    assert G.separators_get("TDANet") is M.TDANet and G.separators_get("FRCNN") is G.FRCNN
    with pytest.raises(NotImplementedError):  # exists in the reference, outside this package's scope
        G.separators_get("DPTNet")
