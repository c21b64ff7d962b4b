"""Input pipeline: the oracle against the reference-generated fixture (CPU), the device kernels and the
double-buffered host pipeline against the oracle and the fixture (GPU)."""
import os
import random
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import datas_oracle as O  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "datas_small.npz"))


def test_oracle_matches_reference_fixture():
    """tests/golden/datas_small.npz was produced by the reference's own transform.py / normalize_tensor_wav (oracle/make_golden_datas.py)."""
    roi = G["roi"]
    for b in range(roi.shape[0]):
        assert np.array_equal(O.mouth_pipeline(roi[b]), G["val"][b])
        assert np.array_equal(O.mouth_pipeline(roi[b], int(G["off_y"][b]), int(G["off_x"][b]), bool(G["flip"][b])), G["train"][b])
        m, s = O.wav_normalize(G["mix"][b], G["src"][b])
        assert np.abs(m - G["mix_n"][b]).max() < 2e-6 and np.abs(s - G["src_n"][b]).max() < 2e-6
    assert G["flip"].min() == 0 and G["flip"].max() == 1  # the fixture exercises both branches


def test_train_augmentation_draw_order_matches_reference():
    """draw_train_augmentation consumes Python's `random` exactly like RandomCrop + HorizontalFlip (transform.py:120-121,145)."""
    from rtfs_net_b200 import datas

    random.seed(6)  # the seed the fixture's "train" pipeline ran under
    oy, ox, fl = datas.draw_train_augmentation(3, 96, 96)
    assert oy == list(G["off_y"]) and ox == list(G["off_x"]) and fl == list(G["flip"])


@pytest.mark.gpu
def test_mouth_preprocess_kernel_vs_reference_fixture():
    from rtfs_net_b200 import datas

    roi = torch.from_numpy(G["roi"]).cuda()
    val = datas.mouth_preprocess(roi).cpu().numpy()[:, 0]
    # the reference computes in float64 and casts to fp32 afterwards (core.py:89 mouth.type_as(wav)): 1 ulp of fp32 at |x| <= 3.6
    assert np.abs(val - G["val"].astype(np.float32)).max() <= 5e-7
    tr = datas.mouth_preprocess(roi, list(G["off_y"]), list(G["off_x"]), list(G["flip"])).cpu().numpy()[:, 0]
    assert np.abs(tr - G["train"].astype(np.float32)).max() <= 5e-7


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 50, 96, 96), (1, 7, 120, 100), (3, 25, 88, 88)])
def test_mouth_preprocess_kernel_shapes(shape):
    from rtfs_net_b200 import datas

    rng = np.random.default_rng(sum(shape))
    roi = rng.integers(0, 256, size=shape, dtype=np.uint8)
    B, T, H, W = shape
    oy = [int(rng.integers(0, H - 88 + 1)) for _ in range(B)]
    ox = [int(rng.integers(0, W - 88 + 1)) for _ in range(B)]
    fl = [int(rng.integers(0, 2)) for _ in range(B)]
    out = datas.mouth_preprocess(torch.from_numpy(roi).cuda(), oy, ox, fl).cpu().numpy()
    assert out.shape == (B, 1, T, 88, 88)
    for b in range(B):
        ref = O.mouth_pipeline(roi[b], oy[b], ox[b], bool(fl[b])).astype(np.float32)
        assert np.abs(out[b, 0] - ref).max() <= 5e-7
    centre = datas.mouth_preprocess(torch.from_numpy(roi).cuda()).cpu().numpy()
    assert np.abs(centre[0, 0] - O.mouth_pipeline(roi[0]).astype(np.float32)).max() <= 5e-7


@pytest.mark.gpu
def test_wav_normalize_kernel():
    from rtfs_net_b200 import datas

    m, s = datas.wav_normalize(torch.from_numpy(G["mix"]).cuda(), torch.from_numpy(G["src"]).cuda())
    assert np.abs(m.cpu().numpy() - G["mix_n"]).max() < 5e-6 and np.abs(s.cpu().numpy() - G["src_n"]).max() < 5e-6
    # full-size utterances, mixture only
    rng = np.random.default_rng(3)
    mix = (0.1 * rng.standard_normal((4, 32000))).astype(np.float32)
    m2, s2 = datas.wav_normalize(torch.from_numpy(mix).cuda())
    assert s2 is None
    for b in range(4):
        assert np.abs(m2[b].cpu().numpy() - O.wav_normalize(mix[b])[0]).max() < 5e-6
    with pytest.raises(ValueError):
        datas.wav_normalize(torch.from_numpy(mix))  # CPU tensor: no fallback


@pytest.mark.gpu
def test_device_input_pipeline_double_buffered():
    """Ragged last batch, 5 batches through 2 slots: every batch equals the oracle on its own raw data."""
    from rtfs_net_b200 import datas

    rng = np.random.default_rng(5)
    sizes = [4, 4, 4, 4, 2]
    raws = [((0.1 * rng.standard_normal((n, 8000))).astype(np.float32), (0.1 * rng.standard_normal((n, 1, 8000))).astype(np.float32),
             rng.integers(0, 256, size=(n, 10, 96, 96), dtype=np.uint8)) for n in sizes]
    pipe = datas.DeviceInputPipeline(4, 8000, 10, n_src=1, device="cuda", train=False)
    seen = 0
    for (mix, src, roi), (dm, ds, dv) in zip(raws, pipe.run(raws)):
        n = mix.shape[0]
        assert dm.shape == (n, 8000) and ds.shape == (n, 1, 8000) and dv.shape == (n, 1, 10, 88, 88)
        for b in range(n):
            om, os_ = O.wav_normalize(mix[b], src[b])
            assert np.abs(dm[b].cpu().numpy() - om).max() < 5e-6 and np.abs(ds[b].cpu().numpy() - os_).max() < 5e-6
            assert np.abs(dv[b, 0].cpu().numpy() - O.mouth_pipeline(roi[b]).astype(np.float32)).max() <= 5e-7
        seen += 1
    assert seen == len(sizes) and pipe.h2d_bytes == sum(n * (8000 * 4 * 2 + 10 * 96 * 96) for n in sizes)
