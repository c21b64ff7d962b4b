"""Generates tests/golden/datas_small.npz by running the REFERENCE's own input arithmetic (authoring container only):
`get_preprocessing_pipelines()` of /root/reference/src/datas/transform.py ("val" and, seeded, "train") and `normalize_tensor_wav`
of src/datas/avspeech_dataset.py on small seeded inputs, and asserts that oracle/datas_oracle.py reproduces them.

    python oracle/make_golden_datas.py

cv2 and soundfile are absent here: two-line stand-ins cover what the imported code uses (cv2.flip(frame, 1) = horizontal flip).
"""
import os
import random
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

cv2 = types.ModuleType("cv2")
cv2.flip = lambda img, code: img[:, ::-1].copy() if code == 1 else img[::-1].copy()
cv2.COLOR_RGB2GRAY = 7
cv2.cvtColor = lambda img, code: img
sys.modules["cv2"] = cv2
sys.modules["soundfile"] = types.ModuleType("soundfile")
sys.path.insert(0, REF)
from src.datas.transform import get_preprocessing_pipelines  # noqa: E402
from src.datas.avspeech_dataset import normalize_tensor_wav  # noqa: E402

from oracle import datas_oracle as O  # noqa: E402

rng = np.random.default_rng(7)
B, T, H, W = 3, 2, 96, 96
roi = rng.integers(0, 256, size=(B, T, H, W), dtype=np.uint8)
pipes = get_preprocessing_pipelines()

val = np.stack([pipes["val"](roi[b]) for b in range(B)])
for b in range(B):
    assert np.array_equal(val[b], O.mouth_pipeline(roi[b]))

# "train": record the decisions the reference's transforms draw (same call order as rtfs_net_b200.datas.draw_train_augmentation)
random.seed(6)
train = np.stack([pipes["train"](roi[b].copy()) for b in range(B)])
random.seed(6)
off_y, off_x, flip = [], [], []
for b in range(B):
    off_x.append(random.randint(0, W - 88))
    off_y.append(random.randint(0, H - 88))
    flip.append(1 if random.random() < 0.5 else 0)
for b in range(B):
    assert np.array_equal(train[b], O.mouth_pipeline(roi[b], off_y[b], off_x[b], bool(flip[b]))), b

L, n_src = 4000, 2
mix = (0.1 * rng.standard_normal((B, L))).astype(np.float32) + 0.03
src = (0.07 * rng.standard_normal((B, n_src, L))).astype(np.float32) - 0.01
mix_n, src_n = [], []
for b in range(B):
    m = torch.from_numpy(mix[b])
    s = torch.from_numpy(src[b])
    m_std = m.std(-1, keepdim=True)
    mix_n.append(normalize_tensor_wav(m, eps=1e-8, std=m_std).numpy())
    src_n.append(normalize_tensor_wav(s, eps=1e-8, std=m_std).numpy())
    om, os_ = O.wav_normalize(mix[b], src[b])
    assert np.abs(om - mix_n[-1]).max() < 2e-6 and np.abs(os_ - src_n[-1]).max() < 2e-6

out = os.path.join(ROOT, "tests", "golden", "datas_small.npz")
np.savez_compressed(out, roi=roi, val=val.astype(np.float64), train=train.astype(np.float64), off_y=np.array(off_y), off_x=np.array(off_x),
                    flip=np.array(flip), mix=mix, src=src, mix_n=np.stack(mix_n), src_n=np.stack(src_n))
print("wrote", out, os.path.getsize(out), "bytes; decisions", off_y, off_x, flip)
