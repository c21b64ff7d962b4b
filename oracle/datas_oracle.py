"""TEST INFRASTRUCTURE ONLY (CPU oracle) -- numpy restatement of the reference's per-utterance input arithmetic.

Follows src/datas/transform.py:63-167 (Normalize / CenterCrop / RandomCrop / HorizontalFlip / get_preprocessing_pipelines) and
src/datas/avspeech_dataset.py:10-14,128-131,167-170 (normalize_tensor_wav with the mixture's unbiased std).  Pinned against the
reference's own functions by oracle/make_golden_datas.py (tests/golden/datas_small.npz).
"""
import numpy as np

CROP = 88
MEAN, STD = 0.421, 0.165


def center_offsets(h, w, th=CROP, tw=CROP):
    """transform.py:96-99: int(round(w - tw) / 2.0)."""
    return int(round((h - th)) / 2.0), int(round((w - tw)) / 2.0)


def mouth_pipeline(frames, off_y=None, off_x=None, flip=False, crop=CROP, mean=MEAN, std=STD):
    """frames (T,H,W) uint8 -> (T,crop,crop) float64, as Compose([Normalize(0,255), Crop, [Flip], Normalize(mean,std)])."""
    x = (frames - 0.0) / 255.0                      # transform.py:78 (numpy promotes uint8 to float64)
    t, h, w = x.shape
    if off_y is None:
        off_y, off_x = center_offsets(h, w, crop, crop)
    x = x[:, off_y:off_y + crop, off_x:off_x + crop]  # transform.py:100 / 122
    if flip:
        x = x[:, :, ::-1]                            # cv2.flip(frame, 1), transform.py:146-147
    return (x - mean) / std                          # transform.py:78


def wav_normalize(mix, sources=None, eps=1e-8):
    """mix (L,), sources (n_src, L) float32 -> normalised float32 arrays (avspeech_dataset.py:10-14,128-131)."""
    mix = np.asarray(mix, dtype=np.float32)
    std = mix.astype(np.float64).std(ddof=1)        # torch.std: unbiased
    m = (mix - mix.mean(dtype=np.float64)) / (std + eps)
    s = None
    if sources is not None:
        sources = np.asarray(sources, dtype=np.float32)
        s = (sources - sources.mean(-1, keepdims=True, dtype=np.float64)) / (std + eps)
    return m.astype(np.float32), (None if s is None else s.astype(np.float32))
