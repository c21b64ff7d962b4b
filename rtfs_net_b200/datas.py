"""Input pipeline of the RTFS-Net path with the per-utterance arithmetic on the device.

Reference: `AVSpeechDataset.__getitem__` (src/datas/avspeech_dataset.py:115-185) decodes a wav + a mouth-ROI `.npz` per utterance
and runs `normalize_tensor_wav` (avspeech_dataset.py:10-14) and the lip-reading transforms of `get_preprocessing_pipelines()`
(src/datas/transform.py:151-167) in numpy inside the DataLoader workers, shipping float64 (T, 88, 88) frames to the GPU.  Here the
host only stages the RAW data -- uint8 ROIs (4.6x smaller than the cropped fp32 frames, 9x smaller than the float64 ones) and fp32
waveforms -- in pinned memory, two batches deep; one asynchronous copy each on a copy stream, then two kernels
(csrc/input.cuh: `rtfs_mouth_preprocess`, `rtfs_wav_normalize`) produce what the reference's collate would have produced.

The random crop / flip decisions of the "train" pipeline are drawn on the host with Python's `random` in the reference's order
(RandomCrop: delta_w then delta_h, transform.py:120-121; HorizontalFlip: one `random.random()`, transform.py:145), so a run seeded
like the reference makes the same decisions.
"""
import random

import torch

from . import _lib

CROP = 88                    # transform.py:153
MEAN, STD = 0.421, 0.165     # transform.py:154
EPS = 1e-8                   # avspeech_dataset.py:116


def draw_train_augmentation(n, H, W, crop=CROP, flip_ratio=0.5):
    """Crop offsets and flip decisions of `n` utterances, in the reference's call order (transform.py:120-121,145)."""
    off_x, off_y, flip = [], [], []
    for _ in range(n):
        off_x.append(random.randint(0, W - crop))
        off_y.append(random.randint(0, H - crop))
        flip.append(1 if random.random() < flip_ratio else 0)
    return off_y, off_x, flip


def _stream():
    return torch.cuda.current_stream().cuda_stream


def mouth_preprocess(roi, off_y=None, off_x=None, flip=None, crop=CROP, mean=MEAN, std=STD):
    """roi (B,T,H,W) uint8 cuda -> (B,1,T,crop,crop) fp32: the "val" pipeline (centre crop) or, with offsets / flips, "train"."""
    if roi.dtype != torch.uint8 or not roi.is_cuda or roi.ndim != 4:
        raise ValueError("mouth_preprocess expects a (B,T,H,W) uint8 CUDA tensor")
    roi = roi.contiguous()
    B, T, H, W = roi.shape
    out = torch.empty(B, 1, T, crop, crop, device=roi.device, dtype=torch.float32)

    def dev(v):
        return None if v is None else torch.as_tensor(v, dtype=torch.int32).to(roi.device, non_blocking=True).contiguous()

    oy, ox, fl = dev(off_y), dev(off_x), dev(flip)
    with torch.cuda.device(roi.device):
        _lib.check(_lib.lib().rtfs_mouth_preprocess(roi.data_ptr(), out.data_ptr(), None if oy is None else oy.data_ptr(),
                                                    None if ox is None else ox.data_ptr(), None if fl is None else fl.data_ptr(),
                                                    B, T, H, W, crop, mean, std, _stream()), "rtfs_mouth_preprocess")
    return out


def wav_normalize(mix, sources=None, eps=EPS):
    """mix (B,L), sources (B,n_src,L) fp32 cuda -> normalised copies (avspeech_dataset.py:128-131 / 167-170)."""
    if not mix.is_cuda or mix.dtype != torch.float32 or mix.ndim != 2:
        raise ValueError("wav_normalize expects a (B,L) fp32 CUDA mixture")
    mix = mix.contiguous()
    B, L = mix.shape
    n_src = 0
    src_out = None
    if sources is not None:
        sources = sources.contiguous()
        if sources.shape[0] != B or sources.shape[-1] != L or sources.ndim != 3:
            raise ValueError("sources must be (B,n_src,L)")
        n_src = sources.shape[1]
        src_out = torch.empty_like(sources)
    mix_out = torch.empty_like(mix)
    with torch.cuda.device(mix.device):
        _lib.check(_lib.lib().rtfs_wav_normalize(mix.data_ptr(), None if sources is None else sources.data_ptr(), mix_out.data_ptr(),
                                                 None if src_out is None else src_out.data_ptr(), B, L, n_src, eps, _stream()), "rtfs_wav_normalize")
    return mix_out, src_out


class DeviceInputPipeline:
    """Double-buffered pinned-memory staging + device-side preprocessing for batches of raw utterances.

        pipe = DeviceInputPipeline(batch, samples, frames, n_src=1, device="cuda", train=False)
        for mixture, sources, mouth in pipe.run(iterable_of_raw_batches):   # mixture (B,L), sources (B,n_src,L), mouth (B,1,T,88,88)
            ...

    A raw batch is (mix float32 (B,L), sources float32 (B,n_src,L) or None, roi uint8 (B,T,H,W)) as numpy arrays or CPU tensors.
    While the consumer works on batch i on the compute stream, batch i+1 is copied on the copy stream; an event per slot keeps a
    pinned buffer from being overwritten before its copy has finished.
    """

    def __init__(self, batch, samples, frames, roi_hw=(96, 96), n_src=1, device="cuda", train=False, normalize_audio=True):
        self.device = torch.device(device)
        self.train = train
        self.normalize_audio = normalize_audio
        self.n_src = n_src
        H, W = roi_hw
        self.slots = []
        for _ in range(2):
            self.slots.append({
                "mix": torch.empty(batch, samples, dtype=torch.float32).pin_memory(),
                "src": torch.empty(batch, n_src, samples, dtype=torch.float32).pin_memory() if n_src else None,
                "roi": torch.empty(batch, frames, H, W, dtype=torch.uint8).pin_memory(),
                "d_mix": torch.empty(batch, samples, dtype=torch.float32, device=self.device),
                "d_src": torch.empty(batch, n_src, samples, dtype=torch.float32, device=self.device) if n_src else None,
                "d_roi": torch.empty(batch, frames, H, W, dtype=torch.uint8, device=self.device),
                "copied": torch.cuda.Event(),
                "consumed": torch.cuda.Event(),
                "aug": None,
                "n": 0,
            })
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.h2d_bytes = 0

    def _stage(self, slot, raw):
        mix, src, roi = (torch.as_tensor(x) if x is not None else None for x in raw)
        n = mix.shape[0]
        slot["consumed"].synchronize()  # the kernels that read this slot's device buffers have finished
        slot["mix"][:n].copy_(mix)
        if src is not None and slot["src"] is not None:
            slot["src"][:n].copy_(src)
        slot["roi"][:n].copy_(roi)
        slot["n"] = n
        slot["aug"] = draw_train_augmentation(n, roi.shape[-2], roi.shape[-1]) if self.train else None
        with torch.cuda.stream(self.copy_stream):
            slot["d_mix"][:n].copy_(slot["mix"][:n], non_blocking=True)
            if slot["src"] is not None:
                slot["d_src"][:n].copy_(slot["src"][:n], non_blocking=True)
            slot["d_roi"][:n].copy_(slot["roi"][:n], non_blocking=True)
            slot["copied"].record(self.copy_stream)
        self.h2d_bytes += n * (mix.shape[1] * 4 * (1 + (self.n_src if src is not None else 0)) + roi[0].numel())

    def _finish(self, slot):
        n = slot["n"]
        torch.cuda.current_stream(self.device).wait_event(slot["copied"])
        mix, src = slot["d_mix"][:n], (slot["d_src"][:n] if slot["d_src"] is not None else None)
        if self.normalize_audio:
            mix, src = wav_normalize(mix, src)
        else:
            mix, src = mix.clone(), (src.clone() if src is not None else None)
        aug = slot["aug"]
        mouth = mouth_preprocess(slot["d_roi"][:n], *(aug if aug is not None else (None, None, None)))
        slot["consumed"].record(torch.cuda.current_stream(self.device))
        return mix, src, mouth

    def run(self, raw_batches):
        it = iter(raw_batches)
        cur = 0
        try:
            self._stage(self.slots[cur], next(it))
        except StopIteration:
            return
        while True:
            nxt = None
            try:
                nxt = next(it)
            except StopIteration:
                pass
            if nxt is not None:
                self._stage(self.slots[cur ^ 1], nxt)  # copy of batch i+1 overlaps the kernels of batch i
            yield self._finish(self.slots[cur])
            if nxt is None:
                return
            cur ^= 1
