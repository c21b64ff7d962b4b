#!/usr/bin/env python
"""bench.py -- throughput of the RTFS-Net model-forward hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): utterances/sec (2 s @ 16 kHz, batch 32) per GPU, RTFS-Net 4-layer
inference (configs[1]).  One step = one AVNet forward over a batch of 32 synthetic utterances
(+ 50-frame lip embeddings) per GPU; N GPUs = N independent batch shards (weak scaling, no
data-path collective).  Prints ONE JSON line on rank 0.

  value     device-timed (CUDA events), inputs resident in HBM
  e2e       through the public nn.Module API with pinned HOST inputs: H2D + forward + D2H per step
  roofline  the dominant kernel stage: algorithmic bytes / live CUDA-event duration vs measured HBM peak
  cpu_baseline  the CPU oracle port of the reference forward on the host cores (bounded sample)

--impl reference times the reference's CPU implementation of the path (the oracle port; the Python
reference itself cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

BATCH = 32
SECONDS = 2
L = 16000 * SECONDS
TV = 25 * SECONDS
REPEATS = 4
METRIC = "utterances/sec (2s@16kHz, batch 32) per GPU; SI-SDR within 0.01 dB of ref"
WORKLOAD = "RTFSNet 4-layer inference, batch 32, 2 s @ 16 kHz synthetic, 1xB200"


def load_state_dict():
    g = np.load(os.path.join(ROOT, "tests", "golden", "state_dict_rtfs.npz"))
    return {k: torch.from_numpy(g[k]) for k in g.files}


def make_inputs(batch, seed):
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(batch, L, generator=g), torch.rand(batch, 512, TV, generator=g)


def measured_traffic(stage):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the stage's kernel at B=32 from the committed
    ncu --set full capture (profiles/r02_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")  # ncu capture of this round's kernels (tools/make_traffic.py)
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f)
    v = d.get("stages", {}).get(stage)
    return None if v is None else float(v["dram_bytes"])


def dprnn_flops(B):
    """Algorithmic FLOPs of one fused dual-path RNN launch, averaged over the two launches of a block pass (SURVEY.md
    App. F): a path with O sequences of S positions runs the SRU stack on L = S - 7 unfolded steps, each step
    unfold(8) o Linear 512 -> 256 (k = 4 gates), three layers 64 -> 192 (k = 3: identity highway) and its share of the
    ConvTranspose1d 64 -> 64 x 8 taps; 2 FLOPs per MAC.  Frequency path: O = B*Tc, S = Fc; time path: O = B*Fc, S = Tc.
    (= (933.9 + 3*87.6 + 233.5 + 989.9 + 3*92.8 + 247.5) / 2 MMAC per utterance at 2 s.)"""
    T = L // 128 + 1
    Tc, Fc = (T - 2) // 2 + 1, 64
    per_step = 512 * 256 + 3 * 64 * 192 + 64 * 512
    macs = B * Tc * (Fc - 7) * per_step + B * Fc * (Tc - 7) * per_step
    return 2.0 * macs / 2.0


def measured_tensor_peak():
    """TF32 dense peak: half the measured bf16 cuBLAS rate (sustained figure: the kernel is timed inside a long step)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        v = d.get("bf16_tflops_sustained", d.get("bf16_tflops"))
        if v:
            return float(v) / 2.0, "measured (MEASURED_PEAKS.json bf16_tflops_sustained / 2: tf32 runs at half the bf16 rate)"
    return 1125.0, "fallback (nominal 2.25 PFLOP/s bf16 / 2)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------- algorithmic bytes
def stage_bytes(B):
    """Algorithmic HBM bytes per launch of each kernel stage (fp32).  Stages as launched today:
    DW_S2_POOL = one pass over gLN(d0) producing le0 (H), d1 (G) and the pool (G); DPRNN_FUSED = one dual-path RNN
    (reads g [+ pool], writes g' [+ g]; average of the two paths); TFAR_GLOBAL = the 4-conv launch (5G) and le1 (2G).
    A = 4*256*T*F, H = 4*64*T*F, G = 4*64*Tc*Fc bytes per utterance."""
    T, Fq = L // 128 + 1, 129
    Tc, Fc = (T - 2) // 2 + 1, 64
    A, H, G = 4 * 256 * T * Fq * B, 4 * 64 * T * Fq * B, 4 * 64 * Tc * Fc * B
    R = REPEATS
    # residual conv: reads lec, d0 (2H), gec, ggc (2G), the block input x for the recomputed gateway (A), writes out (A)
    # and -- passes 2..R-1 only -- reads the addend a1 (A).  The last pass has no addend; in the first pass (CAF fused in
    # the epilogue) the addend IS x, read once.  RESID_OUT is timed over passes 2..R: average bytes of those launches.
    resid_mid = (((R - 2) * 3 * A + 2 * A) / (R - 1) if R > 1 else 2 * A) + 2 * H + 2 * G
    return {
        "RTFS_SG_ENC_CONV": A, "RTFS_SG_BOTTLENECK": 2 * A, "RTFS_SG_GATE_PROJ": A + H, "RTFS_SG_DW_S1": 2 * H,
        "RTFS_SG_DW_S2_POOL": 2 * H + 2 * G, "RTFS_SG_DPRNN_FUSED": 3 * G, "RTFS_SG_DPRNN_PREP": 3 * G, "RTFS_SG_DPRNN_GEMM0": 5 * G,
        "RTFS_SG_DPRNN_SCAN": 5 * G, "RTFS_SG_DPRNN_GEMML": 4 * G, "RTFS_SG_DPRNN_CONVT": 3 * G,
        "RTFS_SG_ATT_QKV": 2.5 * G, "RTFS_SG_ATT_CORE": 2.5 * G, "RTFS_SG_ATT_PROJ": 3 * G,
        "RTFS_SG_TFAR_GLOBAL": 3.5 * G, "RTFS_SG_TFAR_LE0": 2 * H, "RTFS_SG_TFAR_CAT_GLOBAL": 5 * G,
        "RTFS_SG_TFAR_CAT_LOCAL": 2 * H + 2 * G, "RTFS_SG_RESID_OUT": resid_mid, "RTFS_SG_RESID_OUT_CAF": 2 * A + 2 * H + 2 * G,
        "RTFS_SG_CAF_APPLY": 3 * A, "RTFS_SG_MASK": 3 * A, "RTFS_SG_DEC_GEMM": A * (1 + 18 / 256), "RTFS_SG_MASK_DEC": A * (2 + 18 / 256),
    }, (4 * A + 14 * H + 36 * G), ((6 + 4 * REPEATS) * A + 14 * REPEATS * H + 36 * REPEATS * G)


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def config_block(world):
    """The `config` object of both arms (identical keys so that the driver can compare them)."""
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH, "repeats": REPEATS, "samples": L, "video_frames": TV,
            "parallelism": f"dp{world} (batch shards, no collective)",
            "l2": "no explicit flush: one step streams ~45 GB through a 7.3 GB working set >> 126 MB L2"}


def check_against_golden(model, dev):
    """One utterance of the committed reference golden case through the model before anything is timed: a bench number of
    a build that no longer matches the reference is worthless."""
    from conftest import load_case, rel_l2

    case = load_case("rtfs4_b1_1s")
    with torch.no_grad():
        out = model(case["wav"].to(dev), case["lip"].to(dev))
    err = rel_l2(out, case["out_ref_fp32"])
    if not err <= 1e-3:
        raise SystemExit(f"bench.py: golden check failed before timing (waveform rel-L2 {err:.3e} > 1e-3)")
    return err


def gpu_eager_rate(dev, n_utt, steps):
    """What a user of the reference gets on this GPU today (BASELINE.md 4.6 / SURVEY.md 8d): the same forward as eager
    PyTorch library ops (cuDNN / cuBLAS / ATen, TF32 matmuls as under set_float32_matmul_precision('high')) -- here the
    oracle port run on the device, because the Python reference cannot travel to the GPU box.  Its SRU recurrence is a
    Python loop over time steps (the upstream `sru` CUDA kernel is not available offline), which is stated in the line."""
    from oracle import rtfs_oracle as O

    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    try:
        sd = {k: v.to(dev) for k, v in load_state_dict().items()}
        wav, lip = make_inputs(n_utt, 1)
        wav, lip = wav.to(dev), lip.to(dev)
        with torch.no_grad():
            O.avnet_forward(sd, wav, lip, REPEATS)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                O.avnet_forward(sd, wav, lip, REPEATS)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return n_utt / (ms * 1e-3), ms


def train_step_block(dev, rank, world, steps=5, warmup=2):
    """BASELINE configs[2]: RTFS-Net-6 training step, batch 16 per GPU, SNR loss, gradient all-reduce (NCCL when world > 1),
    clip + AdamW -- forward with tape, hand-written CUDA backward, one flat all-reduce, fused optimizer kernel
    (rtfs_net_b200/train.py).  Returns the `train_step` object of the JSON line (rank 0) or None."""
    from conftest import audionet_conf
    from rtfs_net_b200 import AVNet, _lib, shard
    from rtfs_net_b200.train import Trainer

    R, B = 6, 16
    model = AVNet(print_macs=False, **audionet_conf(R))
    model.load_state_dict(load_state_dict(), strict=True)
    model = model.to(dev)
    if world > 1:  # train.py:145 sync_batchnorm=True
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    tr = Trainer(model, lr=1e-3, weight_decay=0.1, clip=5.0)
    g = torch.Generator().manual_seed(2000 + rank)
    tgt = (0.1 * torch.randn(B, 1, L, generator=g)).to(dev)
    wav = (tgt[:, 0] + 0.1 * torch.randn(B, L, generator=g).to(dev))
    lip = torch.rand(B, 512, TV, generator=g).to(dev)
    losses = []
    for _ in range(warmup):
        losses.append(float(tr.step(wav, tgt, lip)))
    torch.cuda.synchronize()
    shard.barrier()
    evs = []
    launches0 = 0
    for _ in range(steps):
        ev = []
        losses.append(tr.step(wav, tgt, lip, events=ev))
        evs.append(ev)
    torch.cuda.synchronize()
    shard.barrier()
    total = evs[0][0].elapsed_time(evs[-1][3])
    fwd = sum(e[0].elapsed_time(e[1]) for e in evs) / steps
    bwd = sum(e[1].elapsed_time(e[2]) for e in evs) / steps
    opt = sum(e[2].elapsed_time(e[3]) for e in evs) / steps
    total = shard.max_over_ranks(total, dev)
    losses = [float(x) for x in losses]
    peak_mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    grad_bytes = int(tr.flat_g.numel() * 4)
    del tr, model
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    T, Fq = L // 128 + 1, 129
    Tc, Fc = (T - 2) // 2 + 1, 64
    A, H, G = 4 * 256 * T * Fq * B, 4 * 64 * T * Fq * B, 4 * 64 * Tc * Fc * B
    fwd_bytes = (6 + 4 * R) * A + 14 * R * H + 36 * R * G
    ms = total / steps
    peak, _ = measured_peaks()
    return {"config": f"RTFSNet {R}-layer training step (SNR loss + grad allreduce), batch {B}/GPU, 2 s @ 16 kHz synthetic, dp{world}",
            "ms_per_step": ms, "utterances_per_s": B * world / (ms * 1e-3), "forward_loss_ms": fwd, "backward_ms": bwd, "allreduce_optimizer_ms": opt,
            "steps": steps, "warmup": warmup, "loss_first": losses[0], "loss_last": losses[-1],
            "allreduce": {"backend": "nccl" if world > 1 else None, "bytes_per_step": grad_bytes, "what": "one sum all-reduce of the flat fp32 gradient"},
            "algorithmic_bytes": {"what": "3 x forward bytes ((6+4R)A + 14RH + 36RG): forward + data-gradient + weight-gradient passes", "bytes": 3 * fwd_bytes,
                                  "achieved_gbps": 3 * fwd_bytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": 3 * fwd_bytes / (ms * 1e-3) / 1e9 / peak},
            "peak_memory_gib": peak_mem, "sync_batchnorm": world > 1}


# ------------------------------------------------------------------------------- CPU arm
def cpu_forward_rate(n_utt, steps, warmup):
    """utterances/s of the oracle port (CPU restatement of the reference forward) on the host cores."""
    from oracle import rtfs_oracle as O

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = load_state_dict()
    wav, lip = make_inputs(n_utt, 1)
    with torch.no_grad():
        for _ in range(warmup):
            O.avnet_forward(sd, wav, lip, REPEATS)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.avnet_forward(sd, wav, lip, REPEATS)
        dt = time.perf_counter() - t0
    return n_utt * steps / dt, dt / steps, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    n_utt = 2
    rate, sec_per_step, cores = cpu_forward_rate(n_utt, args.steps, max(1, min(args.warmup, 2)))
    sample = f"{n_utt} of the 32 utterances per step (oracle CPU port of the reference forward, fp32, torch CPU ops + OpenMP SRU scan)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "utterances/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(args.gpus),
        "note": f"each timed step is a forward over {n_utt} utterances (a bounded sample of the batch-32 step); ms_per_step is per "
                f"{n_utt}-utterance step, value = utterances/s is directly comparable",
        "cpu_baseline": {"value": rate, "unit": "utterances/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    from conftest import build_model
    from rtfs_net_b200 import _lib, shard

    rank, local_rank, world = shard.init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the RTFS-Net B200 path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    model = build_model(load_state_dict(), REPEATS, dev)
    golden_err = check_against_golden(model, dev)
    wav_h, lip_h = make_inputs(BATCH, 1000 + rank)
    wav_h, lip_h = wav_h.pin_memory(), lip_h.pin_memory()
    wav, lip = wav_h.to(dev), lip_h.to(dev)
    out_h = torch.empty(BATCH, 1, L).pin_memory()
    K, W = args.steps, args.warmup
    ev = lambda: torch.cuda.Event(enable_timing=True)

    with torch.no_grad():
        for _ in range(W):
            out = model(wav, lip)
        torch.cuda.synchronize()
        launches = int(_lib.lib().rtfs_last_launch_count())

        # ---- timed region 1: device-resident inputs
        sampler = ClockSampler(local_rank)
        sampler.start()
        shard.barrier()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(K):
            out = model(wav, lip)
        e1.record()
        torch.cuda.synchronize()
        shard.barrier()
        ms_total = e0.elapsed_time(e1)
        clocks = sampler.stop()

        # ---- timed region 2: end to end through the public API with host buffers: every step copies its inputs from pinned host
        #      memory and its waveforms back (shard.StreamingSeparator: the copies of neighbouring batches overlap the kernels of the
        #      current one; the last batch's copy back is inside the timed region: e3 is recorded after drain())
        sep = shard.StreamingSeparator(model, BATCH, L, TV, dev)
        for _ in range(2):
            sep.submit(wav_h, lip_h, out_h)
        sep.drain()
        torch.cuda.synchronize()
        shard.barrier()
        e2, e3 = ev(), ev()
        e2.record()
        for _ in range(K):
            sep.submit(wav_h, lip_h, out_h)
        sep.drain()
        e3.record()
        torch.cuda.synchronize()
        shard.barrier()
        ms_e2e = e2.elapsed_time(e3)
        e2e_check = float((out_h.to(dev) - out).abs().max())  # the streamed result is the device-resident one
        del sep

        # ---- per-stage device times (same K steps, CUDA events on the launch stream)
        _lib.profile_enable(True)
        for _ in range(K):
            out = model(wav, lip)
        stages = _lib.profile_collect()
        _lib.profile_enable(False)
    assert torch.isfinite(out).all()

    ms_total = shard.max_over_ranks(ms_total, dev)
    ms_e2e = shard.max_over_ranks(ms_e2e, dev)
    train_block = None
    if not args.no_train:
        del model
        torch.cuda.empty_cache()
        try:
            train_block = train_step_block(dev, rank, world)
        except Exception as e:  # the extra block must never take the headline line down
            train_block = {"unavailable": repr(e)[:300]}
    if rank != 0:
        return
    n_utt = BATCH * world * K
    value = n_utt / (ms_total * 1e-3)
    e2e = n_utt / (ms_e2e * 1e-3)

    peak, peak_src = measured_peaks()
    sbytes, block_bytes, fwd_bytes = stage_bytes(BATCH)
    per_stage = {k: {"ms_per_launch": v[0] / v[1], "launches_per_step": v[1] / K, "ms_per_step": v[0] / K} for k, v in stages.items() if v[1] > 0}
    top = max(per_stage, key=lambda k: per_stage[k]["ms_per_step"])
    for k, v in per_stage.items():
        if k in sbytes:
            v["gbps"] = sbytes[k] / (v["ms_per_launch"] * 1e-3) / 1e9
    achieved = sbytes.get(top, 0.0) / (per_stage[top]["ms_per_launch"] * 1e-3) / 1e9
    block_stage_names = [n for n in _lib.STAGE_NAMES if n not in ("RTFS_SG_STFT", "RTFS_SG_ENC_CONV", "RTFS_SG_BOTTLENECK", "RTFS_SG_CAF_VIDEO",
                                                                   "RTFS_SG_CAF_APPLY", "RTFS_SG_MASK", "RTFS_SG_DEC_GEMM", "RTFS_SG_DEC_ISTFT", "RTFS_SG_MASK_DEC")]
    block_ms = sum(per_stage[n]["ms_per_step"] for n in block_stage_names if n in per_stage) / REPEATS
    stage_sum = sum(v["ms_per_step"] for v in per_stage.values())

    # the fused dual-path RNN keeps its intermediates on chip (3G of HBM traffic): it is bounded by the tensor pipe and the serial
    # recurrence, every other stage by HBM.  `roofline` describes the stage with the largest share of the step, `roofline_hbm` the
    # largest HBM-bound one.
    tensor_stages = ("RTFS_SG_DPRNN_FUSED",)
    top_hbm = max((k for k in per_stage if k not in tensor_stages and k in sbytes), key=lambda k: per_stage[k]["ms_per_step"])

    def roofline_of(st):
        ms_l = per_stage[st]["ms_per_launch"]
        common = {"kernel": st, "ms_per_launch": ms_l, "share_of_step": per_stage[st]["ms_per_step"] / stage_sum}
        if st in tensor_stages:
            flops = dprnn_flops(BATCH)
            tf_peak, tf_src = measured_tensor_peak()
            a = flops / (ms_l * 1e-3) / 1e12
            return {"bound": "tensor", "achieved": a, "peak": tf_peak, "unit": "TFLOP/s", "frac": a / tf_peak, "traffic": measured_traffic(st) if BATCH == 32 else None,
                    "peak_source": tf_src, **common}
        a = sbytes.get(st, 0.0) / (ms_l * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": measured_traffic(st) if BATCH == 32 else None,
                "peak_source": peak_src, **common}

    line = {
        "metric": METRIC, "value": value, "unit": "utterances/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 storage, tf32 tensor-core contractions", "data": "synthetic",
        "config": config_block(world),
        "e2e": {"value": e2e, "unit": "utterances/s", "h2d_bytes_per_step": int(wav_h.numel() * 4 + lip_h.numel() * 4), "d2h_bytes_per_step": int(out_h.numel() * 4),
                "ms_per_step": ms_e2e / K, "api": "rtfs_net_b200.shard.StreamingSeparator (double-buffered copies around model.forward)",
                "max_abs_diff_vs_device_resident": e2e_check},
        "gpu_launches": launches * K,
        "clocks": clocks,
        "roofline": roofline_of(top),
        "roofline_hbm": roofline_of(top_hbm),
        "roofline_block": {"what": "RTFS block pass kernel chain, algorithmic 4A+14H+36G", "ms_per_pass": block_ms, "achieved": block_bytes / (block_ms * 1e-3) / 1e9,
                           "peak": peak, "unit": "GB/s", "frac": block_bytes / (block_ms * 1e-3) / 1e9 / peak},
        "roofline_forward": {"what": "whole forward, algorithmic (6+4R)A+14RH+36RG", "achieved": fwd_bytes / (ms_total / K * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": fwd_bytes / (ms_total / K * 1e-3) / 1e9 / peak},
        "stages": {k.replace("RTFS_SG_", "").lower(): {kk: round(vv, 4) for kk, vv in v.items()} for k, v in per_stage.items()},
    }
    if train_block is not None:
        line["train_step"] = train_block
    line["golden_check"] = {"case": "rtfs4_b1_1s (reference-generated fixture)", "waveform_rel_l2": golden_err, "bound": 1e-3}
    if world == 1 and not args.no_eager:
        try:
            rate, ms = gpu_eager_rate(dev, BATCH, 2)
            line["gpu_eager_baseline"] = {"value": rate, "unit": "utterances/s", "ms_per_step": ms, "batch": BATCH,
                                          "what": "oracle port of the reference forward as eager PyTorch library ops on this GPU, TF32 matmuls; "
                                                  "SRU recurrence = Python loop over time steps (upstream sru CUDA kernel not available offline)"}
        except Exception as e:  # the baseline leg must never take the bench line down
            line["gpu_eager_baseline"] = {"unavailable": repr(e)[:200]}
    if world == 1 and not args.no_cpu:
        rate, sec, cores = cpu_forward_rate(4, 1, 1)
        line["cpu_baseline"] = {"value": rate, "unit": "utterances/s", "cores": cores, "kind": "port",
                                "sample": "4 of the 32 utterances, one forward after one warm-up (oracle CPU port of the reference forward, fp32)"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eager", action="store_true", help="skip the gpu_eager_baseline leg")
    ap.add_argument("--no-train", action="store_true", help="skip the train_step block (RTFS-Net-6 training step, batch 16 per GPU)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", 0)))
        return
    run_gpu(args)
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
