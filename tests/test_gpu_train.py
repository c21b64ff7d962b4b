"""GPU tests of the training step (pytest -m gpu): loss / optimizer kernels against torch, the backward of every stage against
autograd through the CPU oracle (oracle/rtfs_oracle.py is differentiable: plain torch ops, the SRU scan as a torch loop), and the
full AVNet gradient for every parameter.

Tolerances.  The reference trains under torch.set_float32_matmul_precision("high") (train.py:8): its own GPU gradients carry TF32
operand rounding (2^-11 relative per operand).  Ours: forward contractions TF32 (as the inference path), data-gradient GEMMs
TF32, weight-gradient reductions 3xTF32.  Gradients are compared with fp32 CPU autograd of the oracle as relative L2 per tensor;
the bound GRAD_TOL is stated next to each assertion and the measured values go to gpurun_out/train_report.txt.
"""
import os

import pytest
import torch

from conftest import ROOT, audionet_conf, rel_l2

pytestmark = pytest.mark.gpu

BLK = "refinement_module.audio_net.blocks."
REPORT = os.path.join(ROOT, "gpurun_out", "train_report.txt")
# Two regimes (measured, gpurun_out/train_report.txt):
#  * smooth: PReLU slopes set to 1 -- no kink sits behind a TF32 contraction, our gradient and the fp32 oracle's differ by the
#    TF32 operand rounding only (measured 3e-4 .. 8e-4, as the dual-path RNN, which has no kinks, shows with the real weights);
#  * kinked (the real slopes 0.25, ReLU of the mask head): the forward activations differ from the fp32 oracle by ~5e-4 (TF32), so
#    a fraction ~3e-4 of the units behind a contraction sit on the other side of their kink; each such unit changes its local
#    derivative by (1 - slope), which is a relative L2 error of (1 - slope) * sqrt(fraction) ~ 1.3e-2 in the gradient that passes
#    through -- the same discrepancy the reference's own TF32 GPU run has against its CPU run.  Bound 5e-2.
# A parameter whose gradient is small next to the typical gradient of the test (scalar slopes, sums with heavy cancellation)
# is held to the same ABSOLUTE error instead: |ours - ref| <= tol * typical |ref| (typical = median gradient norm of the test).
# In the kinked regime a few small-gradient parameters of the attention query/key convs exceed the bound (their gradient passes
# the soft-max and is a small difference of large terms): at most 1 in 25 parameters may, none beyond 0.25.
GRAD_TOL = 5e-3       # smooth regime: per-parameter gradient, relative L2 vs fp32 oracle autograd (measured 3e-4 .. 3e-3)
GRAD_TOL_KINK = 5e-2  # kinked regime (measured 1e-2 .. 4e-2)
STAGE_TOL = 2e-3      # smooth regime: activation gradients inside one block pass (measured 3e-4 .. 5e-4)
STAGE_TOL_KINK = 3e-2


def report(line):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(line + "\n")
    print(line)


@pytest.fixture(scope="module")
def O():
    from oracle import rtfs_oracle

    return rtfs_oracle


def build_train_model(sd, repeats, device="cuda"):
    """Video-block dropout 0 so that train() mode is deterministic (the oracle has no dropout)."""
    from rtfs_net_b200 import AVNet

    conf = audionet_conf(repeats)
    conf["video_params"]["layers"]["layer_1"]["dropout"] = 0.0
    m = AVNet(print_macs=False, **conf)
    m.load_state_dict(sd, strict=True)
    return m.to(device)


def leaf_sd(sd, dtype=torch.float32):
    return {k: (v.clone().to(dtype).requires_grad_(True) if v.dtype.is_floating_point else v.clone()) for k, v in sd.items()}


def smooth_sd(sd, smooth):
    """smooth: every PReLU slope of the audio block that sits behind a TF32 contraction becomes 1 (identity, no kink)."""
    out = {k: v.clone() for k, v in sd.items()}
    if smooth:
        for k in out:
            if k.startswith(BLK) and (k.endswith("act.weight") or k.endswith("projection.full_layer.4.weight")):
                out[k].fill_(1.0)
    return out


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


# ------------------------------------------------------------------------------------------------- loss / optimizer
def test_snr_loss_kernel(O):
    from rtfs_net_b200.train import pit_snr_loss

    g = torch.Generator().manual_seed(3)
    tgt = 0.1 * torch.randn(4, 1, 16000, generator=g) + 0.01
    est = (tgt + 0.03 * torch.randn(4, 1, 16000, generator=g)).requires_grad_(True)
    ref = O.neg_snr(est, tgt).mean()
    ref.backward()
    e2 = est.detach().cuda().requires_grad_(True)
    ours = pit_snr_loss(e2, tgt.cuda())
    ours.backward()
    report(f"snr loss: ours {float(ours):.6f} ref {float(ref):.6f}; grad rel_l2 {rel_l2(e2.grad, est.grad):.3e}")
    assert abs(float(ours) - float(ref)) < 1e-4
    assert rel_l2(e2.grad, est.grad) < 1e-4


def test_adamw_kernel_matches_torch():
    from rtfs_net_b200 import _lib

    g = torch.Generator().manual_seed(5)
    p0 = torch.randn(10007, generator=g)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3, weight_decay=0.1)
    p = p0.clone().cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    gn = torch.zeros(1, dtype=torch.float64, device="cuda")
    for step in range(1, 4):
        grad = 3.0 * torch.randn(10007, generator=g)  # norm >> 5: the clip is active
        ref_p.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 5.0)
        opt.step()
        gd = grad.cuda()
        _lib.check(_lib.lib().rtfs_adamw_step(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 1e-3, 0.9, 0.999, 1e-8, 0.1, step, 5.0, 1.0,
                                              gn.data_ptr(), torch.cuda.current_stream().cuda_stream), "adamw")
    e = rel_l2(p, ref_p.detach())
    report(f"adamw + clip, 3 steps: rel_l2 {e:.3e}")
    assert e < 1e-6


# ------------------------------------------------------------------------------------------------- stages
def _param_grad_report(tag, ours, ref, tol, prefix_filter=None):
    """Relative L2 per parameter.  Gradients that are zero by symmetry (the LN offsets of the attention keys: a shift common to
    all keys leaves the soft-max unchanged) are compared absolutely against the typical gradient size."""
    keys = [k for k, g in ref.items() if g is not None and k in ours and (prefix_filter is None or k.startswith(prefix_filter))]
    typical = float(torch.tensor([float(ref[k].norm()) for k in keys]).median())
    worst = []
    for k in keys:
        gref = ref[k]
        e = rel_l2(ours[k], gref)
        e_abs = float((ours[k].double().cpu() - gref.double()).norm()) / typical
        ok = e < tol or e_abs < tol
        report(f"{tag} dparam {k:<90} rel_l2={e:.3e} |ref|={float(gref.norm()):.3e} abs/typical={e_abs:.2e}" + ("" if ok else "  <-- FAIL"))
        if not ok:
            worst.append((k, e))
    return worst


@pytest.mark.parametrize("which,dim", [(0, 4), (1, 3)])
def test_dprnn_backward(golden_sd, O, which, dim):
    from rtfs_net_b200.train import ModuleHarness, slot_grads_to_param_grads

    model = build_train_model(golden_sd, 4)
    B, Tc = 2, 63
    g = torch.Generator().manual_seed(7 + which)
    x = torch.randn(B, 64, Tc, 64, generator=g)
    dout = torch.randn(B, 64, Tc, 64, generator=g)
    sd = leaf_sd(golden_sd)
    xr = x.clone().requires_grad_(True)
    ref = O.dual_path_rnn(sd, BLK + f"globalatt.{which}.", xr, dim)
    ref.backward(dout)
    h = ModuleHarness(model, B, 2 * Tc + 1, "cuda")
    out, d_in, sg = h.dprnn(which, nhwc(x).cuda(), nhwc(dout).cuda())
    e_out = rel_l2(out.permute(0, 3, 1, 2).cpu() - x, ref.detach() - x)
    e_in = rel_l2(d_in.permute(0, 3, 1, 2), xr.grad)
    report(f"dprnn[{which}] train-forward update rel_l2={e_out:.3e}; d_input rel_l2={e_in:.3e}")
    pg = slot_grads_to_param_grads(model, {k: v for k, v in sg.items() if k.startswith("RTFS_P_RF_" if which == 0 else "RTFS_P_RT_")}, "cuda")
    worst = _param_grad_report(f"dprnn[{which}]", pg, {k: v.grad for k, v in sd.items() if v.dtype.is_floating_point}, GRAD_TOL, BLK + f"globalatt.{which}.")
    assert e_out < 3e-3 and e_in < STAGE_TOL
    assert not worst, worst


@pytest.mark.parametrize("smooth", [True, False])
def test_mhsa_backward(golden_sd, O, smooth):
    from rtfs_net_b200.train import ModuleHarness, slot_grads_to_param_grads

    golden_sd = smooth_sd(golden_sd, smooth)
    gtol, stol = (GRAD_TOL, STAGE_TOL) if smooth else (GRAD_TOL_KINK, STAGE_TOL_KINK)
    model = build_train_model(golden_sd, 4)
    B, Tc = 2, 63
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 64, Tc, 64, generator=g)
    dout = torch.randn(B, 64, Tc, 64, generator=g)
    sd = leaf_sd(golden_sd)
    xr = x.clone().requires_grad_(True)
    ref = O.mhsa2d(sd, BLK + "globalatt.2.", xr)
    ref.backward(dout)
    h = ModuleHarness(model, B, 2 * Tc + 1, "cuda")
    out, d_in, sg = h.mhsa(nhwc(x).cuda(), nhwc(dout).cuda())
    e_out = rel_l2(out.permute(0, 3, 1, 2).cpu() - x, ref.detach() - x)
    e_in = rel_l2(d_in.permute(0, 3, 1, 2), xr.grad)
    tag = "mhsa[smooth]" if smooth else "mhsa[kinked]"
    allowed = 0 if smooth else 3
    report(f"{tag} train-forward update rel_l2={e_out:.3e}; d_input rel_l2={e_in:.3e}")
    pg = slot_grads_to_param_grads(model, {k: v for k, v in sg.items() if k.startswith("RTFS_P_AT_")}, "cuda")
    worst = _param_grad_report(tag, pg, {k: v.grad for k, v in sd.items() if v.dtype.is_floating_point}, gtol, BLK + "globalatt.2.")
    assert e_out < 3e-3 and e_in < stol
    assert len(worst) <= allowed and all(e < 0.25 for _, e in worst), worst


GRAD_TAPS = [
    # (oracle tap, backward scratch buffer, full resolution?)
    ("p_pre", "RTFS_BW_HT", True), ("d0_pre", "RTFS_BW_HDE", True), ("d1_pre", "RTFS_BW_GDD1N", False),
    ("g0", "RTFS_BW_GDG0", False), ("g1", "RTFS_BW_GDG1", False), ("g2", "RTFS_BW_GDG2", False), ("g3", "RTFS_BW_GDG3", False),
    ("f0", "RTFS_BW_HDF0", True), ("f1", "RTFS_BW_GDF1", False),
    ("f0_g_pre", "RTFS_BW_GT1", False), ("f0_e_pre", "RTFS_BW_GT2", False), ("f1_l_pre", "RTFS_BW_GT3", False),
    ("f1_g_pre", "RTFS_BW_GT4", False), ("f1_e_pre", "RTFS_BW_GT5", False),
]


@pytest.mark.parametrize("T,smooth", [(63, True), (126, True), (126, False)])
def test_block_backward(golden_sd, O, T, smooth):
    """One RTFS block pass: gradient w.r.t. the input, 14 intermediate activation gradients, every block parameter."""
    from rtfs_net_b200.train import ModuleHarness, slot_grads_to_param_grads

    golden_sd = smooth_sd(golden_sd, smooth)
    gtol, stol = (GRAD_TOL, STAGE_TOL) if smooth else (GRAD_TOL_KINK, STAGE_TOL_KINK)
    model = build_train_model(golden_sd, 4)
    B, Fq = 2, 129
    Tc = (T - 2) // 2 + 1
    g = torch.Generator().manual_seed(13)
    x = torch.randn(B, 256, T, Fq, generator=g)
    dout = torch.randn(B, 256, T, Fq, generator=g)
    sd = leaf_sd(golden_sd)
    xr = x.clone().requires_grad_(True)
    taps = {}
    ref = O.rtfs_block(sd, BLK, xr, taps)
    for v in taps.values():
        v.retain_grad()
    ref.backward(dout)
    h = ModuleHarness(model, B, T, "cuda")
    out, d_in, sg = h.block(nhwc(x).cuda(), nhwc(dout).cuda())
    e_out = rel_l2(out.permute(0, 3, 1, 2), ref.detach())
    Tn, T = T, f"{T}{' smooth' if smooth else ' kinked'}"
    report(f"block[T{T}] train-forward rel_l2={e_out:.3e}")
    bad = []
    for tap, buf, full in GRAD_TAPS:
        shape = (B, Tn, Fq, 64) if full else (B, Tc, 64, 64)
        ours = h.scratch_view(buf, shape).permute(0, 3, 1, 2)
        e = rel_l2(ours, taps[tap].grad)
        report(f"block[T{T}] d{tap:<10} rel_l2={e:.3e}")
        if not e < stol:
            bad.append((tap, e))
    e_in = rel_l2(d_in.permute(0, 3, 1, 2), xr.grad)
    report(f"block[T{T}] d_input    rel_l2={e_in:.3e}")
    pg = slot_grads_to_param_grads(model, sg, "cuda")
    worst = _param_grad_report(f"block[T{T}]", pg, {k: v.grad for k, v in sd.items() if v.dtype.is_floating_point}, gtol, BLK)
    assert e_out < 3e-3
    assert not bad, bad
    assert e_in < stol
    assert len(worst) <= (0 if smooth else 6) and all(e < 0.25 for _, e in worst), worst


# ------------------------------------------------------------------------------------------------- whole model
def _full_grad_case(golden_sd, O, repeats, train_mode, B=2, L=16000, Tv=25, seed=17, smooth=False):
    golden_sd = smooth_sd(golden_sd, smooth)
    model = build_train_model(golden_sd, repeats)
    model.train(train_mode)
    g = torch.Generator().manual_seed(seed)
    wav = 0.1 * torch.randn(B, L, generator=g)
    lip = torch.rand(B, 512, Tv, generator=g)
    tgt = 0.1 * torch.randn(B, 1, L, generator=g)
    sd = leaf_sd(golden_sd)
    O.BN_TRAIN = train_mode
    try:
        ref_out = O.avnet_forward(sd, wav, lip, repeats)
        ref_loss = O.neg_snr(ref_out, tgt).mean()
        ref_loss.backward()
    finally:
        O.BN_TRAIN = False
    from rtfs_net_b200.train import pit_snr_loss

    out = model(wav.cuda(), lip.cuda())
    loss = pit_snr_loss(out, tgt.cuda())
    loss.backward()
    tag = f"full[R{repeats} {'train' if train_mode else 'eval'}-BN{' smooth' if smooth else ''}]"
    report(f"{tag} waveform rel_l2={rel_l2(out.detach(), ref_out.detach()):.3e}; loss ours {float(loss):.5f} ref {float(ref_loss):.5f}")
    ours = {k: p.grad for k, p in model.named_parameters()}
    missing = [k for k, v in ours.items() if v is None]
    assert not missing, f"parameters without a gradient (DDP find_unused_parameters=False would fail): {missing}"
    ref = {k: v.grad for k, v in sd.items() if v.dtype.is_floating_point and v.grad is not None}
    worst = _param_grad_report(tag, ours, ref, GRAD_TOL_KINK)
    errs = sorted(rel_l2(ours[k], ref[k]) for k in ref if k in ours)
    report(f"{tag} per-parameter gradient rel_l2: median {errs[len(errs) // 2]:.3e}, 90th percentile {errs[int(0.9 * len(errs))]:.3e}, max {errs[-1]:.3e}")
    assert rel_l2(out.detach(), ref_out.detach()) <= 1e-3
    # the bulk sits at the kink level or below; with B = 2 the batch-statistics BatchNorm1d layers of the (eager torch) video
    # block normalise over as few as 14 values and amplify the TF32 noise of the gradient that reaches them
    assert errs[len(errs) // 2] < 2.5e-2 and errs[int(0.9 * len(errs))] < 7e-2
    return model, sd


def test_full_model_gradients_eval_bn(golden_sd, O):
    """Autograd through AVNet.forward with running-statistics BatchNorm (fine-tuning in eval mode), R = 2."""
    _full_grad_case(golden_sd, O, 2, False)


def test_full_model_gradients_train_mode(golden_sd, O):
    """model.train(): batch-statistics BatchNorm in the CAF cell and the video block; R = 2; also checks the running statistics."""
    model, sd = _full_grad_case(golden_sd, O, 2, True)
    cell = model.refinement_module.crossmodal_fusion.fusion_module.audio_lstm
    bn = cell.key_embed.full_layer[3]
    assert int(bn.num_batches_tracked) == 1
    assert float((bn.running_var - 1.0).abs().max()) > 0  # moved away from the initial (0, 1)


def test_full_model_gradients_r6(golden_sd, O):
    """BASELINE configs[2] depth (RTFS-Net-6), train mode, one step's gradients."""
    _full_grad_case(golden_sd, O, 6, True, B=2, L=16000)


def test_full_model_gradients_smooth(golden_sd, O):
    """R = 2, train mode, PReLU slopes of the audio block set to 1: what is left of the kink effect is the ReLU of the mask head."""
    _full_grad_case(golden_sd, O, 2, True, smooth=True)


def test_trainer_step_decreases_loss(golden_sd):
    """Native step (forward + SNR loss + backward + clip + fused AdamW on the flat buffers): a few steps on one fixed batch."""
    from rtfs_net_b200.train import Trainer

    model = build_train_model(golden_sd, 2)
    tr = Trainer(model, lr=1e-3)
    g = torch.Generator().manual_seed(23)
    tgt = 0.1 * torch.randn(2, 1, 16000, generator=g)
    wav = (tgt[:, 0] + 0.1 * torch.randn(2, 16000, generator=g)).cuda()
    lip = torch.rand(2, 512, 25, generator=g).cuda()
    losses = [float(tr.step(wav, tgt.cuda(), lip)) for _ in range(6)]
    report("trainer losses: " + " ".join(f"{v:.4f}" for v in losses))
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0]


def test_fast_sync_batchnorm_native_path_matches_batchnorm():
    """train.FastSyncBatchNorm on CUDA (torch's fused batch-norm kernels around one collective, no host round trip) in a one-rank
    NCCL group against BatchNorm1d: output, input / parameter gradients, running statistics."""
    import torch.distributed as dist

    from rtfs_net_b200.train import FastSyncBatchNorm, use_fast_sync_batchnorm

    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=0, world_size=1)
        created = True
    try:
        g = torch.Generator().manual_seed(5)
        x0 = (torch.randn(4, 512, 50, generator=g) * 1.5 + 0.3).cuda()
        dy = torch.randn(4, 512, 50, generator=g).cuda()
        net = torch.nn.Sequential(torch.nn.SyncBatchNorm(512)).cuda()
        assert use_fast_sync_batchnorm(net) == 1
        bn = net[0]
        bn.force_sync = True
        ref = torch.nn.BatchNorm1d(512).cuda()
        with torch.no_grad():
            bn.weight.normal_(1.0, 0.2)
            bn.bias.normal_(0.0, 0.2)
            ref.weight.copy_(bn.weight)
            ref.bias.copy_(bn.bias)
        bn.train()
        ref.train()
        x = x0.clone().requires_grad_(True)
        xr = x0.clone().requires_grad_(True)
        y, yr = bn(x), ref(xr)
        y.backward(dy)
        yr.backward(dy)
        errs = dict(y=rel_l2(y, yr), dx=rel_l2(x.grad, xr.grad), dw=rel_l2(bn.weight.grad, ref.weight.grad), db=rel_l2(bn.bias.grad, ref.bias.grad),
                    rm=rel_l2(bn.running_mean, ref.running_mean), rv=rel_l2(bn.running_var, ref.running_var))
        report("fast SyncBatchNorm (native CUDA path) vs BatchNorm1d: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
        assert max(errs.values()) < 1e-5 and int(bn.num_batches_tracked) == int(ref.num_batches_tracked)
    finally:
        if created:
            dist.destroy_process_group()
