"""Training step of the RTFS-Net path on B200: forward with a tape, hand-written CUDA backward, SNR loss, gradient
all-reduce, clip + AdamW (BASELINE configs[2]; reference: src/system/core.py:94-117, train.py:81-146,
src/losses/matrix.py:22-53, src/losses/pit_wrapper.py:26-51, src/system/optimizers.py:58-75).

Two ways in:

* drop-in: `AVNet.forward` in train() mode (or with autograd enabled) returns a tensor whose grad_fn is `_AVNetFn` below, so the
  reference's `System.training_step` -> `loss.backward()` -> Lightning DDP all-reduce -> `optimizer.step()` run unchanged; every
  parameter receives a gradient (DDP find_unused_parameters=False);
* native: `Trainer.step(wav, target, mouth)` = forward + on-device SNR loss + backward + ONE NCCL all-reduce of the flat 2.96 MB
  gradient + fused clip/AdamW kernel over the flat parameter buffer.

The parameter slots the kernels read are differentiable torch views/permutes of the live parameters (weights.prepare(train=True)):
the CUDA backward returns gradients per slot and autograd carries them back to the parameters.  The 1-D video block is eager
torch (library ops, as in the inference path) and differentiates through autograd; the gradient w.r.t. its output comes from the
CAF backward kernels.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .weights import DERIVED, prepare

CAF = "refinement_module.crossmodal_fusion.fusion_module.audio_lstm."
_MK_PERM = None


def _mk_perm(device):
    """mask head: slot row r holds natural output channel perm[r] (weights.py: interleaved real/imag rows)."""
    global _MK_PERM
    if _MK_PERM is None or _MK_PERM.device != device:
        _MK_PERM = torch.stack([torch.arange(128), torch.arange(128) + 128], 1).reshape(-1).to(device)
    return _MK_PERM


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _event():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _ptr(t):
    return None if t is None else t.data_ptr()


def _table(tensors):
    return (ctypes.c_void_p * len(tensors))(*[_ptr(t) for t in tensors])


class GradBuffers:
    """One zeroed flat buffer with a view per slot that receives a gradient (every non-derived slot)."""

    def __init__(self, slots, device):
        self.names = [n for n in _lib.PARAM_NAMES if n not in DERIVED and slots[n] is not None]
        sizes = [slots[n].numel() for n in self.names]
        self.flat = torch.zeros(sum(sizes), device=device, dtype=torch.float32)
        self.views = {}
        o = 0
        for n, s in zip(self.names, sizes):
            self.views[n] = self.flat[o:o + s].view(slots[n].shape)
            o += s
        self.table = _table([self.views.get(n) for n in _lib.PARAM_NAMES])

    def slot_order(self):
        """Gradients in the layout of the slots (the mask head comes back in natural channel order)."""
        out = dict(self.views)
        perm = _mk_perm(self.flat.device)
        out["RTFS_P_MK_W"] = out["RTFS_P_MK_W"][perm]
        out["RTFS_P_MK_B"] = out["RTFS_P_MK_B"][perm]
        return out


class TrainBuffers:
    """Tape + backward scratch for one (B, L, Tv, R) geometry (allocated once, reused every step)."""

    def __init__(self, B, L, Tv, R, device):
        self.key = (B, L, Tv, R, str(device))
        self.plan = _lib.train_plan(B, L, Tv, R)
        self.tape = torch.empty(self.plan["tape_bytes"], dtype=torch.uint8, device=device)
        self.scratch = torch.empty(self.plan["bwd_bytes"], dtype=torch.uint8, device=device)

    def tape_view(self, name, nbytes, dtype):
        o = self.plan["tape_offsets"][name]
        return self.tape[o:o + nbytes].view(dtype)

    def scratch_view(self, name, nbytes, dtype):
        o = self.plan["bwd_offsets"][name]
        return self.scratch[o:o + nbytes].view(dtype)

    def pass_view(self, i, name, shape):
        """Buffer `name` (enum rtfs_ws) of block pass i as a float tensor (diagnostics / tests)."""
        n = 1
        for s in shape:
            n *= s
        o = self.plan["tape_offsets"]["RTFS_TP_PASS0"] + i * self.plan["pass_bytes"] + self.plan["pass_offsets"][name]
        return self.tape[o:o + 4 * n].view(torch.float32).view(*shape)


def _bn_modules(model):
    cell = model.refinement_module.crossmodal_fusion.fusion_module.audio_lstm
    return cell.key_embed.full_layer[3], cell.value_embed.full_layer[3]


def _sync_bn(bn):
    return isinstance(bn, torch.nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class _AVNetFn(torch.autograd.Function):
    """Audio path of AVNet.forward (tdavnet.py:86-97) for autograd: inputs = mixture, video-block output, the six CAF BatchNorm
    branch parameters and the prepared parameter slots."""

    @staticmethod
    def forward(ctx, rt, wav, video, wk, gk, bk, wv, gv, bv, *slots):
        model = rt.model
        B, L = wav.shape
        Tv = video.shape[-1]
        R = model.refinement_module.audio_params["repeats"]
        dev = wav.device
        bufs = rt.train_buffers(B, L, Tv, R, dev)
        names = _lib.PARAM_NAMES
        slot = dict(zip(names, slots))
        table = _table(slots)
        out = torch.empty(B, L, device=dev, dtype=torch.float32)
        lib = _lib.lib()
        video = video.contiguous()
        with torch.cuda.device(dev):
            early = getattr(rt, "_audio_phase0", None)
            rt._audio_phase0 = None
            if early is not None and early[0] == (B, L, Tv, R, wav.data_ptr(), bufs.tape.data_ptr()):
                # forward_train enqueued the audio-only part before the video block ran (and keeps its slot tensors alive in `early`)
                _lib.check(lib.rtfs_avnet_train_forward(table, wav.data_ptr(), video.data_ptr(), out.data_ptr(), bufs.tape.data_ptr(),
                                                        B, L, Tv, R, 2, _stream()), "rtfs_avnet_train_forward(2)")
            else:
                _lib.check(lib.rtfs_avnet_train_forward(table, wav.data_ptr(), video.data_ptr(), out.data_ptr(), bufs.tape.data_ptr(),
                                                        B, L, Tv, R, 0, _stream()), "rtfs_avnet_train_forward(0)")
            # ---- CAF BatchNorm statistics (layers/fusion.py:210-228): y = w*a per channel => mean_y = w*mean_a, var_y = w^2*var_a
            T = L // 128 + 1
            sums_l = bufs.tape_view("RTFS_TP_CAFSUM", 256 * 2 * 8, torch.float64).view(256, 2).clone()
            n_l = float(B * T * 129)
            bn_k, bn_v = _bn_modules(model)
            training = model.training
            sync = training and _sync_bn(bn_k)
            sums_g, n_g = sums_l, n_l
            if sync:  # SyncBatchNorm (train.py:145): the statistics are those of the global batch
                pack = torch.cat([sums_l.reshape(-1), torch.tensor([n_l], dtype=torch.float64, device=dev)])
                dist.all_reduce(pack)
                sums_g, n_g = pack[:-1].view(256, 2), pack[-1]  # n_g stays on the device: no host sync in the middle of the forward
            mean = sums_g[:, 0] / n_g
            var = (sums_g[:, 1] / n_g - mean * mean).clamp_(min=0.0)
            stats = {}
            for tag, bn, w, gm, be in (("K", bn_k, wk, gk, bk), ("V", bn_v, wv, gv, bv)):
                w64, g64, b64 = w.detach().double().reshape(-1), gm.detach().double(), be.detach().double()
                if training:
                    mu_y, var_y = w64 * mean, w64 * w64 * var
                    if bn.track_running_stats and bn.running_mean is not None:
                        mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
                        with torch.no_grad():
                            bn.running_mean.mul_(1 - mom).add_(mom * mu_y.float())
                            unbias = n_g / torch.clamp(n_g - 1.0, min=1.0) if torch.is_tensor(n_g) else n_g / max(n_g - 1.0, 1.0)
                            bn.running_var.mul_(1 - mom).add_(mom * (var_y * unbias).float())
                            bn.num_batches_tracked += 1
                else:
                    mu_y, var_y = bn.running_mean.double(), bn.running_var.double()
                sig = torch.sqrt(var_y + bn.eps)
                s = g64 * w64 / sig
                tt = b64 - g64 * mu_y / sig
                slot[f"RTFS_P_CAF_S{tag}"].copy_(s.float())
                slot[f"RTFS_P_CAF_T{tag}"].copy_(tt.float())
                stats[tag] = (w64, g64, sig, mu_y)
            _lib.check(lib.rtfs_avnet_train_forward(table, wav.data_ptr(), video.data_ptr(), out.data_ptr(), bufs.tape.data_ptr(),
                                                    B, L, Tv, R, 1, _stream()), "rtfs_avnet_train_forward(1)")
        ctx.rt, ctx.bufs, ctx.geom = rt, bufs, (B, L, Tv, R)
        ctx.slots, ctx.table = slots, table
        ctx.wav, ctx.video = wav, video
        ctx.bn = dict(training=training, sync=sync, mean=mean, var=var, sums_l=sums_l, n_l=n_l, n_g=n_g, stats=stats)
        return out

    @staticmethod
    def backward(ctx, d_out):
        B, L, Tv, R = ctx.geom
        bufs, slots, table = ctx.bufs, ctx.slots, ctx.table
        dev = d_out.device
        slot = dict(zip(_lib.PARAM_NAMES, slots))
        grads = GradBuffers(slot, dev)
        d_out = d_out.contiguous().view(B, L)
        d_video = torch.empty_like(ctx.video)
        lib = _lib.lib()
        bn = ctx.bn
        with torch.cuda.device(dev):
            _lib.check(lib.rtfs_avnet_backward(table, grads.table, ctx.wav.data_ptr(), ctx.video.data_ptr(), d_out.data_ptr(), d_video.data_ptr(),
                                               None, None, None, bufs.tape.data_ptr(), bufs.scratch.data_ptr(), B, L, Tv, R, 0, _stream()),
                       "rtfs_avnet_backward(0)")
            # ---- BatchNorm backward of the two CAF branches from the channel sums (S1 = sum g, S2 = sum g*a)
            cs_l = bufs.scratch_view("RTFS_BW_CSUM", 256 * 4 * 8, torch.float64).view(256, 4).clone()
            cs_g = cs_l
            if bn["sync"]:
                cs_g = cs_l.clone()
                dist.all_reduce(cs_g)
            mean, var, n_g = bn["mean"], bn["var"], bn["n_g"]
            sum_a_l, sum_a2_l = bn["sums_l"][:, 0], bn["sums_l"][:, 1]
            c0 = torch.zeros(256, dtype=torch.float64, device=dev)
            c1 = torch.zeros(256, dtype=torch.float64, device=dev)
            pg = {}
            for col, tag in ((0, "K"), (2, "V")):
                w, g, sig, mu_y = bn["stats"][tag]
                s = g * w / sig
                S1_l, S2_l, S1_g, S2_g = cs_l[:, col], cs_l[:, col + 1], cs_g[:, col], cs_g[:, col + 1]
                if bn["training"]:
                    m1 = S1_g / n_g
                    m2 = (w / sig) * (S2_g - mean * S1_g) / n_g  # mean of g * yhat
                    c0 += s * m1
                    c1 += s * m2 * w / sig
                    d_beta = S1_l
                    d_gamma = (w / sig) * (S2_l - mean * S1_l)
                    d_w = (g / sig) * (S2_l - m1 * sum_a_l - m2 * (w / sig) * (sum_a2_l - mean * sum_a_l))
                else:  # running statistics: the normalisation is a constant affine map
                    d_beta = S1_l
                    d_gamma = (w * S2_l - mu_y * S1_l) / sig
                    d_w = g * S2_l / sig
                pg[tag] = (d_w.float(), d_gamma.float(), d_beta.float())
            mu32, c032, c132 = mean.float().contiguous(), c0.float().contiguous(), c1.float().contiguous()
            _lib.check(lib.rtfs_avnet_backward(table, grads.table, ctx.wav.data_ptr(), ctx.video.data_ptr(), d_out.data_ptr(), d_video.data_ptr(),
                                               mu32.data_ptr(), c032.data_ptr(), c132.data_ptr(), bufs.tape.data_ptr(), bufs.scratch.data_ptr(),
                                               B, L, Tv, R, 1, _stream()), "rtfs_avnet_backward(1)")
        by_slot = grads.slot_order()
        slot_grads = [by_slot.get(n) if (slot[n] is not None and slot[n].requires_grad) else None for n in _lib.PARAM_NAMES]
        wk_shape = (256, 1, 1, 1)
        return (None, None, d_video, pg["K"][0].view(wk_shape), pg["K"][1], pg["K"][2], pg["V"][0].view(wk_shape), pg["V"][1], pg["V"][2], *slot_grads)


def live_tensors(model):
    """{reference key: live parameter / buffer} (no detach: weights.prepare(train=True) differentiates through them)."""
    d = dict(model.named_parameters())
    d.update(dict(model.named_buffers()))
    return d


def _param_versions(model):
    return tuple(p._version for p in model.parameters())


def prepared_slots(rt, device, stash=False):
    """(live tensors, parameter slots) of the training path.  weights.prepare(train=True) re-derives every weight image from the live
    parameters with ~400 small torch ops: 6.5 ms of HOST time per step (tools/prep_time.py).  Trainer.step therefore calls this
    with stash=True right after it has enqueued the optimizer kernel -- the host is ahead of the GPU there, the ops queue behind the
    update they depend on -- and the next forward picks the result up, provided no parameter was touched in between (tensor
    versions; the fused optimizer writes through raw pointers and does not bump them)."""
    model = rt.model
    key = (_param_versions(model), str(device), torch.is_grad_enabled())
    cached = getattr(rt, "_prepared_next", None)
    rt._prepared_next = None
    if not stash and cached is not None and cached[0] == key:
        return cached[1], cached[2]
    live = live_tensors(model)
    slots = prepare(live, device, train=True)
    if stash:
        rt._prepared_next = (key, live, slots)
    return live, slots


def forward_train(rt, wav, mouth):
    """AVNet.forward on the training path (called by nn._Runtime.forward when the model trains / autograd is on)."""
    model = rt.model
    if wav.ndim == 1:
        wav = wav[None]
    elif wav.ndim == 3:
        wav = wav[:, 0]
    wav = wav.contiguous()
    rm = model.refinement_module
    live, slots = prepared_slots(rt, wav.device)
    # The audio-only part of the forward (encoder, bottleneck, first block pass) is enqueued BEFORE the video block: that block is
    # ~150 eager library launches (more under SyncBatchNorm) and bound by the host, which now works while the GPU is busy.
    B, L, Tv = wav.shape[0], wav.shape[1], mouth.shape[-1]
    R = rm.audio_params["repeats"]
    bufs = rt.train_buffers(B, L, Tv, R, wav.device)
    slot_list = [slots[n] for n in _lib.PARAM_NAMES]
    with torch.cuda.device(wav.device):
        _lib.check(_lib.lib().rtfs_avnet_train_forward(_table(slot_list), wav.data_ptr(), None, None, bufs.tape.data_ptr(), B, L, Tv, R, 0, _stream()),
                   "rtfs_avnet_train_forward(0, audio)")
    rt._audio_phase0 = ((B, L, Tv, R, wav.data_ptr(), bufs.tape.data_ptr()), slot_list)
    video = rm.video_net.get_block(0)(model.video_bottleneck(mouth.contiguous()))  # eager torch ops + autograd
    q = CAF
    bnp = [live[q + "key_embed.full_layer.2.weight"], live[q + "key_embed.full_layer.3.weight"], live[q + "key_embed.full_layer.3.bias"],
           live[q + "value_embed.full_layer.2.weight"], live[q + "value_embed.full_layer.3.weight"], live[q + "value_embed.full_layer.3.bias"]]
    out = _AVNetFn.apply(rt, wav, video, *bnp, *slot_list)
    return out.view(wav.shape[0], 1, wav.shape[1])


# ------------------------------------------------------------------------------------------------------------ loss
def snr_loss(est, target, need_grad=True):
    """PITLossWrapper(pairwise_neg_snr, pit_from='pw_mtx') for n_src = 1 (train.py:99): mean over the batch of the negative SNR
    of zero-meaned signals.  Returns (loss scalar tensor, d loss / d est or None) from one pair of kernels."""
    B, L = est.shape[0], est.shape[-1]
    e, t = est.reshape(B, L).contiguous(), target.reshape(B, L).contiguous()
    loss = torch.empty(B, device=e.device, dtype=torch.float32)
    sums = torch.empty(B * 5, device=e.device, dtype=torch.float64)
    d_est = torch.empty_like(e) if need_grad else None
    with torch.cuda.device(e.device):
        _lib.check(_lib.lib().rtfs_snr_loss(e.data_ptr(), t.data_ptr(), loss.data_ptr(), _ptr(d_est), sums.data_ptr(), B, L, 1.0 / B, _stream()), "rtfs_snr_loss")
    return loss.mean(), d_est


class _SnrLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, est, target):
        loss, d_est = snr_loss(est.detach(), target, True)
        ctx.save_for_backward(d_est)
        ctx.shape = est.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        (d_est,) = ctx.saved_tensors
        return (d_est * g).view(ctx.shape), None


def pit_snr_loss(est, target):
    """Differentiable drop-in for loss_func["train"] of the reference (n_src = 1)."""
    return _SnrLossFn.apply(est, target)


# ------------------------------------------------------------------------------------------------------------ trainer
class FlatParams:
    """Parameters and gradients of a module as views of two flat fp32 buffers (device-agnostic: the CPU gloo tests cover the
    data-parallel exchange).  Every parameter starts on a 256-byte boundary: the kernels read parameters with 16-byte vector
    loads; the padding stays zero (zero gradient, zero moments, decay of zero)."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        dev = self.params[0].device
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 63) // 64 * 64
        self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                k = p.numel()
                self.flat_p[o:o + k].copy_(p.reshape(-1))
                p.data = self.flat_p[o:o + k].view(p.shape)
                p.grad = self.flat_g[o:o + k].view(p.shape)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def exchange(self):
        """One sum all-reduce of the flat gradient (NCCL over NVLink on the GPU box, gloo in the CPU tests); the 1/world factor
        is folded into the optimizer kernel."""
        if self.world > 1:
            dist.all_reduce(self.flat_g)


class _SyncBNFn(torch.autograd.Function):
    """Batch normalisation over the GLOBAL batch with one small all-reduce each way and no host synchronisation."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, group):
        C = x.shape[1]
        dims = [0] + list(range(2, x.dim()))
        shape = [1, C] + [1] * (x.dim() - 2)
        xd = x.double()
        pack = torch.cat([xd.sum(dims), (xd * xd).sum(dims), xd.new_tensor([x.numel() / C])])
        dist.all_reduce(pack, group=group)
        n = pack[-1]
        mean = pack[:C] / n
        var = (pack[C:2 * C] / n - mean * mean).clamp_(min=0.0)
        rstd = torch.rsqrt(var + eps)
        xhat = ((xd - mean.view(shape)) * rstd.view(shape)).to(x.dtype)
        y = xhat if weight is None else xhat * weight.view(shape) + bias.view(shape)
        ctx.save_for_backward(xhat, weight, rstd.to(x.dtype), n)
        ctx.group, ctx.dims, ctx.shape = group, dims, shape
        ctx.mark_non_differentiable(mean, var, n)
        return y, mean, var, n

    @staticmethod
    def backward(ctx, dy, _dmean, _dvar, _dn):
        xhat, weight, rstd, n = ctx.saved_tensors
        C = xhat.shape[1]
        dyd = dy.double()
        s1 = dyd.sum(ctx.dims)
        s2 = (dyd * xhat.double()).sum(ctx.dims)
        pack = torch.cat([s1, s2])
        dist.all_reduce(pack, group=ctx.group)
        m1 = (pack[:C] / n).to(dy.dtype).view(ctx.shape)
        m2 = (pack[C:] / n).to(dy.dtype).view(ctx.shape)
        scale = rstd if weight is None else weight * rstd
        dx = scale.view(ctx.shape) * (dy - m1 - xhat * m2)
        if weight is None:
            return dx, None, None, None, None
        return dx, s2.to(weight.dtype), s1.to(weight.dtype), None, None  # LOCAL sums: the gradient exchange adds the ranks


class _SyncBNNativeFn(torch.autograd.Function):
    """The CUDA path: torch's own fused batch-norm kernels (the ones torch.nn.SyncBatchNorm launches), 4 launches + 1 collective each
    way, without its boolean-index filter of empty ranks (every rank of this trainer holds at least one utterance)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, group):
        x = x.contiguous()
        C = x.shape[1]
        mean_l, invstd_l = torch.batch_norm_stats(x, eps)
        count = torch.full((1,), x.numel() // C, dtype=mean_l.dtype, device=x.device)
        pack = torch.cat([mean_l, invstd_l, count])
        world = dist.get_world_size(group)
        allp = torch.empty(world, 2 * C + 1, dtype=pack.dtype, device=x.device)
        dist.all_gather_into_tensor(allp, pack, group=group)
        mean_all, invstd_all, count_all = allp[:, :C], allp[:, C:2 * C], allp[:, 2 * C]
        mean, invstd = torch.batch_norm_gather_stats_with_counts(x, mean_all, invstd_all, running_mean, running_var, momentum, eps, count_all)
        y = torch.batch_norm_elemt(x, weight, bias, mean, invstd, eps)
        ctx.save_for_backward(x, weight, mean, invstd, count_all.to(torch.int32))
        ctx.group = group
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, invstd, counts = ctx.saved_tensors
        dy = dy.contiguous()
        sum_dy, sum_dy_xmu, gw, gb = torch.batch_norm_backward_reduce(dy, x, mean, invstd, weight, True, weight is not None, weight is not None)
        C = sum_dy.numel()
        pack = torch.cat([sum_dy, sum_dy_xmu])
        dist.all_reduce(pack, group=ctx.group)
        dx = torch.batch_norm_backward_elemt(dy, x, mean, invstd, weight, pack[:C], pack[C:], counts)
        return dx, gw, gb, None, None, None, None, None  # gw / gb are LOCAL sums: the gradient exchange adds the ranks


class FastSyncBatchNorm(torch.nn.SyncBatchNorm):
    """torch.nn.SyncBatchNorm with the same parameters, buffers and statistics (biased batch variance for the output, unbiased for
    running_var, global element count), minus its host synchronisation: torch's forward filters empty ranks out of the gathered
    statistics with a boolean index -- an `aten::nonzero`, i.e. one device -> host round trip per layer.  The VP block holds 26 of
    them: measured 12 ms of a 33 ms data-parallel forward (tools/prof_train_dp.py).  CUDA tensors run torch's fused batch-norm
    kernels around one all-gather / all-reduce (`_SyncBNNativeFn`), CPU tensors (the gloo tests) a plain-tensor restatement in fp64
    (`_SyncBNFn`); nothing is read on the host either way."""

    force_sync = False  # tests: take the synchronised path in a one-rank process group too

    def forward(self, x):
        on = dist.is_available() and dist.is_initialized() and (self.force_sync or dist.get_world_size(self.process_group) > 1)
        if not (self.training and on):
            return super().forward(x)
        if x.is_cuda:
            mom = 0.0
            if self.track_running_stats and self.running_mean is not None:
                self.num_batches_tracked += 1
                mom = self.momentum if self.momentum is not None else 1.0 / float(self.num_batches_tracked)
            return _SyncBNNativeFn.apply(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps, mom, self.process_group)
        y, mean, var, n = _SyncBNFn.apply(x, self.weight, self.bias, self.eps, self.process_group)
        if self.track_running_stats and self.running_mean is not None:
            with torch.no_grad():
                self.num_batches_tracked += 1
                mom = self.momentum if self.momentum is not None else 1.0 / self.num_batches_tracked.double()
                self.running_mean.mul_(1 - mom).add_((mom * mean).to(self.running_mean.dtype))
                self.running_var.mul_(1 - mom).add_((mom * var * (n / torch.clamp(n - 1.0, min=1.0))).to(self.running_var.dtype))
        return y


def use_fast_sync_batchnorm(model):
    """Re-classes every torch.nn.SyncBatchNorm of `model` in place (parameters, buffers and state_dict keys untouched)."""
    k = 0
    for m in model.modules():
        if type(m) is torch.nn.SyncBatchNorm:
            m.__class__ = FastSyncBatchNorm
            k += 1
    return k


class Trainer:
    """Native training step: forward + SNR loss + backward + gradient all-reduce + clip + AdamW.

    Parameters live in ONE flat buffer (the nn.Parameters become views of it), gradients in another, so the data-parallel
    exchange is a single NCCL all-reduce of 2.96 MB (the reference's Lightning DDP does the same with buckets,
    train.py:135-146) and the optimizer is one fused kernel launch."""

    def __init__(self, model, lr=1e-3, weight_decay=0.1, betas=(0.9, 0.999), eps=1e-8, clip=5.0):
        self.model = model
        self.lr, self.wd, self.betas, self.eps, self.clip = lr, weight_decay, betas, eps, clip
        use_fast_sync_batchnorm(model)  # (no-op unless the model was converted with SyncBatchNorm.convert_sync_batchnorm)
        self.flat = FlatParams(model)
        self.flat_p, self.flat_g, self.params, self.world = self.flat.flat_p, self.flat.flat_g, self.flat.params, self.flat.world
        dev = self.flat_p.device
        self.m = torch.zeros_like(self.flat_p)
        self.v = torch.zeros_like(self.flat_p)
        self.gnorm = torch.zeros(1, device=dev, dtype=torch.float64)
        self.t = 0

    def exchange(self):
        self.flat.exchange()

    def step(self, wav, target, mouth, events=None):
        """One optimisation step on this rank's batch shard; returns the (local) loss as a 0-d tensor.
        events: optional list that receives 4 recorded CUDA events (start, after forward+loss, after backward, end)."""
        model = self.model
        model.train()
        mark = (lambda: events.append(_event())) if events is not None else (lambda: None)
        mark()
        self.flat_g.zero_()
        est = model(wav, mouth)
        loss, d_est = snr_loss(est.detach(), target, True)
        mark()
        est.backward(d_est.view_as(est))
        mark()
        self.exchange()
        self.t += 1
        with torch.cuda.device(self.flat_p.device):
            _lib.check(_lib.lib().rtfs_adamw_step(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.flat_p.numel(),
                                                  self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.t, self.clip, 1.0 / self.world,
                                                  self.gnorm.data_ptr(), _stream()), "rtfs_adamw_step")
        mark()
        rt = getattr(model, "_runtime", None)
        if rt is not None:  # next step's weight images, derived while the GPU still works through this step's queue
            prepared_slots(rt, self.flat_p.device, stash=True)
        return loss


# ------------------------------------------------------------------------------------------------- module-level (tests)
class ModuleHarness:
    """Module-level training entries (rtfs_block_/dprnn_/mhsa_ train_forward + backward) on one pass region; returns the
    gradients per slot.  Used by the gradient-parity tests to localise a failing stage."""

    def __init__(self, model, B, T, device):
        self.model, self.B, self.T, self.dev = model, B, T, device
        L = (T - 1) * 128
        self.plan = _lib.train_plan(B, L, 1, 1)
        self.ws = torch.zeros(self.plan["pass_bytes"], dtype=torch.uint8, device=device)
        self.scratch = torch.zeros(self.plan["bwd_bytes"], dtype=torch.uint8, device=device)
        with torch.no_grad():
            self.slots = prepare({k: v.detach() for k, v in live_tensors(model).items()}, device, train=True)
        # eval-statistics CAF slots are not used by these entries
        self.table = _table([self.slots[n] for n in _lib.PARAM_NAMES])

    def ws_view(self, name, shape):
        n = 1
        for s in shape:
            n *= s
        o = self.plan["pass_offsets"][name]
        return self.ws[o:o + 4 * n].view(torch.float32).view(*shape)

    def scratch_view(self, name, shape):
        n = 1
        for s in shape:
            n *= s
        o = self.plan["bwd_offsets"][name]
        return self.scratch[o:o + 4 * n].view(torch.float32).view(*shape)

    def _run(self, fwd, bwd, x, d_out):
        grads = GradBuffers(self.slots, self.dev)
        out = torch.empty_like(x)
        d_in = torch.empty_like(x)
        with torch.cuda.device(self.dev):
            fwd(out)
            bwd(grads, out, d_in)
        torch.cuda.synchronize()
        return out, d_in, grads.slot_order()

    def block(self, x, d_out):
        """x, d_out: (B,T,F,256) channels-last storage."""
        lib, B, T = _lib.lib(), self.B, self.T
        return self._run(
            lambda out: _lib.check(lib.rtfs_block_train_forward(self.table, x.data_ptr(), None, out.data_ptr(), self.ws.data_ptr(), B, T, _stream()), "block fwd"),
            lambda g, out, d_in: _lib.check(lib.rtfs_block_backward(self.table, g.table, x.data_ptr(), d_out.data_ptr(), d_in.data_ptr(), self.ws.data_ptr(),
                                                                     self.scratch.data_ptr(), B, T, _stream()), "block bwd"),
            x, d_out)

    def dprnn(self, which, g_in, d_out):
        lib, B, T = _lib.lib(), self.B, self.T
        return self._run(
            lambda out: _lib.check(lib.rtfs_dprnn_train_forward(self.table, which, g_in.data_ptr(), out.data_ptr(), self.ws.data_ptr(), B, T, _stream()), "dprnn fwd"),
            lambda g, out, d_in: _lib.check(lib.rtfs_dprnn_backward(self.table, g.table, which, g_in.data_ptr(), d_out.data_ptr(), d_in.data_ptr(),
                                                                     self.ws.data_ptr(), self.scratch.data_ptr(), B, T, _stream()), "dprnn bwd"),
            g_in, d_out)

    def mhsa(self, g_in, d_out):
        lib, B, T = _lib.lib(), self.B, self.T
        return self._run(
            lambda out: _lib.check(lib.rtfs_mhsa_train_forward(self.table, g_in.data_ptr(), out.data_ptr(), self.ws.data_ptr(), B, T, _stream()), "mhsa fwd"),
            lambda g, out, d_in: _lib.check(lib.rtfs_mhsa_backward(self.table, g.table, g_in.data_ptr(), d_out.data_ptr(), d_in.data_ptr(), self.ws.data_ptr(),
                                                                    self.scratch.data_ptr(), B, T, _stream()), "mhsa bwd"),
            g_in, d_out)


def slot_grads_to_param_grads(model, slot_grads, device):
    """Push per-slot gradients through the (differentiable) slot preparation: {parameter name: gradient}."""
    live = {k: (v.detach().clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in live_tensors(model).items()}
    slots = prepare(live, device, train=True)
    outs, gs = [], []
    for n, g in slot_grads.items():
        if slots[n] is not None and slots[n].requires_grad:
            outs.append(slots[n])
            gs.append(g.to(device))
    torch.autograd.backward(outs, gs)
    return {k: v.grad for k, v in live.items() if v.dtype.is_floating_point and v.grad is not None}
