"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI of
librtfs_b200.so; the checker is the CPU oracle (oracle/rtfs_oracle.py) and the committed golden
vectors produced by the reference itself (tests/golden/, oracle/make_golden.py).

Tolerances (north_star): separated waveform <= 1e-3 relative L2 and |delta SI-SDR| <= 0.01 dB against
the reference's fp32 forward.  The contractions run in TF32 (as the reference does on GPU under
torch.set_float32_matmul_precision("high")), so single stages are held to 3e-3 and pure-fp32
stages (STFT, iSTFT, depthwise convs, norms) to 2e-5.
"""
import os

import pytest
import torch

from conftest import ROOT, build_model, load_case, rel_l2, strided

pytestmark = pytest.mark.gpu

TOL_TF32 = 3e-3
TOL_FP32 = 2e-5
BLK = "refinement_module.audio_net.blocks."
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.txt")


def report(line):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(line + "\n")
    print(line)


@pytest.fixture(scope="module")
def O():
    from oracle import rtfs_oracle

    return rtfs_oracle


@pytest.fixture(scope="module")
def model(golden_sd):
    assert torch.cuda.is_available()
    return build_model(golden_sd, 4)


def ws_view(model, name, shape):
    """A workspace buffer of the last call as a torch tensor (debug/diagnostic access)."""
    from rtfs_net_b200 import _lib

    rt = model._runtime
    B, T, Tv, _ = rt._ws_key
    _, offs = _lib.ws_plan(B, (T - 1) * 128, Tv)
    n = 1
    for s in shape:
        n *= s
    return rt._ws[offs[name]: offs[name] + 4 * n].view(torch.float32).view(*shape)


def test_library_loaded():
    from rtfs_net_b200 import _lib

    assert _lib.lib().rtfs_abi_version() == _lib.ABI_VERSION


def test_encoder(model, golden_sd, O):
    case = load_case("rtfs4_b2_2s")
    with torch.no_grad():
        a0 = model.encoder(case["wav"].cuda())
        ref = O.encoder(golden_sd, case["wav"])
    e = rel_l2(a0, ref)
    report(f"encoder a0 rel_l2={e:.3e}")
    assert a0.shape == ref.shape
    assert e < TOL_FP32
    assert rel_l2(strided(a0.cpu()), case["tap_a0"]) < TOL_FP32


def test_bottleneck(model, golden_sd, O):
    case = load_case("rtfs4_b2_2s")
    with torch.no_grad():
        a0 = O.encoder(golden_sd, case["wav"])
        ref = O.audio_bottleneck(golden_sd, a0)
        a1 = model.audio_bottleneck(a0.cuda())
    e = rel_l2(a1, ref)
    report(f"bottleneck a1 rel_l2={e:.3e}")
    assert e < TOL_TF32


@pytest.mark.parametrize("which,dim", [(0, 4), (1, 3)])
@pytest.mark.parametrize("Tc,B", [(125, 2), (63, 2), (50, 3), (126, 1), (250, 1)])  # odd tile counts, half-empty last tiles, one
# utterance; 250 frames: one sequence per 256-position tile on the time path
def test_dprnn(model, golden_sd, O, which, dim, Tc, B):
    g = torch.Generator().manual_seed(3 + which)
    x = torch.randn(B, 64, Tc, 64, generator=g)
    with torch.no_grad():
        ref = O.dual_path_rnn(golden_sd, BLK + f"globalatt.{which}.", x, dim)
        out = model.refinement_module.audio_net.blocks.globalatt[which](x.cuda())
    e = rel_l2(out, ref)
    e_delta = rel_l2(out.cpu() - x, ref - x)
    report(f"dprnn which={which} Tc={Tc} rel_l2={e:.3e} (update only: {e_delta:.3e})")
    assert e_delta < TOL_TF32


@pytest.mark.parametrize("Tc", [125, 63, 250])
def test_mhsa(model, golden_sd, O, Tc):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, Tc, 64, generator=g)
    with torch.no_grad():
        ref = O.mhsa2d(golden_sd, BLK + "globalatt.2.", x)
        out = model.refinement_module.audio_net.blocks.globalatt[2](x.cuda())
    e_delta = rel_l2(out.cpu() - x, ref - x)
    report(f"mhsa Tc={Tc} update rel_l2={e_delta:.3e}")
    assert e_delta < TOL_TF32


STAGES = [
    # (workspace buffer, oracle tap, full resolution?, tolerance)
    ("RTFS_WS_P_PRE", "p_pre", True, TOL_TF32),
    ("RTFS_WS_D0_PRE", "d0_pre", True, TOL_TF32),
    ("RTFS_WS_D1_PRE", "d1_pre", False, TOL_TF32),
    ("RTFS_WS_POOL", "pool", False, TOL_TF32),
    ("RTFS_WS_G0", "g0", False, TOL_TF32),
    ("RTFS_WS_G1", "g1", False, TOL_TF32),
    ("RTFS_WS_G2", "g2", False, TOL_TF32),
    ("RTFS_WS_G3", "g3", False, TOL_TF32),
    ("RTFS_WS_GE0", "f0_e_pre", False, TOL_TF32),
    ("RTFS_WS_GG0", "f0_g_pre", False, TOL_TF32),
    ("RTFS_WS_LE0_PRE", "f0_l_pre", True, TOL_TF32),
    ("RTFS_WS_GE1", "f1_e_pre", False, TOL_TF32),
    ("RTFS_WS_GG1", "f1_g_pre", False, TOL_TF32),
    ("RTFS_WS_LE1", "f1_l_pre", False, TOL_TF32),
    ("RTFS_WS_GEC", "c0_e_pre", False, TOL_TF32),
    ("RTFS_WS_GGC", "c0_g_pre", False, TOL_TF32),
    ("RTFS_WS_LEC_PRE", "c0_l_pre", True, TOL_TF32),
]


def _block_stages(model, golden_sd, O, x, tag):
    B, _, T, Fq = x.shape
    Tc = (T - 2) // 2 + 1
    taps = {}
    with torch.no_grad():
        ref = O.rtfs_block(golden_sd, BLK, x, taps)
        out = model.refinement_module.audio_net.blocks(x.cuda())
    torch.cuda.synchronize()
    worst = []
    for buf, tap, full, tol in STAGES:
        shape = (B, T, Fq, 64) if full else (B, Tc, 64, 64)
        ours = ws_view(model, buf, shape).permute(0, 3, 1, 2)
        e = rel_l2(ours, taps[tap])
        report(f"block[{tag}] {tap:<10} rel_l2={e:.3e}")
        if not e < tol:
            worst.append((tap, e))
    e = rel_l2(out, ref)
    report(f"block[{tag}] out        rel_l2={e:.3e}")
    assert not worst, worst
    assert e < TOL_TF32
    return out


def test_block_small_golden(model, golden_sd, O):
    """The reference's own TDANetBlock output on a (1,256,19,129) input (tests/golden/block_small.npz)."""
    case = load_case("block_small")
    out = _block_stages(model, golden_sd, O, case["x"], "small")
    e = rel_l2(out, case["y"])
    report(f"block[small] vs reference golden rel_l2={e:.3e}")
    assert e < TOL_TF32


@pytest.mark.parametrize("T", [251, 126])
def test_block_full(model, golden_sd, O, T):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 256, T, 129, generator=g)
    _block_stages(model, golden_sd, O, x, f"T{T}")


def test_caf(model, golden_sd, O):
    g = torch.Generator().manual_seed(13)
    a = torch.randn(2, 256, 251, 129, generator=g)
    v = torch.rand(2, 512, 50, generator=g)
    p = "refinement_module.crossmodal_fusion.fusion_module.audio_lstm."
    with torch.no_grad():
        ref = O.caf(golden_sd, p, a, v)
        out, v_out = model.refinement_module.crossmodal_fusion.get_fusion_block(0)(a.cuda(), v.cuda())
    e = rel_l2(out, ref)
    report(f"caf rel_l2={e:.3e}")
    assert e < TOL_FP32 * 5
    assert torch.equal(v_out.cpu(), v)


def test_video_block(model, golden_sd, O):
    g = torch.Generator().manual_seed(17)
    v = torch.rand(2, 512, 50, generator=g)
    with torch.no_grad():
        ref = O.video_block(golden_sd, "refinement_module.video_net.blocks.", v)
        out = model.refinement_module.video_net.get_block(0)(v.cuda())
    e = rel_l2(out, ref)
    report(f"video block rel_l2={e:.3e}")
    assert e < TOL_TF32


def test_mask_decoder(model, golden_sd, O):
    g = torch.Generator().manual_seed(19)
    refined = torch.randn(2, 256, 251, 129, generator=g)
    a0 = torch.randn(2, 256, 251, 129, generator=g)
    with torch.no_grad():
        zref = O.s3_mask(golden_sd, refined, a0)
        z = model.mask_generator(refined.cuda(), a0.cuda())
        e = rel_l2(z, zref)
        report(f"mask rel_l2={e:.3e}")
        assert z.shape == zref.shape
        assert e < TOL_TF32
        wref = O.decoder(golden_sd, zref, 32000)
        w = model.decoder(zref.cuda(), torch.Size([2, 32000]))
        e = rel_l2(w, wref)
        report(f"decoder rel_l2={e:.3e}")
        assert w.shape == wref.shape
        assert e < TOL_FP32 * 5


def _sisdr_db(O, est, target):
    return -O.neg_sisdr(est.cpu().float(), target.cpu().float())


@pytest.mark.parametrize("tag", ["rtfs4_b2_2s", "rtfs4_b1_1s", "rtfs12_b1_2s"])
def test_full_forward_golden(golden_sd, O, tag):
    """End-to-end parity against the reference's own fp32 forward (tests/golden/<tag>.npz)."""
    case = load_case(tag)
    R = case["repeats"]
    m = build_model(golden_sd, R)
    with torch.no_grad():
        out = m(case["wav"].cuda(), case["lip"].cuda())
    ref = case["out_ref_fp32"]
    assert out.shape == ref.shape
    e = rel_l2(out, ref)
    target = case["wav"][:, None, :]
    d = float((_sisdr_db(O, out, target) - _sisdr_db(O, ref, target)).abs().max())
    report(f"forward[{tag}] waveform rel_l2={e:.3e} |dSI-SDR|={d:.2e} dB")
    assert e <= 1e-3
    assert d <= 0.01


def test_forward_stage_taps(golden_sd, O):
    """Boundary tensors of the full forward against the reference's strided golden samples."""
    case = load_case("rtfs4_b2_2s")
    m = build_model(golden_sd, 4)
    with torch.no_grad():
        out = m(case["wav"].cuda(), case["lip"].cuda())
    B, T = 2, 251
    a0 = ws_view(m, "RTFS_WS_A0", (B, T, 129, 256)).permute(0, 3, 1, 2)
    e = rel_l2(strided(a0.cpu()), case["tap_a0"])
    report(f"forward taps a0 rel_l2={e:.3e}")
    assert e < TOL_FP32
    best = 1.0
    for buf in ("RTFS_WS_XA", "RTFS_WS_XB"):
        x = ws_view(m, buf, (B, T, 129, 256)).permute(0, 3, 1, 2)
        best = min(best, rel_l2(strided(x.cpu()), case["tap_refined"]))
    report(f"forward taps refined rel_l2={best:.3e}")
    assert best < TOL_TF32
    # the fused S^3 mask + decoder epilogue never writes the masked embedding: re-form it with the module-level entry from the
    # refined features and a0 the forward left in the workspace
    refined = min((ws_view(m, buf, (B, T, 129, 256)) for buf in ("RTFS_WS_XA", "RTFS_WS_XB")),
                  key=lambda x: rel_l2(strided(x.permute(0, 3, 1, 2).cpu()), case["tap_refined"]))
    with torch.no_grad():
        z = m.mask_generator(refined.permute(0, 3, 1, 2).clone(), a0.clone())[:, 0]
    e = rel_l2(strided(z.cpu()), case["tap_masked"])
    report(f"forward taps masked rel_l2={e:.3e}")
    assert e < TOL_TF32


def test_stft_istft_round_trip(golden_sd):
    """Size-independent property at BASELINE batch size: with an encoder that copies (Re, Im) into
    channels 0/1 and a decoder that reads them back, decoder(encoder(wav)) == wav."""
    sd = {k: v.clone() for k, v in golden_sd.items()}
    sd["encoder.conv.full_layer.2.weight"].zero_()
    sd["encoder.conv.full_layer.2.weight"][0, 0, 1, 1] = 1.0
    sd["encoder.conv.full_layer.2.weight"][1, 1, 1, 1] = 1.0
    sd["decoder.decoder.weight"].zero_()
    sd["decoder.decoder.weight"][0, 0, 1, 1] = 1.0
    sd["decoder.decoder.weight"][1, 1, 1, 1] = 1.0
    m = build_model(sd, 4)
    g = torch.Generator().manual_seed(23)
    wav = (0.1 * torch.randn(32, 32000, generator=g)).cuda()
    with torch.no_grad():
        a0 = m.encoder(wav)
        back = m.decoder(a0, wav.shape)
    e = rel_l2(back[:, 0], wav)
    report(f"stft->istft round trip B=32 rel_l2={e:.3e}")
    assert e < 1e-5


def test_batch_independence_full_size(golden_sd):
    """BASELINE config 2 size (B=32, 2 s): utterances are independent, so a batch of 32 must
    reproduce the 2-utterance golden case wherever those utterances sit in the batch."""
    case = load_case("rtfs4_b2_2s")
    m = build_model(golden_sd, 4)
    g = torch.Generator().manual_seed(29)
    wav = 0.1 * torch.randn(32, 32000, generator=g)
    lip = torch.rand(32, 512, 50, generator=g)
    wav[5], wav[31] = case["wav"][0], case["wav"][1]
    lip[5], lip[31] = case["lip"][0], case["lip"][1]
    with torch.no_grad():
        out = m(wav.cuda(), lip.cuda())
    assert out.shape == (32, 1, 32000)
    assert torch.isfinite(out).all()
    e0 = rel_l2(out[5], case["out_ref_fp32"][0])
    e1 = rel_l2(out[31], case["out_ref_fp32"][1])
    report(f"B=32 batch independence rel_l2={e0:.3e},{e1:.3e}")
    assert e0 <= 1e-3 and e1 <= 1e-3


def test_errors(model):
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            model.encoder(torch.zeros(1, 32000))  # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        model.encoder(torch.zeros(1, 32000, device="cuda"))  # module-level entries are inference-only: autograd is enabled here
    with pytest.raises(ValueError):
        with torch.no_grad():
            model(torch.zeros(1, 1000, device="cuda"), torch.zeros(1, 512, 50, device="cuda"))  # too short: validated up front


def test_forward_4s_vs_oracle(golden_sd, O):
    """BASELINE config 4 geometry (4 s -> T = 501, T' = 250: one 250-step sequence per fused-RNN tile, 250 attention
    tokens) against the CPU oracle computed on the fly."""
    g = torch.Generator().manual_seed(31)
    wav = 0.1 * torch.randn(1, 64000, generator=g)
    lip = torch.rand(1, 512, 100, generator=g)
    m = build_model(golden_sd, 4)
    with torch.no_grad():
        out = m(wav.cuda(), lip.cuda())
        ref = O.avnet_forward(golden_sd, wav, lip, 4)
    e = rel_l2(out, ref)
    d = float((_sisdr_db(O, out, wav[:, None, :]) - _sisdr_db(O, ref, wav[:, None, :])).abs().max())
    report(f"forward[4s b1 R4] waveform rel_l2={e:.3e} |dSI-SDR|={d:.2e} dB")
    assert out.shape == (1, 1, 64000)
    assert e <= 1e-3 and d <= 0.01


@pytest.mark.parametrize("L", [16000 + 128 * 3, 24000])
def test_forward_ragged_lengths(golden_sd, O, L):
    """Lengths that give odd / even frame counts and a partially filled last fused-RNN tile (B = 3)."""
    g = torch.Generator().manual_seed(37)
    wav = 0.1 * torch.randn(3, L, generator=g)
    lip = torch.rand(3, 512, max(8, L // 640), generator=g)
    m = build_model(golden_sd, 4)
    with torch.no_grad():
        out = m(wav.cuda(), lip.cuda())
        ref = O.avnet_forward(golden_sd, wav, lip, 4)
    e = rel_l2(out, ref)
    report(f"forward[L={L} b3] waveform rel_l2={e:.3e}")
    assert out.shape == (3, 1, L)
    assert e <= 1e-3


def test_kernel_generations_agree(golden_sd):
    """The first-generation kernels (mma.sync GEMMs, strip depthwise, unfused RNN, separate CAF pass) stay in the library
    behind environment switches as the measured baseline; both generations must give the same waveform."""
    import subprocess
    import sys

    code = (
        "import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from conftest import build_model, load_case\n"
        "import numpy as np\n"
        "g = np.load(%r); sd = {k: torch.from_numpy(g[k]) for k in g.files}\n"
        "c = load_case('rtfs4_b2_2s'); m = build_model(sd, 4)\n"
        "with torch.no_grad(): out = m(c['wav'].cuda(), c['lip'].cuda())\n"
        "torch.save(out.cpu(), sys.argv[1])\n"
    ) % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden", "state_dict_rtfs.npz"))
    outs = {}
    variants = (
        ("new", {}),
        ("old", {"RTFS_LEGACY_GEMM": "1", "RTFS_LEGACY_DW": "1", "RTFS_UNFUSED_CAF": "1", "RTFS_NO_VIDEO_GRAPH": "1", "RTFS_LEGACY_FRONTEND": "1"}),
        # second generation: tcgen05 GEMMs one tile per CTA, unfused RNN around them, scalar TF-AR conv, mma.sync attention convs
        ("mid", {"RTFS_PERSIST_MASK": "0", "RTFS_UNFUSED_DPRNN": "1", "RTFS_SCALAR_TFAR": "1", "RTFS_LEGACY_ATT": "1", "RTFS_LEGACY_FRONTEND": "1"}),
        # fused RNN tilings: one 256-position tile per CTA everywhere (round-1 shape) / 128-position tiles as independent CTAs;
        # the default runs two 128-position pipelines per persistent CTA (dprnn_fused.cuh)
        ("df256", {"RTFS_DF_TILE": "256"}),
        ("df128", {"RTFS_DF_TILE": "128"}),
    )
    for tag, env in variants:
        path = os.path.join(ROOT, "gpurun_out", f"gen_{tag}.pt")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        subprocess.run([sys.executable, "-c", code, path], check=True, env={**os.environ, **env}, timeout=600)
        outs[tag] = torch.load(path)
    for tag in ("old", "mid", "df256", "df128"):
        e = rel_l2(outs["new"], outs[tag])
        report(f"kernel generations new vs {tag} rel_l2={e:.3e}")
        assert e <= 1e-3


# ------------------------------------------------------------------------------- round 2: sizes and variety
def test_forward_r6_vs_oracle(golden_sd, O):
    """RTFS-Net-6 (BASELINE configs[2] geometry) against the CPU oracle."""
    g = torch.Generator().manual_seed(41)
    wav = 0.1 * torch.randn(2, 16000, generator=g)
    lip = torch.rand(2, 512, 25, generator=g)
    m = build_model(golden_sd, 6)
    with torch.no_grad():
        out = m(wav.cuda(), lip.cuda())
        ref = O.avnet_forward(golden_sd, wav, lip, 6)
    e = rel_l2(out, ref)
    d = float((_sisdr_db(O, out, wav[:, None, :]) - _sisdr_db(O, ref, wav[:, None, :])).abs().max())
    report(f"forward[1s b2 R6] waveform rel_l2={e:.3e} |dSI-SDR|={d:.2e} dB")
    assert e <= 1e-3 and d <= 0.01


def test_b32_eight_utterances_vs_oracle(golden_sd, O):
    """BASELINE config 2 at size (B=32, 2 s, R=4): eight utterances spread over the batch against the CPU oracle, the
    other 24 against the same utterances run as a second batch in a different order (batch independence)."""
    g = torch.Generator().manual_seed(43)
    wav = 0.1 * torch.randn(32, 32000, generator=g)
    lip = torch.rand(32, 512, 50, generator=g)
    m = build_model(golden_sd, 4)
    picks = [0, 3, 9, 14, 17, 22, 27, 31]
    with torch.no_grad():
        out = m(wav.cuda(), lip.cuda()).cpu()
        ref = O.avnet_forward(golden_sd, wav[picks], lip[picks], 4)
        perm = torch.randperm(32, generator=g)
        out_p = m(wav[perm].cuda(), lip[perm].cuda()).cpu()
    errs = [rel_l2(out[i], ref[j]) for j, i in enumerate(picks)]
    report("B=32 eight utterances vs oracle rel_l2 max=%.3e" % max(errs))
    assert max(errs) <= 1e-3
    e_perm = max(rel_l2(out_p[k], out[int(perm[k])]) for k in range(32))
    report(f"B=32 permuted batch vs original rel_l2 max={e_perm:.3e}")
    # Not bit-equal, by construction: an utterance's rows fall into different 128-row tiles at a different batch position,
    # so the fp32 partial sums behind the fp64 gLN statistics group differently (1e-7 relative); every later TF32 operand
    # rounding turns such a perturbation into occasional one-ulp (2^-10) flips, which saturate at the TF32 noise level after
    # a few layers (measured 2.4e-4, the same size as the distance to the fp32 oracle).  Bound: the north-star 1e-3.
    assert e_perm <= 1e-3


def test_cfg4_at_size(golden_sd, O):
    """BASELINE config 4 at size (B=64, 4 s, R=12; 29 GB workspace): two utterances of the batch against the CPU oracle."""
    g = torch.Generator().manual_seed(47)
    wav = 0.1 * torch.randn(64, 64000, generator=g)
    lip = torch.rand(64, 512, 100, generator=g)
    m = build_model(golden_sd, 12)
    with torch.no_grad():
        out = m(wav.cuda(), lip.cuda()).cpu()
        ref = O.avnet_forward(golden_sd, wav[[7, 63]], lip[[7, 63]], 12)
    assert out.shape == (64, 1, 64000) and torch.isfinite(out).all()
    e = max(rel_l2(out[7], ref[0]), rel_l2(out[63], ref[1]))
    d = float((_sisdr_db(O, out[[7, 63]], wav[[7, 63], None, :]) - _sisdr_db(O, ref, wav[[7, 63], None, :])).abs().max())
    report(f"forward[cfg4 B=64 4s R12] two utterances vs oracle rel_l2={e:.3e} |dSI-SDR|={d:.2e} dB")
    del m
    torch.cuda.empty_cache()
    assert e <= 1e-3 and d <= 0.01


def test_non_default_stream_and_cuda_graph(golden_sd):
    """The library launches on the caller's stream and never synchronises: the forward runs on a side stream and can be
    captured into a CUDA graph; both reproduce the default-stream result bit for bit up to
    the fp64 statistic atomics."""
    case = load_case("rtfs4_b2_2s")
    m = build_model(golden_sd, 4)
    wav, lip = case["wav"].cuda(), case["lip"].cuda()
    with torch.no_grad():
        base = m(wav, lip).clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            out_side = m(wav, lip).clone()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        e_side = rel_l2(out_side, base)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out_graph = m(wav, lip)
        out_graph.zero_()
        graph.replay()
        torch.cuda.synchronize()
        e_graph = rel_l2(out_graph, base)
    report(f"side stream vs default rel_l2={e_side:.3e}; CUDA-graph replay vs default rel_l2={e_graph:.3e}")
    assert e_side <= 1e-5 and e_graph <= 1e-5
    assert rel_l2(out_graph, case["out_ref_fp32"]) <= 1e-3


def test_trained_like_weight_scales(golden_sd, O):
    """TF32 margin beyond random-init weights: every weight tensor gets its own gain (log-uniform in [0.5, 2]), biases
    and norm offsets become non-zero, BatchNorm running statistics move away from (0, 1) -- the waveform must still be
    within the 1e-3 bar of the fp32 oracle."""
    g = torch.Generator().manual_seed(53)
    sd = {}
    for k, v in golden_sd.items():
        v = v.clone()
        if v.dtype.is_floating_point and "pos_enc" not in k and "scale_x" not in k:
            if k.endswith("running_var"):
                v = v * float(torch.empty(1).uniform_(0.5, 2.0, generator=g))
            elif k.endswith("running_mean") or k.endswith(".bias") or k.endswith(".beta") or k.endswith("norm.bias"):
                v = v + 0.05 * torch.randn(v.shape, generator=g)
            else:
                v = v * float(2.0 ** torch.empty(1).uniform_(-1.0, 1.0, generator=g))
        sd[k] = v
    wav = 0.1 * torch.randn(2, 16000, generator=g)
    lip = torch.rand(2, 512, 25, generator=g)
    m = build_model(sd, 4)
    with torch.no_grad():
        out = m(wav.cuda(), lip.cuda())
        ref = O.avnet_forward(sd, wav, lip, 4)
    e = rel_l2(out, ref)
    d = float((_sisdr_db(O, out, wav[:, None, :]) - _sisdr_db(O, ref, wav[:, None, :])).abs().max())
    report(f"forward[trained-like scales] waveform rel_l2={e:.3e} |dSI-SDR|={d:.2e} dB")
    assert e <= 1e-3 and d <= 0.01


@pytest.mark.parametrize("Tv", [50, 100, 25, 37, 8])
def test_video_block_kernel(model, golden_sd, O, Tv):
    """The VP block as one kernel (csrc/video.cuh, fp32 FMAs) against the CPU oracle, and against the torch-module path it replaces."""
    g = torch.Generator().manual_seed(59 + Tv)
    v = torch.rand(3, 512, Tv, generator=g)
    with torch.no_grad():
        ref = O.video_block(golden_sd, "refinement_module.video_net.blocks.", v)
        out = model._runtime.video_block(v.cuda())
        torch_path = model.refinement_module.video_net.get_block(0)(v.cuda())
    e = rel_l2(out, ref)
    report(f"video block kernel Tv={Tv} rel_l2={e:.3e} (torch modules on the GPU: {rel_l2(torch_path, ref):.3e})")
    assert out.shape == ref.shape
    assert e < TOL_FP32


def test_streaming_separator_matches_direct_calls(golden_sd):
    """shard.StreamingSeparator (double-buffered host -> device -> host loop, the bench's end-to-end leg): five different host
    batches come back equal to the direct model calls, whatever the slot reuse."""
    from rtfs_net_b200 import shard

    m = build_model(golden_sd, 4)
    g = torch.Generator().manual_seed(77)
    B, L, Tv = 3, 16000, 25
    batches = [((0.1 * torch.randn(B, L, generator=g)).pin_memory(), torch.rand(B, 512, Tv, generator=g).pin_memory(), torch.empty(B, 1, L).pin_memory())
               for _ in range(5)]
    sep = shard.StreamingSeparator(m, B, L, Tv)
    for w, l, o in batches:
        sep.submit(w, l, o)
    sep.drain()
    with torch.no_grad():
        for i, (w, l, o) in enumerate(batches):
            ref = m(w.cuda(), l.cuda()).cpu()
            e = rel_l2(o, ref)
            report(f"streaming separator batch {i} vs direct call rel_l2={e:.3e}")
            assert e <= 1e-5  # same kernels, same batch position: differences only from the order of the fp64 statistics atomics
