"""CPU-side tests (python -m pytest tests -m "not gpu"): the oracle against the reference's golden
vectors, the host logic (state_dict mirror, weight preparation, registry, error behaviour), and
the C-ABI library surface (loads, exports every declared symbol; no compute calls without a GPU).
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import GOLD, ROOT, audionet_conf, load_case, rel_l2, strided


@pytest.fixture(scope="module")
def O():
    from oracle import rtfs_oracle

    return rtfs_oracle


# ------------------------------------------------------------------------------- oracle vs golden
def test_oracle_block_matches_reference_golden(golden_sd, O):
    case = load_case("block_small")
    with torch.no_grad():
        y = O.rtfs_block(golden_sd, "refinement_module.audio_net.blocks.", case["x"])
    assert rel_l2(y, case["y"]) < 2e-5


def test_oracle_forward_matches_reference_golden(golden_sd, O):
    """The 1 s / B=1 fixture (T = 126 even: exercises the 2-wide pooling windows)."""
    case = load_case("rtfs4_b1_1s")
    taps = {}
    with torch.no_grad():
        out = O.avnet_forward(golden_sd, case["wav"], case["lip"], case["repeats"], taps)
    assert out.shape == case["out_ref_fp32"].shape
    assert rel_l2(out, case["out_ref_fp32"]) < 2e-5
    assert rel_l2(out, case["out_ref_fp64"]) < 2e-5
    for name in ("a0", "a1", "video", "caf", "masked"):
        assert rel_l2(strided(taps[name]), case["tap_" + name]) < 2e-5, name
    assert rel_l2(strided(taps["refined"]), case["tap_refined"]) < 2e-5


def test_oracle_sru_c_and_torch_scans_agree():
    from oracle import sru_ref

    g = torch.Generator().manual_seed(0)
    L, B, d = 13, 5, 32
    for n_in, k in ((512, 4), (64, 3)):
        x = torch.randn(L, B, n_in, generator=g)
        w = torch.randn(n_in, 2 * d * k, generator=g) / n_in ** 0.5
        wc = torch.randn(4 * d, generator=g)
        bias = torch.randn(4 * d, generator=g)
        U = (x.reshape(L * B, n_in) @ w).view(L, B, 2 * d, k)
        h_t, c_t = sru_ref.sru_scan_torch(U, x if k == 3 else None, wc, bias, d, 2, k)
        h, c = sru_ref.sru_layer_forward(x, w, wc, bias, d, True)
        assert rel_l2(h, h_t) < 1e-5 and rel_l2(c, c_t) < 1e-5


def test_oracle_stft_istft_round_trip(O):
    g = torch.Generator().manual_seed(1)
    for L in (16000, 32000, 16100):
        wav = torch.randn(2, L, generator=g)
        spec = O.stft_spec(wav)
        assert spec.shape == (2, 2, L // 128 + 1, 129)
        ref = torch.stft(wav, 256, 128, window=torch.hann_window(256), return_complex=True)
        assert rel_l2(spec[:, 0], ref.real.transpose(1, 2)) < 1e-5
        back = O.istft(spec[:, 0], spec[:, 1], L)
        assert rel_l2(back, wav) < 1e-5


def test_oracle_neg_sisdr_known_answer(O):
    t = torch.randn(3, 1, 4000, generator=torch.Generator().manual_seed(2))
    assert float(O.neg_sisdr(2.5 * t, t).max()) < -60.0  # scale invariance: perfect estimate
    noise = torch.randn(3, 1, 4000, generator=torch.Generator().manual_seed(3))
    noise = noise - (noise * t).sum(-1, keepdim=True) / (t * t).sum(-1, keepdim=True) * t
    est = t + noise * (t.norm(dim=-1, keepdim=True) / noise.norm(dim=-1, keepdim=True)) * 0.1
    assert torch.allclose(-O.neg_sisdr(est, t), torch.full((3,), 20.0), atol=0.3)


# ------------------------------------------------------------------------------- host logic
def test_state_dict_keys_and_shapes_match_reference(golden_sd):
    from rtfs_net_b200 import AVNet

    m = AVNet(print_macs=False, **audionet_conf(4))
    sd = m.state_dict()
    # the golden state_dict was written through the sru shim (no `scale_x`); upstream cells carry that buffer (App. B)
    extra = set(sd.keys()) - set(golden_sd.keys())
    assert set(golden_sd.keys()) <= set(sd.keys()) and all(k.endswith(".scale_x") for k in extra) and len(extra) == 8
    for k, v in golden_sd.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    assert sum(p.numel() for p in m.parameters()) == 740210
    m.load_state_dict(golden_sd, strict=True)
    assert "window" not in " ".join(sd.keys())  # STFT windows are non-persistent buffers


def test_variants_share_one_block():
    from rtfs_net_b200 import AVNet

    n = [sum(p.numel() for p in AVNet(print_macs=False, **audionet_conf(r)).parameters()) for r in (4, 6, 12)]
    assert n == [740210] * 3


def test_analytic_macs_match_published_table():
    from rtfs_net_b200 import AVNet

    for r, published in ((4, 21.9), (6, 30.5), (12, 56.4)):
        macs = AVNet(print_macs=False, **audionet_conf(r)).get_MACs()
        assert abs(sum(macs.values()) / 1e9 - published) < 0.15


def test_registry():
    import rtfs_net_b200 as R

    assert R.get("avnet") is R.AVNet and R.get("AVNet") is R.AVNet
    with pytest.raises(ValueError):
        R.get("nope")
    with pytest.raises(ValueError):
        R.register_model(R.AVNet)


def test_unsupported_configs_fail_loudly():
    """STFT-domain (CUDA path) configurations outside the RTFS-Net form raise instead of falling back to eager modules."""
    from rtfs_net_b200 import AVNet

    conf = audionet_conf(4)
    conf["enc_dec_params"]["decoder_type"] = "ConvolutionalDecoder"  # STFT encoder with a 1-D decoder
    with pytest.raises((NotImplementedError, TypeError)):
        AVNet(print_macs=False, **conf)
    conf = audionet_conf(4)
    conf["audio_params"]["layers"]["layer_1"]["rnn_type"] = "LSTM"
    with pytest.raises(NotImplementedError):
        AVNet(print_macs=False, **conf)
    conf = audionet_conf(4)
    conf["audio_params"]["shared"] = False  # non-shared 2-D blocks
    with pytest.raises(NotImplementedError):
        AVNet(print_macs=False, **conf)
    conf = audionet_conf(4)
    conf["fusion_params"]["fusion_type"] = "ConcatFusion"  # 2-D audio block with a non-CAF fusion
    with pytest.raises(NotImplementedError):
        AVNet(print_macs=False, **conf)
    conf = audionet_conf(4)
    conf["audio_params"]["layers"]["layer_3"]["layer_type"] = "CBAMBlock"  # exists in the reference, out of scope here
    with pytest.raises(NotImplementedError):
        AVNet(print_macs=False, **conf)
    conf = audionet_conf(4)
    conf["audio_params"]["audio_net"] = "NoSuchNet"
    with pytest.raises(ValueError):
        AVNet(print_macs=False, **conf)


def test_no_cpu_fallback(golden_sd):
    from rtfs_net_b200 import AVNet

    m = AVNet(print_macs=False, **audionet_conf(4)).eval()
    with torch.no_grad():
        with pytest.raises(RuntimeError):
            m(torch.zeros(1, 32000), torch.zeros(1, 512, 50))
    with pytest.raises(RuntimeError):  # autograd / train(): the training path is CUDA-only as well
        m(torch.zeros(1, 32000), torch.zeros(1, 512, 50))
    with pytest.raises(NotImplementedError):  # module-level calls are inference entries
        m.encoder(torch.zeros(1, 32000))


def test_tf32_round():
    from rtfs_net_b200.weights import tf32_round

    x = torch.randn(10000, generator=torch.Generator().manual_seed(4)) * 3
    r = tf32_round(x)
    assert (r.view(torch.int32) & 0x1FFF).abs().sum() == 0
    assert float(((r - x).abs() / x.abs()).max()) <= 2.0 ** -11 * 1.0001
    assert torch.equal(tf32_round(r), r)
    assert torch.equal(tf32_round(-x), -r)


def test_weight_preparation_layouts(golden_sd):
    from rtfs_net_b200 import _lib
    from rtfs_net_b200.weights import prepare, tf32_round

    pp = prepare(golden_sd, "cpu")
    assert list(pp.keys()) == _lib.PARAM_NAMES
    blk = "refinement_module.audio_net.blocks."
    # unfold(8) + Linear == GEMM over the overlapping row view: check the K/N permutation numerically
    g = torch.Generator().manual_seed(5)
    n = torch.randn(20, 64, generator=g)  # (S, C) one sequence, channels-last
    X = n.t().unfold(1, 8, 1).permute(1, 0, 2).reshape(13, 512)  # nn.Unfold order c*8+k
    W0 = golden_sd[blk + "globalatt.0.rnn.rnn_lst.0.weight"]
    U_ref = (X @ W0).view(13, 64, 4)  # (l, col, m)
    A = torch.stack([n.reshape(-1)[l * 64: l * 64 + 512] for l in range(13)])
    U = A @ pp["RTFS_P_RF_W0"].t()  # (l, m*64+col)
    assert rel_l2(U.view(13, 4, 64).permute(0, 2, 1), U_ref) < 2e-3
    # ConvTranspose1d == GEMM over the overlapping view of the zero-padded sequence
    Y = torch.randn(13, 64, generator=g)
    ct = torch.nn.functional.conv_transpose1d(Y.t()[None], golden_sd[blk + "globalatt.0.linear.weight"])[0].t()  # (20, 64)
    hpad = torch.zeros(20 + 7 + 8, 64)
    hpad[7:20] = Y
    A = torch.stack([hpad.reshape(-1)[s * 64: s * 64 + 512] for s in range(20)])
    assert rel_l2(A @ pp["RTFS_P_RF_CTW"].t(), ct) < 2e-3
    # mask rows interleaved real/imag
    Wm = golden_sd["mask_generator.mask_generator.1.full_layer.2.weight"].reshape(256, 256)
    assert torch.equal(pp["RTFS_P_MK_W"][0::2], tf32_round(Wm[:128])) and torch.equal(pp["RTFS_P_MK_W"][1::2], tf32_round(Wm[128:]))
    # folded eval BatchNorm of the CAF key embedding
    q = "refinement_module.crossmodal_fusion.fusion_module.audio_lstm.key_embed.full_layer."
    a = torch.randn(3, 256, 2, 2, generator=g)
    ref = torch.nn.functional.batch_norm(a * golden_sd[q + "2.weight"].view(1, -1, 1, 1), golden_sd[q + "3.running_mean"],
                                         golden_sd[q + "3.running_var"], golden_sd[q + "3.weight"], golden_sd[q + "3.bias"], False, 0.0, 1e-5)
    ours = a * pp["RTFS_P_CAF_SK"].view(1, -1, 1, 1) + pp["RTFS_P_CAF_TK"].view(1, -1, 1, 1)
    assert rel_l2(ours, ref) < 1e-5


# ------------------------------------------------------------------------------- C ABI surface
def test_library_exports_every_declared_symbol():
    from rtfs_net_b200 import _lib

    declared = set(_lib.declared_functions())
    assert declared, "no functions parsed from include/rtfs_b200.h"
    assert declared == set(_lib.PROTOTYPES.keys())
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -m rtfs_net_b200.build"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert _lib.lib().rtfs_abi_version() == _lib.ABI_VERSION


def test_workspace_plan_is_consistent():
    from rtfs_net_b200 import _lib

    total, offs = _lib.ws_plan(2, 32000, 50)
    order = [offs[n] for n in _lib.WS_NAMES]
    assert order == sorted(order) and order[0] == 0 and all(o % 256 == 0 for o in order)
    assert total >= order[-1] and total > offs["RTFS_WS_ATT"]
    T, Fq = 251, 129
    assert offs["RTFS_WS_A1"] - offs["RTFS_WS_A0"] >= 2 * T * Fq * 256 * 4
    # the buffers whose size depends on the video length come last, so module-level calls that
    # do not know Tv address the same offsets
    _, offs0 = _lib.ws_plan(2, 32000, 0)
    tape = lambda n: n.startswith("RTFS_WS_TF_") or n.startswith("RTFS_WS_TT_")  # training tape: empty in the inference plan
    assert all(offs0[n] == offs[n] for n in _lib.WS_NAMES if n not in ("RTFS_WS_ATT",) and not tape(n))
    assert len({offs[n] for n in _lib.WS_NAMES if tape(n)}) == 1
    big, _ = _lib.ws_plan(32, 32000, 50)
    assert big < 12e9


def test_golden_fixtures_present():
    for f in ("state_dict_rtfs.npz", "rtfs4_b2_2s.npz", "rtfs4_b1_1s.npz", "rtfs12_b1_2s.npz", "block_small.npz", "PINNING.txt"):
        assert os.path.exists(os.path.join(GOLD, f)), f


# ------------------------------------------------------------------------------- docs / bench bookkeeping
def test_documented_runtime_switches_exist_in_the_sources():
    """Every RTFS_* environment switch INTEGRATION.md section 5 documents is read somewhere in the sources (and vice versa)."""
    import re

    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    documented = set(re.findall(r"`(RTFS_[A-Z0-9_]+)=", doc))
    src = ""
    for base, _, files in os.walk(os.path.join(ROOT, "rtfs_net_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".py")):
                src += open(os.path.join(base, f)).read()
    read = set(re.findall(r'(?:env_flag|getenv|environ\.get)\(\s*"(RTFS_[A-Z0-9_]+)"', src))
    read -= {"RTFS_B200_LIB"}  # library location, documented in section 1
    assert documented - read == set(), f"documented but never read: {sorted(documented - read)}"
    assert read - documented == set(), f"read but not documented: {sorted(read - documented)}"


def test_bench_roofline_bookkeeping():
    """Algorithmic bytes / FLOPs used for the roofline: block and forward totals, fused-RNN FLOPs, committed traffic file."""
    import json
    import sys

    sys.path.insert(0, ROOT)
    import bench
    from rtfs_net_b200 import _lib

    per_stage, block, fwd = bench.stage_bytes(32)
    T, F = bench.L // 128 + 1, 129
    Tc, Fc = (T - 2) // 2 + 1, 64
    A, H, G = 4 * 256 * T * F * 32, 4 * 64 * T * F * 32, 4 * 64 * Tc * Fc * 32
    assert block == 4 * A + 14 * H + 36 * G
    assert fwd == (6 + 4 * bench.REPEATS) * A + 14 * bench.REPEATS * H + 36 * bench.REPEATS * G
    # residual conv per pass variant: passes 2..R-1 read the addend (3A), the last pass does not (2A); the first pass
    # (CAF fused, addend aliases x) moves 2A
    R = bench.REPEATS
    assert per_stage["RTFS_SG_RESID_OUT"] == ((R - 2) * 3 * A + 2 * A) / (R - 1) + 2 * H + 2 * G
    assert per_stage["RTFS_SG_RESID_OUT_CAF"] == 2 * A + 2 * H + 2 * G
    assert set(per_stage) <= set(_lib.STAGE_NAMES)
    # per unfolded step (L = S - 7) Linear 512->256, three SRU layers 64->192 (k = 3), ConvTranspose1d share
    # 64x512; average of the two paths; = (1430.1 + 1515.8) / 2 MMAC per utterance
    per_step = 512 * 256 + 3 * 64 * 192 + 64 * 512
    assert bench.dprnn_flops(32) == 32 * (Tc * (Fc - 7) + Fc * (Tc - 7)) * per_step * 2.0 / 2.0
    assert bench.dprnn_flops(1) / 2.0 / 1e6 == pytest.approx((1430.1 + 1515.8) / 2.0, rel=1e-3)
    traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["stages"]
    for name, v in traffic.items():
        assert name in _lib.STAGE_NAMES or name == "RTFS_SG_RESID_OUT_CAF", name
        assert v["dram_bytes"] == pytest.approx(v["dram_read_bytes"] + v["dram_write_bytes"])
    # the dominant HBM kernel moves what the algorithm says it must (within 2 %)
    # (the capture is of a middle pass, which reads the addend: 3A + 2H + 2G)
    assert traffic["RTFS_SG_RESID_OUT"]["dram_bytes"] == pytest.approx(3 * A + 2 * H + 2 * G, rel=0.02)


def test_abi_tables_frozen_in_package_match_header():
    """rtfs_net_b200/_abi.py (what the installed package uses) is in sync with include/rtfs_b200.h."""
    from rtfs_net_b200 import _abi, _lib

    params, ws, stats, stages, fns, ver, tape, bwd = _lib.header_tables()
    assert (_abi.ABI_VERSION, _abi.PARAM_NAMES, _abi.WS_NAMES, _abi.STAT_NAMES, _abi.STAGE_NAMES, _abi.FUNCTIONS, _abi.TAPE_NAMES, _abi.BWD_NAMES) == (
        ver, params, ws, stats, stages, fns, tape, bwd)


def test_sru_scale_x_optional_on_load(golden_sd):
    """Checkpoints of the real `sru` package carry a `scale_x` buffer per cell; shim-written ones do not.
    Both must load with strict=True."""
    from rtfs_net_b200 import AVNet

    m = AVNet(print_macs=False, **audionet_conf(4))
    m.load_state_dict(golden_sd, strict=True)  # without scale_x
    sd = dict(golden_sd)
    n = 0
    for k in list(golden_sd):
        if k.endswith("rnn_lst.0.weight_c") or ".rnn_lst." in k and k.endswith(".weight_c"):
            sd[k[: -len("weight_c")] + "scale_x"] = torch.zeros(1)
            n += 1
    assert n == 8
    m.load_state_dict(sd, strict=True)  # with scale_x
    assert sum(1 for k in m.state_dict() if k.endswith("scale_x")) == 8
