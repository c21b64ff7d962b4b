"""TEST INFRASTRUCTURE ONLY -- pins the oracle against the reference and writes tests/golden/.

Run in the authoring container (needs /root/reference; it does not exist on the GPU box):

    python -m oracle.make_golden            # from the repo root

What it does
  1. puts oracle/shims (sru / timm / thop / pytorch_lightning stand-ins) and /root/reference on
     sys.path and imports the reference's own `src.models.AVNet`;
  2. builds AVNet(print_macs=False, **yaml['audionet']) with torch.manual_seed(0), then
     randomises the affine / BatchNorm-statistics parameters that are constants at init
     (gamma=1, beta=0, running stats 0/1) so that the golden vectors exercise every parameter;
  3. runs the REFERENCE forward on seeded inputs in fp32 (and fp64 to bound oracle noise),
     recording boundary tensors with forward hooks;
  4. runs oracle/rtfs_oracle.py on the same state_dict and asserts agreement tensor by tensor;
  5. writes small fixtures (state_dict + inputs + strided samples of the boundary tensors +
     full output waveforms) to tests/golden/.
"""
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("RTFS_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.path.insert(0, os.path.join(HERE, "shims"))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    from src.models import AVNet  # the reference's own code

    return AVNet


def load_conf(name):
    with open(os.path.join(REF, "config", name)) as f:
        return yaml.safe_load(f)


def randomise_constants(model, seed=123):
    """Affine norm parameters are 1/0 and BN running stats 0/1 at init: perturb them (seeded) so a
    kernel that ignored one of them would fail parity."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(("norm.weight", "norm.gamma", ".gamma")) or (".full_layer.3.weight" in name) or name.endswith(("norm1.weight", "norm2.weight")):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif name.endswith(("norm.bias", "norm.beta", ".beta")) or (".full_layer.3.bias" in name) or name.endswith(("norm1.bias", "norm2.bias")):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif "rnn_lst" in name and name.endswith(".bias"):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(0.1 * torch.randn(b.shape, generator=g))
            elif name.endswith("running_var"):
                b.copy_(1.0 + 0.3 * torch.rand(b.shape, generator=g))


def make_inputs(B, L, Tv, seed=1):
    g = torch.Generator().manual_seed(seed)
    wav = 0.1 * torch.randn(B, L, generator=g)
    lip = torch.rand(B, 512, Tv, generator=g)
    return wav, lip


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def run_reference(model, wav, lip):
    """Reference forward with hooks on the §3.4 boundaries."""
    taps = {}
    hooks = []

    def save(name):
        def fn(mod, inp, out):
            taps.setdefault(name, []).append(out.detach().clone() if torch.is_tensor(out) else out[0].detach().clone())

        return fn

    rm = model.refinement_module
    blk = rm.audio_net.blocks
    hooks.append(model.encoder.register_forward_hook(save("a0")))
    hooks.append(model.audio_bottleneck.register_forward_hook(save("a1")))
    hooks.append(blk.register_forward_hook(save("blk")))
    hooks.append(blk.globalatt[0].register_forward_hook(save("g1")))
    hooks.append(blk.globalatt[1].register_forward_hook(save("g2")))
    hooks.append(blk.globalatt[2].register_forward_hook(save("g3")))
    hooks.append(blk.downsample_layers[0].register_forward_hook(save("d0")))
    hooks.append(blk.downsample_layers[1].register_forward_hook(save("d1")))
    hooks.append(blk.fusion_layers[0].register_forward_hook(save("f0")))
    hooks.append(blk.fusion_layers[1].register_forward_hook(save("f1")))
    hooks.append(rm.video_net.blocks.register_forward_hook(save("video")))
    hooks.append(rm.crossmodal_fusion.fusion_module.register_forward_hook(save("caf")))
    hooks.append(model.mask_generator.register_forward_hook(save("masked")))
    with torch.no_grad():
        out = model(wav, lip)
    for h in hooks:
        h.remove()
    return out, taps


def strided(t, n=4096):
    """A deterministic strided sample of a tensor (keeps fixtures small)."""
    flat = t.reshape(-1)
    step = max(1, flat.numel() // n)
    return flat[::step][:n].contiguous()


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    AVNet = import_reference()
    from oracle import rtfs_oracle as O

    os.makedirs(GOLD, exist_ok=True)
    report = []
    cases = [
        # (tag, yaml, B, seconds)
        ("rtfs4_b2_2s", "lrs2_RTFSNet_4_layer.yaml", 2, 2.0),
        ("rtfs4_b1_1s", "lrs2_RTFSNet_4_layer.yaml", 1, 1.0),
        ("rtfs12_b1_2s", "lrs2_RTFSNet_12_layer.yaml", 1, 2.0),
    ]
    sd_saved = False
    for tag, cfg, B, seconds in cases:
        conf = load_conf(cfg)
        R = conf["audionet"]["audio_params"]["repeats"]
        torch.manual_seed(0)
        model = AVNet(print_macs=False, **conf["audionet"]).eval()
        randomise_constants(model)
        L = int(16000 * seconds)
        Tv = int(25 * seconds)
        wav, lip = make_inputs(B, L, Tv)
        ref_out, taps = run_reference(model, wav, lip)

        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        otaps = {}
        with torch.no_grad():
            ora_out = O.avnet_forward(sd, wav, lip, R, otaps)
        # fp64 reference to bound fp32 oracle noise
        model64 = AVNet(print_macs=False, **conf["audionet"]).double().eval()
        model64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()})
        with torch.no_grad():
            ref64 = model64(wav.double(), lip.double())
            sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
            ora64 = O.avnet_forward(sd64, wav.double(), lip.double(), R)

        checks = {
            "out": rel_l2(ora_out, ref_out),
            "out_fp64_oracle_vs_ref": rel_l2(ora64, ref64),
            "ref_fp32_vs_fp64": rel_l2(ref_out, ref64),
            "a0": rel_l2(otaps["a0"], taps["a0"][0]),
            "a1": rel_l2(otaps["a1"], taps["a1"][0]),
            "blk0": rel_l2(otaps["blk0"], taps["blk"][0]),
            "video": rel_l2(otaps["video"], taps["video"][0]),
            "caf": rel_l2(otaps["caf"], taps["caf"][0]),
            "refined": rel_l2(otaps["refined"], taps["blk"][-1]),
            "masked": rel_l2(otaps["masked"], taps["masked"][0]),
        }
        sisdr_delta = float((O.neg_sisdr(ora_out, ref64.float()) - O.neg_sisdr(ref_out, ref64.float())).abs().max())
        checks["sisdr_delta_db"] = sisdr_delta
        print(tag, {k: f"{v:.3e}" for k, v in checks.items()})
        assert checks["out_fp64_oracle_vs_ref"] < 1e-10, checks
        for k in ("out", "a0", "a1", "blk0", "video", "caf", "refined", "masked"):
            assert checks[k] < 2e-5, (k, checks)
        report.append((tag, checks))

        if not sd_saved:
            np.savez_compressed(os.path.join(GOLD, "state_dict_rtfs.npz"), **{k: v.numpy() for k, v in sd.items()})
            sd_saved = True
        fx = {
            "wav": wav.numpy(),
            "lip": lip.numpy().astype(np.float16).astype(np.float32) if False else lip.numpy(),
            "out_ref_fp32": ref_out.numpy(),
            "out_ref_fp64": ref64.numpy(),
            "repeats": np.int64(R),
        }
        for name in ("a0", "a1", "d0", "d1", "g1", "g2", "g3", "f0", "f1", "video", "caf", "masked"):
            fx["tap_" + name] = strided(taps[name][0]).numpy()
        fx["tap_blk0"] = strided(taps["blk"][0]).numpy()
        fx["tap_refined"] = strided(taps["blk"][-1]).numpy()
        np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), **fx)

    # per-module golden case at a tiny geometry (kept whole): the RTFS block on a (1,256,19,129) input
    conf = load_conf("lrs2_RTFSNet_4_layer.yaml")
    torch.manual_seed(0)
    model = AVNet(print_macs=False, **conf["audionet"]).eval()
    randomise_constants(model)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 256, 19, 129, generator=g)
    with torch.no_grad():
        y = model.refinement_module.audio_net.blocks(x)
    np.savez_compressed(os.path.join(GOLD, "block_small.npz"), x=x.numpy(), y=y.numpy())

    with open(os.path.join(GOLD, "PINNING.txt"), "w") as f:
        f.write("oracle/rtfs_oracle.py vs the reference's src/models (a4cd7361) executed with oracle/shims, rel-L2:\n")
        for tag, checks in report:
            f.write(tag + " " + " ".join(f"{k}={v:.3e}" for k, v in checks.items()) + "\n")
        f.write("SRU recurrence: PARITY UNPINNED (third-party `sru` absent; both sides use oracle/sru_ref.py).\n")
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
