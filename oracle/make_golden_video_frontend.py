"""TEST INFRASTRUCTURE ONLY -- golden vector of the reference's video front-end.

Run in the authoring container (needs /root/reference):    python -m oracle.make_golden_video_frontend

Imports the reference's own `src.models.videomodels.FRCNNVideoModel` (thop shimmed), loads the synthetic state_dict both sides
regenerate from the key names (11.2 M parameters: too large to commit; oracle.make_golden_ctcnet.synthetic_state_dict), runs it in
eval mode on seeded mouth ROIs and writes tests/golden/video_frontend_small.npz.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")


def inputs(B=2, T=3):
    g = torch.Generator().manual_seed(23)
    return torch.rand(B, 1, T, 88, 88, generator=g).half().float()  # stored as fp16 in the fixture: exactly representable


def main():
    from oracle.make_golden import import_reference
    from oracle.make_golden_ctcnet import synthetic_state_dict

    import_reference()
    from src.models.videomodels import FRCNNVideoModel, get

    assert get("frcnnvideomodel") is FRCNNVideoModel
    model = FRCNNVideoModel(backbone_type="resnet", relu_type="prelu", print_macs=False)
    model.eval()  # (the reference's train() override returns None)
    model.load_state_dict(synthetic_state_dict(model.state_dict()), strict=True)
    x = inputs()
    with torch.no_grad():
        front = model.frontend3D(x)
        out = model(x)
    keys = [[k, list(v.shape)] for k, v in model.state_dict().items()]
    np.savez_compressed(os.path.join(GOLD, "video_frontend_small.npz"), x=x.numpy().astype(np.float16), out_ref_fp32=out.numpy(),
                        front_stats=np.array([float(front.mean()), float(front.std())]),
                        keys=np.frombuffer(json.dumps(keys).encode(), dtype=np.uint8), n_params=sum(p.numel() for p in model.parameters()))
    print("wrote video_frontend_small.npz: out", tuple(out.shape), "rms", float(out.pow(2).mean().sqrt()))


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
