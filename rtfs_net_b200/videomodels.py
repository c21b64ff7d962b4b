"""Video front-end: the lip-reading backbone that turns mouth ROIs (B, 1, T, 88, 88) into the (B, 512, T) lip embedding the
RTFS-Net forward consumes (reference: src/models/videomodels/frcnn_videomodel.py:16-72, resnet.py:24-130,
called from src/system/core.py:87-92 under `torch.no_grad()` with frozen parameters).

API-compatible plain PyTorch modules on library (cuDNN) kernels -- this step is NOT accelerated with hand-written kernels: it is
the caller of the hot path, included so that a user of the reference finds the `videomodels` surface (class names, constructor
keywords, state_dict keys of `FRCNNVideoModel` with the ResNet-18 trunk, `get(name)`, `update_frcnn_parameter`).  Checked against
a reference-generated fixture (tests/test_video_frontend.py).  Two inference conveniences the reference does not have, both
off by default and numerically neutral at fp32: `channels_last()` storage for the 2-D trunk and `capture()` = replay of the
~70 library launches from one CUDA graph per input shape.
"""
import math

import torch
import torch.nn as nn


def _act(relu_type, channels):
    if relu_type == "prelu":
        return nn.PReLU(num_parameters=channels)
    if relu_type == "relu":
        return nn.ReLU(inplace=True)
    raise Exception("relu type not implemented")


class BasicBlock(nn.Module):
    """resnet.py:24-66: conv3x3-BN-act-conv3x3-BN (+ projected) residual, act."""

    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, relu_type="relu"):
        super().__init__()
        assert relu_type in ["relu", "prelu"]
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu1 = _act(relu_type, planes)
        self.relu2 = _act(relu_type, planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = self.bn2(self.conv2(self.relu1(self.bn1(self.conv1(x)))))
        return self.relu2(y + (x if self.downsample is None else self.downsample(x)))


class ResNet(nn.Module):
    """resnet.py:69-130 (the trunk: four stages + global average pool; no stem, no classifier)."""

    def __init__(self, block, layers, num_classes=1000, relu_type="relu", gamma_zero=False, avg_pool_downsample=False):
        super().__init__()
        self.inplanes = 64
        self.relu_type = relu_type
        self.gamma_zero = gamma_zero
        self.avg_pool_downsample = avg_pool_downsample
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            setattr(self, f"layer{i + 1}", self._make_layer(block, planes, n, stride=1 if i == 0 else 2))
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        for m in self.modules():  # resnet.py:93-101
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, math.sqrt(2.0 / (m.kernel_size[0] * m.kernel_size[1] * m.out_channels)))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        if gamma_zero:
            for m in self.modules():
                if isinstance(m, BasicBlock):
                    m.bn2.weight.data.zero_()

    def _project(self, inplanes, outplanes, stride):
        if self.avg_pool_downsample:  # resnet.py:16-21
            return nn.Sequential(nn.AvgPool2d(stride, stride, ceil_mode=True, count_include_pad=False),
                                 nn.Conv2d(inplanes, outplanes, 1, 1, bias=False), nn.BatchNorm2d(outplanes))
        return nn.Sequential(nn.Conv2d(inplanes, outplanes, 1, stride, bias=False), nn.BatchNorm2d(outplanes))

    def _make_layer(self, block, planes, blocks, stride=1):
        proj = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            proj = self._project(self.inplanes, planes * block.expansion, stride)
        stage = [block(self.inplanes, planes, stride, proj, relu_type=self.relu_type)]
        self.inplanes = planes * block.expansion
        stage += [block(self.inplanes, planes, relu_type=self.relu_type) for _ in range(1, blocks)]
        return nn.Sequential(*stage)

    def forward(self, x):
        for i in range(1, 5):
            x = getattr(self, f"layer{i}")(x)
        return self.avgpool(x).flatten(1)


class FRCNNVideoModel(nn.Module):
    """frcnn_videomodel.py:16-72: Conv3d(1 -> 64, 5x7x7, stride (1,2,2)) + BatchNorm3d + PReLU + MaxPool3d(1x3x3, stride (1,2,2)),
    then the 2-D trunk per frame; (B, 1, T, 88, 88) -> (B, 512, T)."""

    def __init__(self, backbone_type="resnet", relu_type="prelu", width_mult=1.0, pretrain=None, print_macs=True, *args, **kwargs):
        super().__init__()
        if backbone_type != "resnet":
            raise NotImplementedError("only the ResNet-18 trunk of the shipped configurations is provided (ShuffleNetV2 is out of scope)")
        self.backbone_type = backbone_type
        self.frontend_nout, self.backend_out = 64, 512
        self.trunk = ResNet(BasicBlock, [2, 2, 2, 2], relu_type=relu_type)
        self.frontend3D = nn.Sequential(
            nn.Conv3d(1, self.frontend_nout, (5, 7, 7), (1, 2, 2), (2, 3, 3), bias=False),
            nn.BatchNorm3d(self.frontend_nout),
            nn.PReLU(num_parameters=self.frontend_nout) if relu_type == "prelu" else nn.ReLU(),
            nn.MaxPool3d((1, 3, 3), (1, 2, 2), (0, 1, 1)),
        )
        self.pretrain = pretrain
        self._graphs = {}
        if pretrain:
            self.init_from(pretrain)
        if print_macs:
            self.get_MACs()

    def _forward(self, x):
        B = x.shape[0]
        y = self.frontend3D(x)                                   # (B, 64, T, 22, 22)
        T = y.shape[2]
        y = y.transpose(1, 2).reshape(B * T, y.shape[1], y.shape[3], y.shape[4])  # threeD_to_2D_tensor, frcnn_videomodel.py:10-13
        if getattr(self, "_channels_last", False):
            y = y.contiguous(memory_format=torch.channels_last)
        return self.trunk(y).view(B, T, -1).transpose(1, 2).contiguous()

    def forward(self, x):
        g = self._graphs.get((x.device, tuple(x.shape))) if self._graphs else None
        if g is not None and not torch.is_grad_enabled():
            graph, static_in, static_out = g
            static_in.copy_(x)
            graph.replay()
            return static_out.clone()
        return self._forward(x)

    # ---- inference conveniences (not in the reference)
    def channels_last(self):
        self.trunk.to(memory_format=torch.channels_last)
        self._channels_last = True
        return self

    def capture(self, example):
        """Record the forward for inputs shaped like `example` (CUDA, eval mode, no grad) into a CUDA graph."""
        assert example.is_cuda and not self.training
        static_in = example.clone()
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(2):
                self._forward(static_in)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(graph):
            static_out = self._forward(static_in)
        self._graphs[(example.device, tuple(example.shape))] = (graph, static_in, static_out)
        return self

    # ---- reference surface
    def init_from(self, path):
        update_frcnn_parameter(self, torch.load(path, map_location="cpu")["model_state_dict"])

    def train(self, mode=True):
        super().train(mode)
        if mode:  # frcnn_videomodel.py:77-83: BatchNorm statistics stay frozen
            for m in self.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    m.eval()
        return self

    def get_MACs(self):
        """frcnn_videomodel.py:85-100 with the closed-form count of the conv / linear MACs (thop is not a dependency)."""
        macs = 0

        def hook(m, inp, out):
            nonlocal macs
            if isinstance(m, (nn.Conv2d, nn.Conv3d)):
                macs += out.numel() * (m.in_channels // m.groups) * math.prod(m.kernel_size)

        hs = [m.register_forward_hook(hook) for m in self.modules() if isinstance(m, (nn.Conv2d, nn.Conv3d))]
        with torch.no_grad():
            dev = next(self.parameters()).device
            self._forward(torch.rand(1, 1, 50, 88, 88, device=dev))
        for h in hs:
            h.remove()
        self.macs = macs / 1e6
        self.number_of_parameters = sum(p.numel() for p in self.parameters()) / 1000
        print("Pretrained Video Backbone\nNumber of MACs: {:,.1f}M\nNumber of parameters: {:,.1f}K\n".format(self.macs, self.number_of_parameters))


def update_frcnn_parameter(model, pretrained_dict):
    """frcnn_videomodel.py:103-115: load everything but the TCN head of the lip-reading checkpoint, then freeze."""
    sd = model.state_dict()
    sd.update({k: v for k, v in pretrained_dict.items() if "tcn" not in k})
    model.load_state_dict(sd)
    for p in model.parameters():
        p.requires_grad = False
    return model


def register_model(custom_model):
    """videomodels/__init__.py:23-33."""
    if custom_model.__name__ in globals().keys() or custom_model.__name__.lower() in globals().keys():
        raise ValueError(f"Model {custom_model.__name__} already exists. Choose another name.")
    globals().update({custom_model.__name__: custom_model})


def get(identifier):
    """videomodels/__init__.py:36-50: class from a case-insensitive name."""
    if isinstance(identifier, str):
        cls = {k.lower(): v for k, v in globals().items()}.get(identifier.lower())
        if cls is not None:
            return cls
    raise ValueError(f"Could not interpret model name : {str(identifier)}")
