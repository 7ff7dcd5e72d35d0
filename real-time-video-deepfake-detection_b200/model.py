"""Drop-in for the reference's ``model.DeepfakeEfficientNet`` (model.py:21-102).

A ``torch.nn.Module`` whose parameter/buffer tree reproduces the reference's ``state_dict`` layout
(lukemelas ``efficientnet_pytorch`` names under ``net.``) so checkpoints load with
zero missing keys, and whose ``forward`` runs in libdfd's CUDA kernels.  The modules are parameter
containers: no torch conv/matmul executes on the path.
"""
import warnings

import torch
import torch.nn as nn

from . import arch, runtime


class _Block(nn.Module):
    def __init__(self, b):
        super().__init__()
        if b.cexp != b.cin:
            self._expand_conv = nn.Conv2d(b.cin, b.cexp, 1, bias=False)
            self._bn0 = nn.BatchNorm2d(b.cexp, eps=arch.BN_EPS, momentum=0.01)
        self._depthwise_conv = nn.Conv2d(b.cexp, b.cexp, b.k, stride=b.s, groups=b.cexp, bias=False)
        self._bn1 = nn.BatchNorm2d(b.cexp, eps=arch.BN_EPS, momentum=0.01)
        self._se_reduce = nn.Conv2d(b.cexp, b.se, 1)
        self._se_expand = nn.Conv2d(b.se, b.cexp, 1)
        self._project_conv = nn.Conv2d(b.cexp, b.cout, 1, bias=False)
        self._bn2 = nn.BatchNorm2d(b.cout, eps=arch.BN_EPS, momentum=0.01)


class _EfficientNetB0(nn.Module):
    """Parameter container with lukemelas EfficientNet-B0's attribute names."""

    def __init__(self):
        super().__init__()
        self._conv_stem = nn.Conv2d(3, 32, 3, stride=2, bias=False)
        self._bn0 = nn.BatchNorm2d(32, eps=arch.BN_EPS, momentum=0.01)
        self._blocks = nn.ModuleList([_Block(b) for b in arch.BLOCKS])
        self._conv_head = nn.Conv2d(320, 1280, 1, bias=False)
        self._bn1 = nn.BatchNorm2d(1280, eps=arch.BN_EPS, momentum=0.01)
        self._avg_pooling = nn.AdaptiveAvgPool2d(1)
        self._dropout = nn.Dropout(0.2)
        self._fc = nn.Linear(1280, 1000)


class DeepfakeEfficientNet(nn.Module):
    """EfficientNet-B0 + custom classifier 1280 -> 512 -> 256 -> 1 (model.py:36-61)."""

    def __init__(self, pretrained=True, dropout=0.5, *, device=None, dtype="fp32"):
        super().__init__()
        if pretrained:
            warnings.warn("ImageNet weights cannot be downloaded here; the backbone is randomly initialised "
                          "(load a checkpoint with load_state_dict)", stacklevel=2)
        self.net = _EfficientNetB0()
        in_features = self.net._fc.in_features
        self.net._fc = nn.Sequential(
            nn.Dropout(dropout), nn.Linear(in_features, 512), nn.BatchNorm1d(512), nn.ReLU(),
            nn.Dropout(dropout * 0.7), nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(),
            nn.Dropout(dropout * 0.5), nn.Linear(256, 1))
        self._dfd_device = device
        self._dfd_dtype = dtype
        self._dfd_engine = None
        self._dfd_version = None
        self.eval()

    # -- engine plumbing ------------------------------------------------------------
    def _engine(self):
        from .engine import Engine
        if self._dfd_engine is None:
            dev = runtime.default_device() if self._dfd_device is None else self._dfd_device
            self._dfd_engine = Engine(device=dev, max_streams=1, max_batch=64, max_crop=64)
        ver = tuple(int(p._version) for p in self.parameters()) + tuple(int(b._version) for b in self.buffers())
        if ver != self._dfd_version:
            self._dfd_engine.load_state_dict({k: v.detach().float().cpu() for k, v in self.state_dict().items()})
            self._dfd_version = ver
        return self._dfd_engine

    def _run(self, rgb_input):
        if self.training:
            raise RuntimeError("the B200 path implements inference (eval mode) only")
        eng = self._engine()
        x = rgb_input.to(eng.device)
        if x.dim() != 4 or x.shape[1:] != (3, 224, 224):
            raise ValueError("expected (B, 3, 224, 224) normalized RGB input")
        tdt = torch.bfloat16 if self._dfd_dtype == "bf16" else torch.float32
        x = x.permute(0, 2, 3, 1).contiguous().to(tdt)           # NCHW -> NHWC (layout plumbing only)
        out = []
        for i in range(0, x.shape[0], 64):
            out.append(eng.effnet_forward(x[i:i + 64]))
        return eng, torch.cat(out)

    def forward(self, rgb_input, freq_input=None):
        """(B,3,224,224) -> (B,1) logits; ``freq_input`` is ignored as in the reference (model.py:63-72)."""
        _, logits = self._run(rgb_input)
        return logits.view(-1, 1).to(rgb_input.device)

    def extract_features(self, rgb_input):
        """(B,1280) pooled backbone features (model.py:74-88)."""
        eng = self._engine()
        eng.set_tap("features")
        try:
            self._run(rgb_input[:64])
            f = eng.activation("features").view(-1, 1280).clone()
        finally:
            eng.set_tap("")
        return f.to(rgb_input.device)

    def forward_with_projection(self, rgb_input, freq_input=None):
        return self.forward(rgb_input), None

    def get_feature_extractor(self):
        return self.net._conv_head


def compute_frequency_features(image_bgr_or_rgb, size=224):
    raise NotImplementedError("compute_frequency_features is dead input in the reference (the model ignores it, "
                              "model.py:67); it is out of scope for the B200 hot path")
