"""The EfficientNet-B0 restatement (oracle/effnet.py) against an INDEPENDENT implementation: HF transformers'
EfficientNetModel (a port of the official TF EfficientNet) with the same weights mapped across.

`efficientnet_pytorch` (the reference's dependency, model.py:18) cannot be installed here, and the reference's
tests pin structure only, so this is what pins padding / BN eps / SE / skip / channel plan numerically."""
import pytest
import torch

import dfd_b200  # noqa: F401
from dfd_b200 import arch, synth
from oracle import effnet as oeff

transformers = pytest.importorskip("transformers")


def map_to_hf(sd):
    out = {}

    def bn(dst, src):
        for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            out[f"{dst}.{k}"] = sd[f"{src}.{k}"]

    out["embeddings.convolution.weight"] = sd["net._conv_stem.weight"]
    bn("embeddings.batchnorm", "net._bn0")
    for i, b in enumerate(arch.BLOCKS):
        p, q = f"net._blocks.{i}.", f"encoder.blocks.{i}."
        if b.cexp != b.cin:
            out[q + "expansion.expand_conv.weight"] = sd[p + "_expand_conv.weight"]
            bn(q + "expansion.expand_bn", p + "_bn0")
        out[q + "depthwise_conv.depthwise_conv.weight"] = sd[p + "_depthwise_conv.weight"]
        bn(q + "depthwise_conv.depthwise_norm", p + "_bn1")
        out[q + "squeeze_excite.reduce.weight"] = sd[p + "_se_reduce.weight"]
        out[q + "squeeze_excite.reduce.bias"] = sd[p + "_se_reduce.bias"]
        out[q + "squeeze_excite.expand.weight"] = sd[p + "_se_expand.weight"]
        out[q + "squeeze_excite.expand.bias"] = sd[p + "_se_expand.bias"]
        out[q + "projection.project_conv.weight"] = sd[p + "_project_conv.weight"]
        bn(q + "projection.project_bn", p + "_bn2")
    out["encoder.top_conv.weight"] = sd["net._conv_head.weight"]
    bn("encoder.top_bn", "net._bn1")
    return out


def test_restatement_matches_independent_port():
    from transformers import EfficientNetConfig, EfficientNetModel
    sd = synth.make_state_dict()
    cfg = EfficientNetConfig(width_coefficient=1.0, depth_coefficient=1.0, image_size=224, hidden_dim=1280)
    hf = EfficientNetModel(cfg).eval()
    missing, unexpected = hf.load_state_dict(map_to_hf(sd), strict=True)
    assert not missing and not unexpected
    g = torch.Generator().manual_seed(3)
    x = synth._calib_batch(g, 4).float()
    with torch.no_grad():
        ref = hf(pixel_values=x).pooler_output          # (B, 1280) pooled features
        got = oeff.features(x, sd)
    rel = float((ref - got).abs().max() / ref.abs().max())
    assert rel < 1e-4, rel


def test_classifier_head_matches_torch_modules():
    """The custom _fc (model.py:50-61) restated functionally == the nn.Sequential in eval mode."""
    import torch.nn as nn
    sd = synth.make_state_dict()
    fc = nn.Sequential(nn.Dropout(0.5), nn.Linear(1280, 512), nn.BatchNorm1d(512), nn.ReLU(), nn.Dropout(0.35),
                       nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Dropout(0.25), nn.Linear(256, 1)).eval()
    fc.load_state_dict({k[len("net._fc."):]: v for k, v in sd.items() if k.startswith("net._fc.")})
    f = torch.randn(5, 1280, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        assert float((fc(f) - oeff.classifier(f, sd)).abs().max()) < 1e-5
