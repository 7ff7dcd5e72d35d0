"""CPU: model.py mirror keeps the reference checkpoint layout (model.py:36-61; tests/test_functional.py:62-110,
tests/test_performance.py:234-241 of the reference) and the packer folds BatchNorm correctly."""
import numpy as np
import torch

import dfd_b200  # noqa: F401
from dfd_b200 import arch, synth, weights
from dfd_b200.model import DeepfakeEfficientNet
from oracle import effnet as oeff


def test_architecture_and_parameter_count():
    m = DeepfakeEfficientNet(pretrained=False)
    fc = m.net._fc
    assert len(fc) == 10 and fc[1].in_features == 1280 and fc[1].out_features == 512 and fc[9].out_features == 1
    n = sum(p.numel() for p in m.parameters())
    assert n == 4_796_541 and n < 8_000_000


def test_state_dict_layout_matches_reference_checkpoint():
    m = DeepfakeEfficientNet(pretrained=False)
    want = dict(arch.state_dict_spec())
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == {k: tuple(s) for k, s in want.items()}
    sd = synth.make_state_dict()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    # both container forms the reference accepts (deepfake_detection.py:44-51)
    assert weights.extract_state_dict({"model_state_dict": sd, "epoch": 3}) is sd
    assert weights.extract_state_dict(sd) is sd
    assert weights.check_keys(sd) == ([], [])
    assert weights.check_keys({k: v for k, v in sd.items() if "_fc.9" not in k})[0] == ["net._fc.9.weight", "net._fc.9.bias"]


def test_bn_folding_in_packed_blob():
    """Folded stem / expand / fc parameters reproduce conv+BN of the oracle on random inputs."""
    sd = synth.make_state_dict()
    blob = weights.pack_state_dict(sd)
    off = {n: (o, c) for n, o, c in weights.blob_layout()[0]}
    g = torch.Generator().manual_seed(0)
    # stem: y = conv(x) folded == bn(conv(x))
    x = torch.randn(2, 3, 32, 32, generator=g)
    w = torch.from_numpy(blob[off["stem.w"][0]:off["stem.w"][0] + 27 * 32]).view(3, 3, 3, 32).permute(3, 2, 0, 1)
    b = torch.from_numpy(blob[off["stem.b"][0]:off["stem.b"][0] + 32])
    ref = oeff._bn(torch.nn.functional.conv2d(x, sd["net._conv_stem.weight"]), sd, "net._bn0", oeff.BN_EPS)
    got = torch.nn.functional.conv2d(x, w, b)
    assert float((ref - got).abs().max()) < 1e-4
    # fc1 + BatchNorm1d
    f = torch.randn(4, 1280, generator=g)
    w1 = torch.from_numpy(blob[off["fc1.w"][0]:off["fc1.w"][0] + 512 * 1280]).view(512, 1280)
    b1 = torch.from_numpy(blob[off["fc1.b"][0]:off["fc1.b"][0] + 512])
    ref = torch.nn.functional.batch_norm(torch.nn.functional.linear(f, sd["net._fc.1.weight"], sd["net._fc.1.bias"]),
                                         sd["net._fc.2.running_mean"], sd["net._fc.2.running_var"], sd["net._fc.2.weight"],
                                         sd["net._fc.2.bias"], False, 0.0, oeff.FC_BN_EPS)
    assert float((ref - torch.nn.functional.linear(f, w1, b1)).abs().max()) < 1e-4
    assert blob.dtype == np.float32 and blob.size == weights.blob_layout()[1]
