"""GPU parity: EfficientNet-B0 forward vs the CPU restatement (oracle/effnet.py)."""
import numpy as np
import pytest
import torch

import dfd_b200  # noqa: F401
from dfd_b200 import synth
from oracle import effnet as oeff

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from dfd_b200.engine import Engine
    sd = synth.make_state_dict()
    e = Engine(device=0, max_streams=8, max_batch=32, max_crop=64)
    missing, unexpected = e.load_state_dict(sd)
    assert not missing and not unexpected
    g = torch.Generator().manual_seed(99)
    x = synth._calib_batch(g, 24).float()            # NCHW, normalised
    taps = {}
    ref = oeff.forward(x, sd, taps)
    yield e, sd, x, ref, taps
    e.close()


def test_fp32_logits_and_probability(setup):
    e, sd, x, ref, taps = setup
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    logits = e.effnet_forward(xn).cpu()
    dz = (logits - ref.flatten()).abs()
    dp = (torch.sigmoid(logits) - torch.sigmoid(ref.flatten())).abs()
    print("fp32 max |dlogit|", float(dz.max()), "max |dprob|", float(dp.max()), "logit std", float(ref.std()))
    assert float(dp.max()) <= 1e-4                    # north_star fp32 gate
    assert float(ref.std()) > 0.5                     # the weights are not degenerate


@pytest.mark.parametrize("name", ["stem", "b0.out", "b1.expand", "b1.dw", "b1.out", "b3.out", "b5.dw", "b8.out", "b15.out", "features"])
def test_fp32_activation_taps(setup, name):
    e, sd, x, ref, taps = setup
    xn = x[:4].permute(0, 2, 3, 1).contiguous().cuda()
    e.set_tap(name)
    e.effnet_forward(xn)
    got = e.activation(name).cpu()
    e.set_tap("")
    t = taps[name][:4]
    want = t.permute(0, 2, 3, 1).reshape(-1) if t.dim() == 4 else t.reshape(-1)
    err = float((got - want).abs().max())
    scale = float(want.abs().max())
    print(name, "max abs err", err, "scale", scale)
    assert err <= 2e-4 * max(scale, 1.0)


def test_bf16_logits(setup):
    e, sd, x, ref, taps = setup
    xn = x.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    logits = e.effnet_forward(xn).cpu()
    dz = (logits - ref.flatten()).abs()
    dp = (torch.sigmoid(logits) - torch.sigmoid(ref.flatten())).abs()
    print("bf16 max |dlogit|", float(dz.max()), "mean", float(dz.mean()), "max |dprob|", float(dp.max()), "mean", float(dp.mean()))
    assert float(dz.mean()) < 0.5                     # sanity; the bf16 gate is reported in DESIGN.md


def test_batch_invariance_fp32(setup):
    e, sd, x, ref, taps = setup
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    a = e.effnet_forward(xn).cpu()
    b = torch.cat([e.effnet_forward(xn[i:i + 1]).cpu() for i in range(4)])
    assert float((a[:4] - b).abs().max()) < 1e-4


@pytest.mark.parametrize("name", ["stem", "b0.dw", "b0.out", "b1.expand", "b1.dw", "b2.dw", "b3.dw", "b4.dw", "b5.dw", "b6.dw",
                                  "b8.dw", "b9.out", "b11.dw", "b12.dw", "b15.dw", "b15.out", "features"])
def test_bf16_activation_taps(setup, name):
    """bf16 mode (tcgen05 GEMMs, tiled depthwise): every layer stays within bf16 rounding noise of the fp32
    oracle -- an indexing bug (tile edge, padding, channel tail) would show as an O(1) error."""
    e, sd, x, ref, taps = setup
    xn = x[:3].permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    e.set_tap(name)
    e.effnet_forward(xn)
    got = e.activation(name).cpu()
    e.set_tap("")
    t = taps[name][:3]
    want = t.permute(0, 2, 3, 1).reshape(-1) if t.dim() == 4 else t.reshape(-1)
    err = (got - want).abs()
    scale = float(want.abs().max())
    print(name, "bf16 max abs err", float(err.max()), "mean", float(err.mean()), "scale", scale)
    assert float(err.max()) <= 0.06 * max(scale, 1.0)
    assert float(err.mean()) <= 0.01 * max(scale, 1.0)
