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
    """bf16 mode end to end, reported against (a) the fp32 oracle and (b) the oracle's bf16-STORAGE restatement.

    What is asserted -- and what is NOT.  north_star's bf16 gate is |dp| <= 5e-3 against the reference (fp32).  On the
    fixed-seed random-init weights that gate is NOT met and cannot be: the synthetic network amplifies ANY 2^-9-relative
    perturbation (rounding the input image alone) into |dp| ~ 0.03, so the bf16-storage restatement itself is ~0.05 away
    from fp32 and two bf16 implementations that differ in one rounding (tanh.approx vs exact swish) are ~0.01-0.03 apart.
    The kernels' own arithmetic is therefore gated where the amplification cannot interfere -- per block, on identical
    inputs, in test_bf16_blockwise_exactness -- and this test bounds the end-to-end numbers as a regression guard:
    max |dp| vs fp32 <= 0.08 and no worse than the storage restatement by more than 1.5x.  The parity-green fast mode is
    dtype="fp32" (3xTF32 tensor-core path, <= 1e-4), see test_fp32_*."""
    e, sd, x, ref, taps = setup
    xn = x.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    logits = e.effnet_forward(xn).cpu()
    emu = oeff.forward_bf16_storage(x, sd).flatten()
    ref = ref.flatten()
    p, pe, pr = torch.sigmoid(logits), torch.sigmoid(emu), torch.sigmoid(ref)
    print("bf16 CUDA vs bf16-storage oracle: max |dlogit|", float((logits - emu).abs().max()), "max |dp|", float((p - pe).abs().max()),
          "mean |dp|", float((p - pe).abs().mean()))
    print("bf16-storage oracle vs fp32 oracle (inherent): max |dp|", float((pe - pr).abs().max()), "mean", float((pe - pr).abs().mean()))
    print("bf16 CUDA vs fp32 oracle: max |dp|", float((p - pr).abs().max()), "mean", float((p - pr).abs().mean()))
    print("north_star bf16 gate (5e-3 vs fp32):", "MET" if float((p - pr).abs().max()) <= 5e-3 else "NOT MET on these weights")
    assert float((p - pr).abs().max()) <= 0.08
    assert float((p - pr).abs().mean()) <= 1.5 * float((pe - pr).abs().mean()) + 1e-3    # no worse than bf16 storage itself
    assert float((p - pe).abs().max()) <= 0.06                                           # and near the restatement


@pytest.mark.parametrize("blk", list(range(16)))
def test_bf16_blockwise_exactness(setup, blk):
    """The bf16 kernels' own error, with the network's amplification taken out: block `blk` of the CPU bf16-storage
    restatement is fed the CUDA path's OWN input to that block (the previous tap, exact bf16 values) and must reproduce
    the CUDA depthwise output and block output to within a few bf16 ulps: max <= 2^-6 and mean <= 2^-10 of the tensor
    scale (the old test allowed 6 % / 1 % against the end-to-end fp32 taps).  What may differ: fp32 accumulation order,
    tanh.approx vs exact swish (2^-11), the fixed-order squeeze partials."""
    e, sd, x, ref, taps = setup
    xn = x[:4].permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    got = {}
    names = ["stem" if blk == 0 else f"b{blk - 1}.out", f"b{blk}.dw", f"b{blk}.out"]
    shapes = {}
    k, st, cin, cexp, cout, se = oeff.BLOCKS[blk]
    hin = [112, 112, 56, 56, 28, 28, 14, 14, 14, 14, 14, 14, 7, 7, 7, 7][blk]
    hout = (hin + st - 1) // st
    shapes[names[0]] = (4, hin, hin, cin)
    shapes[names[1]] = (4, hout, hout, cexp)
    shapes[names[2]] = (4, hout, hout, cout)
    for name in names:
        e.set_tap(name)
        e.effnet_forward(xn)
        got[name] = e.activation(name).cpu().reshape(shapes[name]).permute(0, 3, 1, 2).contiguous()
    e.set_tap("")
    out, dw = oeff.block_bf16_storage(got[names[0]], sd, blk, gated_weight=blk < 5)
    for name, want in ((names[1], dw), (names[2], out)):
        d = (got[name] - want).abs()
        scale = max(1.0, float(want.abs().max()))
        print(name, "max", float(d.max()), "mean", float(d.mean()), "scale", scale, "exact frac", float((d == 0).float().mean()))
        assert float(d.max()) <= 2.0 ** -6 * scale, name
        assert float(d.mean()) <= 2.0 ** -10 * scale, name


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


@pytest.mark.parametrize("blk", list(range(1, 16)))
def test_fused_front_matches_two_kernel_path(setup, blk):
    """mbconv_fused.cu (expand 1x1 on tcgen05 -> depthwise, expanded tensor kept in shared memory) must reproduce the
    expand-GEMM + k_dw_tile pair bit for bit at the depthwise output: same bf16 rounding points, same accumulation
    order.  The block output may differ by bf16 rounding only (the SE squeeze partials are summed per tile, and the two
    paths tile the image differently)."""
    e, sd, x, ref, taps = setup
    xn = x[:5].permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    got = {}
    for mode in (1, 0):
        e.set_option("no_fuse", mode)
        for name in (f"b{blk}.dw", f"b{blk}.out"):
            e.set_tap(name)
            e.effnet_forward(xn)
            got[(mode, name)] = e.activation(name).cpu()
    e.set_option("no_fuse", 0)
    e.set_tap("")
    a, b = got[(1, f"b{blk}.dw")], got[(0, f"b{blk}.dw")]
    assert a.shape == b.shape
    if blk == 1:        # block 1's input (b0.out) is the same tensor in both modes: the outputs must be identical
        assert torch.equal(a, b), float((a - b).abs().max())
    # deeper blocks see inputs that already differ by bf16 rounding (squeeze partials are summed per tile and the two
    # paths tile the image differently), so compare within a few bf16 ulps of the tensor's scale
    assert float((a - b).abs().max()) <= 2.0 ** -5 * max(1.0, float(a.abs().max()))
    assert float((a - b).abs().mean()) <= 2.0 ** -9 * max(1.0, float(a.abs().max()))
    a, b = got[(1, f"b{blk}.out")], got[(0, f"b{blk}.out")]
    assert a.shape == b.shape
    assert float((a - b).abs().max()) <= 2.0 ** -5 * max(1.0, float(a.abs().max()))
    assert float((a - b).abs().mean()) <= 2.0 ** -9 * max(1.0, float(a.abs().max()))


@pytest.mark.parametrize("blk", [0, 1, 4, 9, 15])
def test_fused_se_tail_matches_se_kernels(setup, blk):
    """The SE excite variants agree: two kernels (0), one k_se_excite launch (1), 8-CTA cluster kernel (2, default)."""
    e, sd, x, ref, taps = setup
    xn = x[:5].permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    got = {}
    e.set_option("no_gated_w", 1)       # (the per-image gated weights exist only with the cluster kernel: compare like with like)
    for mode in (0, 1, 2):
        e.set_option("se_mode", mode)
        e.set_tap(f"b{blk}.out")
        e.effnet_forward(xn)
        got[mode] = e.activation(f"b{blk}.out").cpu()
    e.set_option("se_mode", 2)
    e.set_option("no_gated_w", 0)
    e.set_tap("")
    for mode in (1, 2):
        assert float((got[0] - got[mode]).abs().max()) <= 2.0 ** -5 * max(1.0, float(got[0].abs().max()))
        assert float((got[0] - got[mode]).abs().mean()) <= 2.0 ** -9 * max(1.0, float(got[0].abs().max()))


def test_fused_paths_logits_close(setup):
    e, sd, x, ref, taps = setup
    xn = x.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    e.set_option("no_fuse", 1); e.set_option("se_mode", 0)
    a = e.effnet_forward(xn).cpu()
    e.set_option("no_fuse", 0); e.set_option("se_mode", 2)
    b = e.effnet_forward(xn).cpu()
    c = e.effnet_forward(xn).cpu()
    print("fused vs unfused max |dlogit|", float((a - b).abs().max()))
    assert torch.equal(b, c)                           # the fused path is deterministic (fixed-order squeeze sums)
    # the two paths round at different points; both must sit equally close to the fp32 oracle
    ea, eb = float((a - ref.flatten()).abs().mean()), float((b - ref.flatten()).abs().mean())
    print("mean |dlogit| vs fp32 oracle: unfused", ea, "fused", eb)
    assert eb <= 1.5 * ea + 0.02


@pytest.mark.parametrize("blk", [0, 1, 2, 3, 4])
def test_gated_weight_project_matches_gated_activation(setup, blk):
    """Blocks 0-4: the project GEMM with per-image SE-gated weights (gemm A_IMG) against the A_SCALE path that gates the
    activation tile; the two differ only in where the bf16 rounding of the gate product happens."""
    e, sd, x, ref, taps = setup
    xn = x[:5].permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    got = {}
    for mode in (1, 0):
        e.set_option("no_gated_w", mode)
        e.set_tap(f"b{blk}.out")
        e.effnet_forward(xn)
        got[mode] = e.activation(f"b{blk}.out").cpu()
    e.set_option("no_gated_w", 0)
    e.set_tap("")
    d = (got[0] - got[1]).abs()
    scale = max(1.0, float(got[1].abs().max()))
    print("gated-w vs gated-a: max", float(d.max()), "mean", float(d.mean()), "scale", scale)
    assert float(d.max()) <= 2.0 ** -5 * scale and float(d.mean()) <= 2.0 ** -9 * scale


def test_batch_invariance_bf16_small_batches(setup):
    """bf16 path: an image's logit does not depend on the batch it is in.  Batches of 1 and 2 take the chunk-split items of
    the fused MBConv kernel (an (image, tile) is split across CTAs so one image still fills the chip), ragged GEMM tiles
    and partially filled SE clusters -- all of which must be bit-identical to the batch-24 result."""
    e, sd, x, ref, taps = setup
    xn = x.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    full = e.effnet_forward(xn).cpu()
    for i in (0, 5, 23):
        one = e.effnet_forward(xn[i:i + 1].contiguous()).cpu()
        assert torch.equal(one, full[i:i + 1]), (i, float(one), float(full[i]))
    two = e.effnet_forward(xn[7:9].contiguous()).cpu()
    assert torch.equal(two, full[7:9])
    seven = e.effnet_forward(xn[10:17].contiguous()).cpu()
    assert torch.equal(seven, full[10:17])
