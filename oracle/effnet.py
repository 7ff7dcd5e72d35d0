"""Oracle (CPU, test infrastructure): EfficientNet-B0 as the reference uses it.

The reference wraps ``efficientnet_pytorch.EfficientNet`` (lukemelas; imported
at reference model.py:18, built at :41-43, ``_fc`` replaced at :50-61, forward
at :63-72).  That package is absent from /root/reference and from this image
(version unpinned by the reference: it is not in requirements.txt; latest
release 0.7.1), so its published algorithm is restated here in plain
``torch.nn.functional`` following SURVEY.md Appendix A:

  * every conv uses *static* TF-"SAME" padding computed for a 224 input:
    out = ceil(in/s); pad = max((out-1)*s + k - in, 0); lo = pad//2 on the
    top/left, hi = pad-lo on the bottom/right (asymmetric for stride 2);
  * BatchNorm2d eps 1e-3 (backbone), BatchNorm1d eps 1e-5 (custom ``_fc``);
  * swish = x*sigmoid(x); SE: avgpool -> reduce(+bias) -> swish ->
    expand(+bias) -> sigmoid -> scale; skip iff stride 1 and cin == cout;
  * block 0 (expand ratio 1) has no ``_expand_conv`` / ``_bn0``.

State-dict key names are lukemelas' under the ``net.`` prefix (the layout of
weights/best_model.pth, reference deepfake_detection.py:44-51).

Pinning: no numeric vectors exist in the reference's tests for this boundary
(tests/test_functional.py:70-110 pin structure only).  tests/test_oracle_effnet.py
cross-checks this restatement against HF ``transformers`` EfficientNetModel
(an independent port of the official TF EfficientNet) with weights mapped
across -- see DESIGN.md.
"""
import math

import torch
import torch.nn.functional as F

# (kernel, stride, cin, cexp, cout, se_channels) -- SURVEY.md Appendix A table
BLOCKS = (
    (3, 1, 32, 32, 16, 8),
    (3, 2, 16, 96, 24, 4),
    (3, 1, 24, 144, 24, 6),
    (5, 2, 24, 144, 40, 6),
    (5, 1, 40, 240, 40, 10),
    (3, 2, 40, 240, 80, 10),
    (3, 1, 80, 480, 80, 20),
    (3, 1, 80, 480, 80, 20),
    (5, 1, 80, 480, 112, 20),
    (5, 1, 112, 672, 112, 28),
    (5, 1, 112, 672, 112, 28),
    (5, 2, 112, 672, 192, 28),
    (5, 1, 192, 1152, 192, 48),
    (5, 1, 192, 1152, 192, 48),
    (5, 1, 192, 1152, 192, 48),
    (3, 1, 192, 1152, 320, 48),
)
BN_EPS = 1e-3
FC_BN_EPS = 1e-5


def same_pad(size, k, s):
    out = math.ceil(size / s)
    pad = max((out - 1) * s + k - size, 0)
    return pad // 2, pad - pad // 2


def _conv_same(x, w, stride, groups=1, bias=None):
    k = w.shape[-1]
    lo_h, hi_h = same_pad(x.shape[2], k, stride)
    lo_w, hi_w = same_pad(x.shape[3], k, stride)
    if lo_h or hi_h or lo_w or hi_w:
        x = F.pad(x, (lo_w, hi_w, lo_h, hi_h))
    return F.conv2d(x, w, bias, stride=stride, groups=groups)


def _bn(x, sd, prefix, eps):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, eps)


def _swish(x):
    return x * torch.sigmoid(x)


def mbconv(x, sd, i, taps=None):
    k, s, cin, cexp, cout, se = BLOCKS[i]
    p = f"net._blocks.{i}."
    inp = x
    if cexp != cin:
        x = _swish(_bn(_conv_same(x, sd[p + "_expand_conv.weight"], 1), sd, p + "_bn0", BN_EPS))
        if taps is not None:
            taps[f"b{i}.expand"] = x
    x = _swish(_bn(_conv_same(x, sd[p + "_depthwise_conv.weight"], s, groups=cexp), sd, p + "_bn1", BN_EPS))
    if taps is not None:
        taps[f"b{i}.dw"] = x
    sq = F.adaptive_avg_pool2d(x, 1)
    sq = _swish(F.conv2d(sq, sd[p + "_se_reduce.weight"], sd[p + "_se_reduce.bias"]))
    sq = F.conv2d(sq, sd[p + "_se_expand.weight"], sd[p + "_se_expand.bias"])
    x = torch.sigmoid(sq) * x
    x = _bn(_conv_same(x, sd[p + "_project_conv.weight"], 1), sd, p + "_bn2", BN_EPS)
    if s == 1 and cin == cout:
        x = x + inp
    if taps is not None:
        taps[f"b{i}.out"] = x
    return x


def features(x, sd, taps=None):
    """(B,3,224,224) normalised RGB -> (B,1280) pooled features."""
    x = _swish(_bn(_conv_same(x, sd["net._conv_stem.weight"], 2), sd, "net._bn0", BN_EPS))
    if taps is not None:
        taps["stem"] = x
    for i in range(len(BLOCKS)):
        x = mbconv(x, sd, i, taps)
    x = _swish(_bn(_conv_same(x, sd["net._conv_head.weight"], 1), sd, "net._bn1", BN_EPS))
    return F.adaptive_avg_pool2d(x, 1).flatten(1)


def classifier(f, sd):
    """Custom ``_fc`` Sequential of reference model.py:50-61 in eval mode
    (dropouts are identity)."""
    x = F.linear(f, sd["net._fc.1.weight"], sd["net._fc.1.bias"])
    x = F.relu(F.batch_norm(x, sd["net._fc.2.running_mean"], sd["net._fc.2.running_var"],
                            sd["net._fc.2.weight"], sd["net._fc.2.bias"], False, 0.0, FC_BN_EPS))
    x = F.linear(x, sd["net._fc.5.weight"], sd["net._fc.5.bias"])
    x = F.relu(F.batch_norm(x, sd["net._fc.6.running_mean"], sd["net._fc.6.running_var"],
                            sd["net._fc.6.weight"], sd["net._fc.6.bias"], False, 0.0, FC_BN_EPS))
    return F.linear(x, sd["net._fc.9.weight"], sd["net._fc.9.bias"])


@torch.no_grad()
def forward(x, sd, taps=None):
    """logits (B,1), as ``DeepfakeEfficientNet.forward`` (model.py:63-72)."""
    f = features(x, sd, taps)
    if taps is not None:
        taps["features"] = f
    return classifier(f, sd)


# ---------------------------------------------------------------------------------------------
# bf16-storage restatement (second oracle for the bf16 mode of the CUDA path)
# ---------------------------------------------------------------------------------------------
def _r(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _fold(sd, prefix, eps):
    g, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    m, v = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    s = g / torch.sqrt(v + eps)
    return s, b - m * s


@torch.no_grad()
def block_bf16_storage(x, sd, i, gated_weight=False):
    """One MBConv block of the bf16-storage restatement: x (bf16-representable fp32, NCHW) -> (block output, depthwise
    output), both rounded where the CUDA bf16 path stores them.  Used by ``forward_bf16_storage`` and, fed with the CUDA
    path's OWN previous-layer tap, by the per-block exactness test (which removes the network's error amplification from
    the comparison: what is left is the kernels' own arithmetic)."""
    k, st, cin, cexp, cout, se = BLOCKS[i]
    p = f"net._blocks.{i}."
    inp = x
    if cexp != cin:
        s, b = _fold(sd, p + "_bn0", BN_EPS)
        w = _r(sd[p + "_expand_conv.weight"] * s.view(-1, 1, 1, 1))
        x = _r(_swish(F.conv2d(x, w) + b.view(1, -1, 1, 1)))
    s, b = _fold(sd, p + "_bn1", BN_EPS)
    w = sd[p + "_depthwise_conv.weight"] * s.view(-1, 1, 1, 1)                 # fp32 weights on the CUDA cores
    y = _swish(_conv_same(x, w, st, groups=cexp) + b.view(1, -1, 1, 1))       # fp32 before it is stored
    sq = F.adaptive_avg_pool2d(y, 1)                                            # squeeze sums the unrounded values
    sq = _swish(F.conv2d(sq, sd[p + "_se_reduce.weight"], sd[p + "_se_reduce.bias"]))
    gate = torch.sigmoid(F.conv2d(sq, sd[p + "_se_expand.weight"], sd[p + "_se_expand.bias"]))
    x = _r(y)
    dw = x
    s, b = _fold(sd, p + "_bn2", BN_EPS)
    wp = (sd[p + "_project_conv.weight"] * s.view(-1, 1, 1, 1)).flatten(1)      # (cout, cexp)
    if gated_weight:
        wg = _r(wp.unsqueeze(0) * gate.flatten(1).unsqueeze(1))                 # (B, cout, cexp) per-image weights
        x = torch.einsum("bchw,boc->bohw", x, wg) + b.view(1, -1, 1, 1)
    else:
        x = F.conv2d(_r(x * gate), _r(wp).view(cout, cexp, 1, 1)) + b.view(1, -1, 1, 1)
    if st == 1 and cin == cout:
        x = x + inp
    return _r(x), dw


@torch.no_grad()
def forward_bf16_storage(x, sd, gated_weight_blocks=5):
    """The same network with every stored activation and every GEMM operand rounded to bfloat16 at exactly the points
    where the CUDA bf16 path rounds (fp32 accumulation, fp32 SE / pooling / classifier, BatchNorm folded into the
    weights before they are rounded).  It separates the two sources of a bf16-vs-fp32 difference: the INHERENT loss of
    bf16 storage (this function vs ``forward``) and the kernels' own error (CUDA bf16 logits vs this function).
    Blocks < gated_weight_blocks fold the SE gate into the project weights (rounded), the others into the activation."""
    x = _r(x)
    s, b = _fold(sd, "net._bn0", BN_EPS)
    w = _r(sd["net._conv_stem.weight"] * s.view(-1, 1, 1, 1))
    x = _r(_swish(_conv_same(x, w, 2) + b.view(1, -1, 1, 1)))
    for i in range(len(BLOCKS)):
        x, _ = block_bf16_storage(x, sd, i, gated_weight=i < gated_weight_blocks)
    s, b = _fold(sd, "net._bn1", BN_EPS)
    w = _r(sd["net._conv_head.weight"] * s.view(-1, 1, 1, 1))
    x = _r(_swish(F.conv2d(x, w) + b.view(1, -1, 1, 1)))
    return classifier(F.adaptive_avg_pool2d(x, 1).flatten(1), sd)
