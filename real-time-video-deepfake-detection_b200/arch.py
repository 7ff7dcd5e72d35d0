"""EfficientNet-B0 layer plan as the reference's checkpoint lays it out.

The reference's face classifier is ``efficientnet_pytorch.EfficientNet`` B0
with a replaced ``_fc`` (reference model.py:36-61); its ``state_dict`` keys are
lukemelas' names under the ``net.`` prefix.  This module
is the single host-side description of that plan used by the weight packer,
the synthetic-weight generator and the launch planner.
"""
import math
from collections import namedtuple

Block = namedtuple("Block", "k s cin cexp cout se hin hout pad_lo")

_RAW = (
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

IMG = 224
STEM_OUT = 32
HEAD_IN, HEAD_OUT = 320, 1280
FC_DIMS = (1280, 512, 256, 1)
BN_EPS = 1e-3       # lukemelas backbone BatchNorm2d
FC_BN_EPS = 1e-5    # nn.BatchNorm1d default in the custom _fc (model.py:53,57)


def same_pad(size, k, s):
    """Static TF-'SAME' padding (lo on top/left, hi on bottom/right)."""
    out = math.ceil(size / s)
    pad = max((out - 1) * s + k - size, 0)
    return pad // 2, pad - pad // 2


def _plan():
    h = math.ceil(IMG / 2)          # stem stride 2 -> 112
    out = []
    for (k, s, cin, cexp, cout, se) in _RAW:
        ho = math.ceil(h / s)
        out.append(Block(k, s, cin, cexp, cout, se, h, ho, same_pad(h, k, s)[0]))
        h = ho
    return tuple(out)


BLOCKS = _plan()
STEM_PAD_LO = same_pad(IMG, 3, 2)[0]     # 0 (pad 0 top/left, 1 bottom/right)


def state_dict_spec():
    """[(key, shape)] for every tensor of the reference checkpoint, in the
    order torch would register them."""
    spec = []

    def bn(prefix, c):
        spec.extend([(prefix + ".weight", (c,)), (prefix + ".bias", (c,)),
                     (prefix + ".running_mean", (c,)), (prefix + ".running_var", (c,)),
                     (prefix + ".num_batches_tracked", ())])

    spec.append(("net._conv_stem.weight", (STEM_OUT, 3, 3, 3)))
    bn("net._bn0", STEM_OUT)
    for i, b in enumerate(BLOCKS):
        p = f"net._blocks.{i}."
        if b.cexp != b.cin:
            spec.append((p + "_expand_conv.weight", (b.cexp, b.cin, 1, 1)))
            bn(p + "_bn0", b.cexp)
        spec.append((p + "_depthwise_conv.weight", (b.cexp, 1, b.k, b.k)))
        bn(p + "_bn1", b.cexp)
        spec.append((p + "_se_reduce.weight", (b.se, b.cexp, 1, 1)))
        spec.append((p + "_se_reduce.bias", (b.se,)))
        spec.append((p + "_se_expand.weight", (b.cexp, b.se, 1, 1)))
        spec.append((p + "_se_expand.bias", (b.cexp,)))
        spec.append((p + "_project_conv.weight", (b.cout, b.cexp, 1, 1)))
        bn(p + "_bn2", b.cout)
    spec.append(("net._conv_head.weight", (HEAD_OUT, HEAD_IN, 1, 1)))
    bn("net._bn1", HEAD_OUT)
    spec.append(("net._fc.1.weight", (512, 1280)))
    spec.append(("net._fc.1.bias", (512,)))
    bn("net._fc.2", 512)
    spec.append(("net._fc.5.weight", (256, 512)))
    spec.append(("net._fc.5.bias", (256,)))
    bn("net._fc.6", 256)
    spec.append(("net._fc.9.weight", (1, 256)))
    spec.append(("net._fc.9.bias", (1,)))
    return spec
