"""Checkpoint loading and parameter packing for libdfd (host side, load time).

Accepts what the reference accepts (deepfake_detection.py:35-76): a path or
dict that is either ``{'model_state_dict': ...}`` (train.py:1034-1055) or a
bare state_dict with lukemelas ``net.*`` keys, non-strict.  BatchNorm is folded
into the preceding conv / linear in float64 and the result is packed into the
float32 blob layout of csrc/effnet_plan.h (``dfd_load_weights``).
"""
import numpy as np
import torch

from . import arch

ALIGN = 64


def _align(v):
    return (v + ALIGN - 1) // ALIGN * ALIGN


def blob_layout():
    """[(name, offset, n)] in blob order + total size; mirrors eff_offsets()."""
    items, p = [], 0

    def take(name, n):
        nonlocal p
        items.append((name, p, n))
        p = _align(p + n)

    take("stem.w", 27 * 32)
    take("stem.b", 32)
    for i, b in enumerate(arch.BLOCKS):
        if b.cexp != b.cin:
            take(f"b{i}.we", b.cexp * b.cin)
            take(f"b{i}.be", b.cexp)
        take(f"b{i}.wd", b.k * b.k * b.cexp)
        take(f"b{i}.bd", b.cexp)
        take(f"b{i}.wr", b.se * b.cexp)
        take(f"b{i}.br", b.se)
        take(f"b{i}.wx", b.cexp * b.se)
        take(f"b{i}.bx", b.cexp)
        take(f"b{i}.wp", b.cout * b.cexp)
        take(f"b{i}.bp", b.cout)
    take("head.w", 1280 * 320)
    take("head.b", 1280)
    take("fc1.w", 512 * 1280)
    take("fc1.b", 512)
    take("fc2.w", 256 * 512)
    take("fc2.b", 256)
    take("fc3.w", 256)
    take("fc3.b", 1)
    return items, p


def extract_state_dict(obj):
    """Checkpoint container -> state_dict (deepfake_detection.py:44-51)."""
    if isinstance(obj, (str, bytes)):
        obj = torch.load(obj, map_location="cpu", weights_only=False)
    if isinstance(obj, dict) and "model_state_dict" in obj:
        obj = obj["model_state_dict"]
    return obj


def check_keys(sd):
    """(missing, unexpected) against the reference layout, like load_state_dict(strict=False)."""
    want = [k for k, _ in arch.state_dict_spec()]
    missing = [k for k in want if k not in sd]
    unexpected = [k for k in sd if k not in set(want)]
    return missing, unexpected


def _fold(sd, conv_w, bn, eps):
    w = sd[conv_w].double()
    g, b = sd[bn + ".weight"].double(), sd[bn + ".bias"].double()
    m, v = sd[bn + ".running_mean"].double(), sd[bn + ".running_var"].double()
    s = g / torch.sqrt(v + eps)
    return w * s.view(-1, *([1] * (w.dim() - 1))), b - m * s


def pack_state_dict(sd):
    """state_dict (net.* keys, float tensors) -> np.float32 blob."""
    sd = {k: (v.detach().cpu() if torch.is_tensor(v) else torch.as_tensor(v)) for k, v in sd.items()}
    layout, total = blob_layout()
    off = {name: (o, n) for name, o, n in layout}
    blob = np.zeros(total, np.float32)

    def put(name, t):
        o, n = off[name]
        a = t.reshape(-1).to(torch.float32).numpy()
        assert a.size == n, (name, a.size, n)
        blob[o:o + n] = a

    w, b = _fold(sd, "net._conv_stem.weight", "net._bn0", arch.BN_EPS)
    put("stem.w", w.permute(2, 3, 1, 0).reshape(27, 32))
    put("stem.b", b)
    for i, blk in enumerate(arch.BLOCKS):
        p = f"net._blocks.{i}."
        if blk.cexp != blk.cin:
            w, b = _fold(sd, p + "_expand_conv.weight", p + "_bn0", arch.BN_EPS)
            put(f"b{i}.we", w.reshape(blk.cexp, blk.cin))
            put(f"b{i}.be", b)
        w, b = _fold(sd, p + "_depthwise_conv.weight", p + "_bn1", arch.BN_EPS)
        put(f"b{i}.wd", w[:, 0].permute(1, 2, 0).reshape(blk.k * blk.k, blk.cexp))
        put(f"b{i}.bd", b)
        put(f"b{i}.wr", sd[p + "_se_reduce.weight"].reshape(blk.se, blk.cexp))
        put(f"b{i}.br", sd[p + "_se_reduce.bias"])
        put(f"b{i}.wx", sd[p + "_se_expand.weight"].reshape(blk.cexp, blk.se))
        put(f"b{i}.bx", sd[p + "_se_expand.bias"])
        w, b = _fold(sd, p + "_project_conv.weight", p + "_bn2", arch.BN_EPS)
        put(f"b{i}.wp", w.reshape(blk.cout, blk.cexp))
        put(f"b{i}.bp", b)
    w, b = _fold(sd, "net._conv_head.weight", "net._bn1", arch.BN_EPS)
    put("head.w", w.reshape(1280, 320))
    put("head.b", b)
    for lin, bn, name in (("net._fc.1", "net._fc.2", "fc1"), ("net._fc.5", "net._fc.6", "fc2")):
        w = sd[lin + ".weight"].double()
        lb = sd[lin + ".bias"].double()
        g, bb = sd[bn + ".weight"].double(), sd[bn + ".bias"].double()
        m, v = sd[bn + ".running_mean"].double(), sd[bn + ".running_var"].double()
        s = g / torch.sqrt(v + arch.FC_BN_EPS)
        put(name + ".w", w * s.view(-1, 1))
        put(name + ".b", (lb - m) * s + bb)
    put("fc3.w", sd["net._fc.9.weight"].reshape(256))
    put("fc3.b", sd["net._fc.9.bias"].reshape(1))
    return blob
