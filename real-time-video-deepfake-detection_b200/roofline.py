"""Algorithmic HBM bytes per kernel launch.

Layer-granular accounting: each kernel reads its input tensor once and writes its
output tensor once (bf16 = 2 B, fp32 = 4 B per element), plus its weights once.
"""
from . import arch


def effnet_bytes(label, m, esz=2, tc=True):
    """Bytes one launch of the kernel labelled `label` (csrc/effnet.cu) must move for batch m."""
    if label == "stem":
        return m * (224 * 224 * 3 + 112 * 112 * 32) * esz + 27 * 32 * 4
    if label == "head":
        return m * 49 * (320 + 1280) * esz + 1280 * 320 * esz
    if label == "pool":
        return m * 49 * 1280 * esz + m * 1280 * 4
    if label == "fc":
        return (512 * 1280 + 256 * 512 + 256) * 4 + m * 1280 * 4
    if not label.startswith("b"):
        return None
    i, kind = label[1:].split(".")
    b = arch.BLOCKS[int(i)]
    min_, mout = m * b.hin * b.hin, m * b.hout * b.hout
    if kind == "expand":
        return (min_ * b.cin + min_ * b.cexp) * esz + b.cexp * b.cin * esz
    if kind == "front":     # fused expand 1x1 + depthwise: block input in, depthwise output out; the expanded tensor stays on chip
        return (min_ * b.cin + mout * b.cexp) * esz + b.cexp * b.cin * esz + b.k * b.k * b.cexp * 4 + m * b.cexp * 4
    if kind == "dw":
        return (min_ * b.cexp + mout * b.cexp) * esz + b.k * b.k * b.cexp * 4 + m * b.cexp * 4
    if kind == "se":
        return 2 * m * b.cexp * 4 + 2 * b.se * b.cexp * 4
    if kind == "scale":
        return 2 * mout * b.cexp * esz
    if kind == "project":
        skip = b.s == 1 and b.cin == b.cout
        return (mout * b.cexp + mout * b.cout * (2 if skip else 1)) * esz + b.cout * b.cexp * esz
    return None


def kernel_bytes(name, n_frames, H, W, m_boxes, mean_box_area, esz=2):
    """name = 'kernel:label' as reported by Engine.profile_stop()."""
    kernel, _, label = name.partition(":")
    if label:
        if kernel == "k_scale":
            label = label.replace("project", "scale")
        return effnet_bytes(label, m_boxes, esz)
    T = 256 * 256
    if kernel == "k_resize256":       # unique 32-B sectors of the 2 tap rows per output row + tile/gray writes
        rows = min(512, H)
        return n_frames * (rows * W * 3 + T * 4)
    if kernel == "k_tile_stats":
        return n_frames * (T * 4 + T * 2)           # tile + gray in, prev gray in/out
    if kernel == "k_canny":
        return n_frames * T
    if kernel == "k_ela":
        return n_frames * T * 3 * 2
    if kernel == "k_fft_rows":
        return n_frames * (T + 129 * 256 * 8)
    if kernel == "k_fft_cols":
        return n_frames * 129 * 256 * 8
    if kernel == "k_clahe_lut":
        return int(m_boxes * mean_box_area * 3)
    if kernel == "k_clahe_hpass":
        return int(m_boxes * (mean_box_area * 3 + (mean_box_area ** 0.5) * 480))
    if kernel == "k_vpass_up_norm":
        return int(m_boxes * ((mean_box_area ** 0.5) * 480 + 224 * 224 * 3 * esz))
    if kernel == "k_pool":
        return effnet_bytes("pool", m_boxes, esz)
    if kernel == "k_fc":
        return effnet_bytes("fc", m_boxes, esz)
    return None
