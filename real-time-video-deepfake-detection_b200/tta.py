"""Host side of test-time augmentation (reference deepfake_detection.py:408-443, ``analyze_face_with_tta``).

The reference draws, per extra prediction and in this order, ``random.random() > 0.5`` (horizontal flip),
``random.uniform(0.9, 1.1)`` (brightness) and ``random.uniform(-3, 3)`` (rotation angle) from Python's global ``random``
module.  ``draw_params`` consumes the generator in exactly that order, so ``random.seed(s)`` before a call reproduces the
reference's augmentations; the pixel work (flip, convertScaleAbs, warpAffine, resize, classifier, mean) runs on the device
(``dfd_face_prep_tta`` / ``dfd_face_probability_tta``).  What stays on the host is the 2 x 3 matrix arithmetic of
``cv2.getRotationMatrix2D`` and of the inversion ``cv2.warpAffine`` performs before its loops -- a dozen double operations
per augmentation, bit-identical with OpenCV (tests/test_oracle_tta.py)."""
import math
import random as _random

import numpy as np

AUG_DTYPE = np.dtype([("flip", "<i4"), ("brightness", "<f4"), ("im", "<f8", (6,))])      # dfd_tta_aug, 56 bytes
assert AUG_DTYPE.itemsize == 56


def draw_params(num_tta_augmentations, rng=_random):
    """[(flip, brightness, angle)] for the ``num_tta_augmentations - 1`` extra predictions (deepfake_detection.py:418-430)."""
    out = []
    for _ in range(num_tta_augmentations - 1):
        flip = rng.random() > 0.5
        brightness = rng.uniform(0.9, 1.1)
        angle = rng.uniform(-3, 3)
        out.append((bool(flip), float(brightness), float(angle)))
    return out


def rotation_matrix(w, h, angle, scale=1.0):
    """cv2.getRotationMatrix2D((w/2, h/2), angle, scale) (OpenCV imgwarp.cpp: the centre is a Point2f)."""
    a = angle * (math.pi / 180)
    alpha, beta = math.cos(a) * scale, math.sin(a) * scale
    cx, cy = float(np.float32(w / 2)), float(np.float32(h / 2))
    return [alpha, beta, (1 - alpha) * cx - beta * cy, -beta, alpha, beta * cx + (1 - alpha) * cy]


def invert_affine(m):
    """The inversion cv2.warpAffine applies to M when WARP_INVERSE_MAP is not set (same operation order)."""
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m0, m1, m3, m4 = a11, m[1] * (-d), m[3] * (-d), a22
    b1 = -m0 * m[2] - m1 * m[5]
    b2 = -m3 * m[2] - m4 * m[5]
    return [m0, m1, b1, m3, m4, b2]


def pack(params, w, h):
    """dfd_tta_aug records for one w x h crop."""
    rec = np.zeros(len(params), AUG_DTYPE)
    for i, (flip, brightness, angle) in enumerate(params):
        rec[i]["flip"] = int(flip)
        rec[i]["brightness"] = np.float32(brightness)       # cv2.convertScaleAbs multiplies in float32
        rec[i]["im"] = invert_affine(rotation_matrix(w, h, angle))
    return rec
