"""CPU: result annotation (reference deepfake_detection.py:552-586, 688-726).

* oracle/overlay.py reproduces the unmodified reference's drawing (tests/golden/overlay.json);
* the product's rasterisation rules (csrc/overlay.cu: thick axis-aligned lines = band + diamond caps, inclusive fills, float32
  blend, stamped text masks), restated here in numpy from the SAME command list the device consumes, are bit-identical with
  OpenCV -- the CUDA kernel is then compared with the oracle on the GPU (tests/test_gpu_overlay.py)."""
import hashlib
import json
import os

import cv2
import numpy as np
import pytest

import dfd_b200  # noqa: F401
from dfd_b200 import overlay, synth
from oracle import overlay as oov

HERE = os.path.dirname(os.path.abspath(__file__))
FRES = {"scores": {"frequency": 0.25, "noise": 0.5, "ela": 0.15, "edge": 0.65, "color": 0.1, "temporal": 0.0}}


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def cases():
    with open(os.path.join(HERE, "golden", "overlay.json")) as f:
        return json.load(f)["cases"]


def apply_commands(frame, cl):
    """The kernel's per-pixel rule (csrc/overlay.cu k_overlay) in numpy."""
    cmds, masks = cl.pack()
    H, W = frame.shape[:2]
    Y, X = np.mgrid[0:H, 0:W]
    out = frame.copy()

    def thick(px, py, a, b, at, r):
        d = np.abs(py - at)
        return (d <= r) & (px >= a - (r - d)) & (px <= b + (r - d))
    for c in cmds:
        if c["op"] == overlay.MASK:
            m = np.zeros((H, W), bool)
            mk = masks[c["mask_off"]:c["mask_off"] + c["mask_w"] * c["mask_h"]].reshape(c["mask_h"], c["mask_w"]) != 0
            ys, xs = np.nonzero(mk)
            ys, xs = ys + c["y0"], xs + c["x0"]
            ok = (ys >= 0) & (ys < H) & (xs >= 0) & (xs < W)
            m[ys[ok], xs[ok]] = True
        else:
            xa, xb = min(c["x0"], c["x1"]), max(c["x0"], c["x1"])
            ya, yb = min(c["y0"], c["y1"]), max(c["y0"], c["y1"])
            if c["op"] == overlay.OUTLINE:
                r = c["thickness"] - 1
                m = thick(X, Y, xa, xb, ya, r) | thick(X, Y, xa, xb, yb, r) | thick(Y, X, ya, yb, xa, r) | thick(Y, X, ya, yb, xb, r)
            else:
                m = (X >= xa) & (X <= xb) & (Y >= ya) & (Y <= yb)
        col = c["color"][:3].astype(np.float32)
        if c["op"] == overlay.BLEND:
            v = np.rint(col * np.float32(c["alpha"]) + out.astype(np.float32) * np.float32(c["beta"]))
            out[m] = np.clip(v, 0, 255).astype(np.uint8)[m]
        else:
            out[m] = c["color"][:3]
    return out


def test_oracle_overlay_reproduces_the_reference():
    for c in cases():
        frame = synth.make_frame("pink", c["h"], c["w"], np.random.RandomState(c["seed"]))
        assert sha(frame) == c["in_sha1"]
        votes = {"fake_count": c["votes"][0], "real_count": c["votes"][1], "total_frames": c["votes"][2]}
        a = oov.draw_detection_overlay(frame.copy(), *c["box"], c["fake_prob"], c["verdict"], votes)
        b = oov.draw_frame_analysis_overlay(frame.copy(), c["fake_prob"], c["verdict"], FRES)
        assert sha(a) == c["detection_sha1"] and sha(b) == c["frame_sha1"]


def test_command_list_rules_match_opencv():
    rng = np.random.RandomState(8)
    for c in cases():
        frame = synth.make_frame("pink", c["h"], c["w"], np.random.RandomState(c["seed"]))
        votes = {"fake_count": c["votes"][0], "real_count": c["votes"][1], "total_frames": c["votes"][2]}
        cl = overlay.CommandList(c["h"], c["w"])
        overlay.detection_overlay(cl, *c["box"], c["fake_prob"], c["verdict"], votes)
        assert sha(apply_commands(frame, cl)) == c["detection_sha1"]
        cl = overlay.CommandList(c["h"], c["w"])
        overlay.frame_analysis_overlay(cl, c["fake_prob"], c["verdict"], FRES)
        assert sha(apply_commands(frame, cl)) == c["frame_sha1"]
    # random boxes, partly outside the frame, every verdict, text clipped at the borders
    for it in range(150):
        H, W = int(rng.randint(120, 400)), int(rng.randint(160, 500))
        frame = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
        box = (int(rng.randint(-40, W)), int(rng.randint(-40, H)), int(rng.randint(1, 300)), int(rng.randint(1, 300)))
        p = float(rng.uniform(0, 1))
        verdict = ["FAKE", "REAL", "UNCERTAIN"][it % 3]
        votes = {"fake_count": int(rng.randint(0, 11)), "real_count": int(rng.randint(0, 11)), "total_frames": int(rng.randint(0, 11))}
        cl = overlay.CommandList(H, W)
        overlay.detection_overlay(cl, *box, p, verdict, votes)
        assert np.array_equal(apply_commands(frame, cl), oov.draw_detection_overlay(frame.copy(), *box, p, verdict, votes)), (it, box)
        cl = overlay.CommandList(H, W)
        overlay.frame_analysis_overlay(cl, p, verdict, FRES)
        assert np.array_equal(apply_commands(frame, cl), oov.draw_frame_analysis_overlay(frame.copy(), p, verdict, FRES)), it
