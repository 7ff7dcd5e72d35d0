"""Generate golden fixtures from the UNMODIFIED reference (build container only).

Run from the repo root in the container that has /root/reference mounted:

    python tests/golden/make_golden.py

It imports the reference's own modules (``frame_analysis`` as is;
``deepfake_detection`` behind three stub modules for packages that are absent
from the image: facenet_pytorch, pytorch_grad_cam, efficientnet_pytorch -- the
stubs only satisfy the module-level constructor calls, none of the golden
values passes through them) and records what the reference computes on seeded
synthetic inputs.  The fixtures are small JSON files; the inputs are
regenerated from the seeds by ``dfd_b200.synth`` at test time.

/root/reference does not exist on the GPU box; tests only read the JSON.
"""
import contextlib
import hashlib
import io
import json
import os
import sys
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import dfd_b200  # noqa: E402
from dfd_b200 import synth  # noqa: E402


def import_reference():
    import torch.nn as nn

    for name in ("facenet_pytorch", "pytorch_grad_cam", "pytorch_grad_cam.utils",
                 "pytorch_grad_cam.utils.model_targets", "pytorch_grad_cam.utils.image"):
        sys.modules[name] = mock.MagicMock()

    class _FakeEff(nn.Module):
        def __init__(self):
            super().__init__()
            self._fc = nn.Linear(1280, 1000)

        @classmethod
        def from_pretrained(cls, name):
            return cls()

        @classmethod
        def from_name(cls, name):
            return cls()

    eff = types.ModuleType("efficientnet_pytorch")
    eff.EfficientNet = _FakeEff
    sys.modules["efficientnet_pytorch"] = eff
    with contextlib.redirect_stdout(io.StringIO()):
        import frame_analysis
        import deepfake_detection
    return frame_analysis, deepfake_detection


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


FORENSIC_CASES = [
    # family, H, W, seed, n_frames
    ("uniform", 720, 1280, 11, 14),
    ("pink", 720, 1280, 12, 14),
    ("blur", 720, 1280, 13, 14),
    ("flat", 720, 1280, 14, 14),
    ("gradient", 720, 1280, 15, 14),
    ("pink", 1080, 1920, 16, 7),
    ("gradient", 1080, 1920, 17, 4),
    ("uniform", 2160, 3840, 18, 2),
    ("pink", 480, 640, 19, 4),
    ("blur", 120, 160, 20, 4),
    ("pink", 333, 517, 21, 4),
    ("gradient", 256, 256, 22, 3),
]


def gen_forensics(frame_analysis, dd):
    out = []
    for fam, h, w, seed, n in FORENSIC_CASES:
        frames = synth.make_sequence(fam, h, w, n, seed=seed)
        # cadence of DeepfakeDetector.analyze_frame_forensics as /analyze drives
        # it (backend_server.py:148,156): frame_count is read BEFORE increment.
        det_frame_count = 0
        an = frame_analysis.FrameForensicAnalyzer(analysis_size=(256, 256))
        rec = []
        for f in frames:
            full = det_frame_count % 3 == 0
            r = an.analyze(f) if full else an.analyze_fast(f)
            det_frame_count += 1
            rec.append({"full": full, "scores": r["scores"], "fake_probability": r["fake_probability"],
                        "frame_number": r["frame_number"], "analysis_type": r["analysis_type"]})
        out.append({"family": fam, "h": h, "w": w, "seed": seed, "n": n,
                    "first_frame_sha1": sha(frames[0]), "frames": rec})
    return out


def gen_tracker(dd):
    rng = np.random.RandomState(77)
    cases = []
    specs = [(10, 0.5), (10, 0.55), (5, 0.75), (10, 0.75)]
    for ci in range(12):
        vw, thr = specs[ci % len(specs)]
        n = int(rng.randint(5, 90))
        kind = ci % 3
        if kind == 0:
            probs = rng.uniform(0, 1, n)
        elif kind == 1:   # straddle the threshold by +-1e-7 and hit it exactly
            probs = thr + rng.choice([-1e-7, 0.0, 1e-7, -0.2, 0.2], n)
        else:             # float32-origin probabilities, as sigmoid().item() yields
            probs = rng.uniform(0, 1, n).astype(np.float32).astype(np.float64)
        with contextlib.redirect_stdout(io.StringIO()):
            t = dd.TemporalTracker(window_size=60, voting_window=vw, detection_threshold=thr)
            steps = []
            for p in probs:
                t.update(float(p))
                vs = t.get_voting_stats()
                steps.append({"verdict": t.get_confidence_level(), "fake": vs["fake_count"],
                              "real": vs["real_count"], "avg": t.get_temporal_average(),
                              "stab": t.get_stability_score()})
        cases.append({"voting_window": vw, "threshold": thr,
                      "probs": [float(p) for p in probs], "steps": steps})
    return cases


CROP_CASES = [(300, 300, 31), (97, 83, 32), (50, 71, 33), (400, 96, 34), (64, 64, 35),
              (223, 410, 36), (900, 700, 37), (40, 40, 38), (161, 159, 39), (8, 8, 40)]


def gen_clahe(dd):
    with contextlib.redirect_stdout(io.StringIO()):
        det = dd.DeepfakeDetector(use_tta=False, num_tta_augmentations=1, detection_threshold=0.55)
    out = []
    for h, w, seed in CROP_CASES:
        rng = np.random.RandomState(seed)
        fam = synth.FAMILIES[seed % 3]       # uniform / pink / blur
        crop = synth.make_frame(fam, h, w, rng)
        res = det.preprocess_face_quality(crop)
        heur = [float(det.apply_heuristics(p, crop)) for p in (0.0, 0.3, 0.5499999, 0.95, 1.0)]
        out.append({"h": h, "w": w, "seed": seed, "family": fam, "in_sha1": sha(crop),
                    "out_sha1": sha(res), "out_sum": int(res.astype(np.int64).sum()),
                    "out_head": res.reshape(-1)[:24].tolist(), "heuristics": heur})
    return out


TTA_CASES = [(300, 300, 51, 3, 7), (97, 83, 52, 3, 8), (50, 71, 53, 5, 9), (400, 96, 54, 2, 10), (223, 410, 55, 4, 11),
             (64, 64, 56, 3, 12)]


def gen_tta(dd):
    """What the unmodified analyze_face (use_tta=True) hands to _single_prediction, under random.seed(py_seed)."""
    import random
    out = []
    for h, w, seed, n_pred, py_seed in TTA_CASES:
        with contextlib.redirect_stdout(io.StringIO()):
            det = dd.DeepfakeDetector(use_tta=True, num_tta_augmentations=n_pred, detection_threshold=0.55)
        rng = np.random.RandomState(seed)
        fam = synth.FAMILIES[seed % 3]
        crop = synth.make_frame(fam, h, w, rng)
        seen = []
        det._single_prediction = lambda face: (seen.append(sha(face)), 0.25 + 0.125 * len(seen))[1]
        random.seed(py_seed)
        with contextlib.redirect_stdout(io.StringIO()):
            p, _, _ = det.analyze_face(crop)
        random.seed(py_seed)
        params = []
        for _ in range(n_pred - 1):       # the reference's draw order (deepfake_detection.py:422-430)
            params.append([random.random() > 0.5, random.uniform(0.9, 1.1), random.uniform(-3, 3)])
        out.append({"h": h, "w": w, "seed": seed, "family": fam, "n_pred": n_pred, "py_seed": py_seed, "in_sha1": sha(crop),
                    "params": params, "face_sha1": seen, "mean_of_stub_predictions": float(p)})
    return out


OVERLAY_CASES = [
    # h, w, seed, box, fake_prob, verdict, votes (fake, real, total)
    (360, 640, 61, (200, 100, 150, 160), 0.8312, "FAKE", (7, 3, 10)),
    (360, 640, 62, (5, 12, 90, 90), 0.1249, "REAL", (0, 4, 4)),
    (480, 640, 63, (500, 400, 200, 120), 0.505, "UNCERTAIN", (0, 0, 0)),
    (720, 1280, 64, (-20, 300, 100, 100), 0.995, "FAKE", (10, 0, 10)),
    (240, 320, 65, (100, 20, 40, 50), 0.0, "REAL", (1, 9, 10)),
]


def gen_overlay(dd):
    """draw_detection_overlay / _draw_frame_analysis_overlay of the unmodified reference on seeded frames."""
    out = []
    with contextlib.redirect_stdout(io.StringIO()):
        det = dd.DeepfakeDetector(use_tta=False, num_tta_augmentations=1, detection_threshold=0.55)
    for h, w, seed, box, p, verdict, votes in OVERLAY_CASES:
        rng = np.random.RandomState(seed)
        frame = synth.make_frame("pink", h, w, rng)
        det.temporal_tracker.get_voting_stats = lambda v=votes: {"fake_count": v[0], "real_count": v[1], "total_frames": v[2]}
        a = det.draw_detection_overlay(frame.copy(), *box, p, verdict)
        fres = {"scores": {"frequency": 0.25, "noise": 0.5, "ela": 0.15, "edge": 0.65, "color": 0.1, "temporal": 0.0}}
        b = det._draw_frame_analysis_overlay(frame.copy(), p, verdict, fres)
        out.append({"h": h, "w": w, "seed": seed, "box": list(box), "fake_prob": p, "verdict": verdict, "votes": list(votes),
                    "in_sha1": sha(frame), "detection_sha1": sha(a), "frame_sha1": sha(b),
                    "detection_changed": int((a != frame).any(axis=2).sum()), "frame_changed": int((b != frame).any(axis=2).sum())})
    return out


def main():
    fa, dd = import_reference()
    import cv2
    import PIL
    meta = {"cv2": cv2.__version__, "numpy": np.__version__, "PIL": PIL.__version__,
            "generator": "tests/golden/make_golden.py", "reference": REF}
    for name, data in (("forensics", gen_forensics(fa, dd)), ("tracker", gen_tracker(dd)),
                       ("clahe", gen_clahe(dd)), ("tta", gen_tta(dd)), ("overlay", gen_overlay(dd))):
        with open(os.path.join(HERE, f"{name}.json"), "w") as f:
            json.dump({"meta": meta, "cases": data}, f, indent=0)
        print(name, "ok")


if __name__ == "__main__":
    main()
