"""GPU parity: test-time augmentation and calibration (reference deepfake_detection.py:408-455)
against the oracle (oracle/faceprep.py ``tta_faces``, pinned to the unmodified reference by tests/golden/tta.json)."""
import random

import numpy as np
import pytest
import torch

import dfd_b200  # noqa: F401
from dfd_b200 import synth, tta
from oracle import effnet as oeff, faceprep

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from dfd_b200.engine import Engine
    e = Engine(device=0, max_streams=8, max_batch=48, max_crop=1024)
    e.load_state_dict(synth.make_state_dict())
    yield e
    e.close()


BOXES = [(100, 50, 300, 300), (0, 0, 97, 83), (400, 300, 50, 71), (640, 10, 400, 96), (1200, 650, 40, 40),
         (800, 200, 223, 410), (20, 20, 900, 700), (1250, 700, 100, 100)]          # the last one is clamped to 30 x 20


@pytest.mark.parametrize("n_pred", [2, 3, 5])
def test_tta_crops_bit_exact_and_probability(eng, n_pred):
    rng = np.random.RandomState(21)
    frame = synth.make_frame("pink", 720, 1280, rng)
    ft = torch.from_numpy(frame).cuda().unsqueeze(0)
    boxes = np.array(BOXES, np.int32)
    random.seed(100 + n_pred)
    params = [tta.draw_params(n_pred) for _ in BOXES]
    params[0][0] = (True, 1.1, 3.0)                     # extremes of the three ranges
    params[1][0] = (False, 0.9, -3.0)
    out = eng.face_prep_tta(ft, boxes, np.zeros(len(BOXES), np.int32), params, "fp32")
    logits = eng.effnet_forward(out)
    prob = eng.face_probability_tta(logits, eng._dev(boxes, torch.int32), n_pred).cpu().numpy()
    sd = synth.make_state_dict()
    for i, (x, y, w, h) in enumerate(BOXES):
        crop = frame[y:y + h, x:x + w]
        faces = faceprep.tta_faces(crop, params[i])
        preds = []
        for j, f in enumerate(faces):
            q = i * n_pred + j
            ref160 = faceprep.resize160(f)
            assert np.array_equal(eng.dbg_face160(q).cpu().numpy(), ref160), (i, j, "augmented 160x160 crop not bit-exact")
            ref = faceprep.to_input(ref160)
            assert np.abs(out[q].cpu().numpy() - ref[0].permute(1, 2, 0).numpy()).max() < 2e-6
            preds.append(torch.sigmoid(oeff.forward(ref, sd).squeeze()).item())
        ch, cw = crop.shape[:2]
        exp = float(faceprep.heuristics(np.mean(preds), ch, cw))                   # deepfake_detection.py:441, 489-502
        # heuristics on the device use the box handed to face_probability_tta: hand it the clamped size like the detector does
        got = float(eng.face_probability_tta(logits[i * n_pred:(i + 1) * n_pred], np.array([[0, 0, cw, ch]], np.int32), n_pred).cpu()[0])
        assert abs(got - exp) <= 1e-4, (i, got, exp)
    assert np.isfinite(prob).all()


@pytest.mark.parametrize("n_pred", [1, 2, 3, 7, 8, 9, 16])
def test_tta_mean_is_numpys_mean(eng, n_pred):
    """np.mean over the per-prediction probabilities in float64, NumPy's summation order: exact."""
    g = torch.Generator().manual_seed(n_pred)
    m = 6
    logits = (torch.randn(m * n_pred, generator=g) * 3).cuda()
    big = np.tile(np.array([[0, 0, 200, 200]], np.int32), (m * n_pred, 1))
    single = eng.face_probability(logits, big).cpu().numpy()                         # the device's own float32 sigmoids
    boxes = np.array([[0, 0, 200, 200], [0, 0, 79, 200], [0, 0, 200, 79], [0, 0, 80, 80], [0, 0, 8, 8], [0, 0, 500, 90]], np.int32)
    got = eng.face_probability_tta(logits, boxes, n_pred).cpu().numpy()
    for i in range(m):
        exp = faceprep.heuristics(np.mean([float(v) for v in single[i * n_pred:(i + 1) * n_pred]]), boxes[i, 3], boxes[i, 2])
        assert got[i] == exp, (i, got[i], exp)


def test_calibrators(eng):
    sk = pytest.importorskip("sklearn.linear_model")
    rng = np.random.RandomState(3)
    raw = rng.uniform(0, 1, 400)
    y = (raw + rng.normal(0, 0.2, 400) > 0.5).astype(int)
    lr = sk.LogisticRegression().fit(raw.reshape(-1, 1), y)
    logits = torch.from_numpy(rng.normal(0, 2, 64).astype(np.float32)).cuda()
    boxes = np.tile(np.array([[0, 0, 100, 100]], np.int32), (64, 1))
    boxes[::5, 2] = 60                                   # heuristic branch after the calibration
    p_raw = eng.face_probability(logits, np.tile(np.array([[0, 0, 100, 100]], np.int32), (64, 1))).cpu().numpy()
    try:
        eng.set_calibrator("logistic", [float(lr.coef_[0][0])], [float(lr.intercept_[0])])
        got = eng.face_probability(logits, boxes).cpu().numpy()
        for i in range(64):
            exp = faceprep.heuristics(lr.predict_proba([[p_raw[i]]])[0][1], boxes[i, 3], boxes[i, 2])
            assert abs(got[i] - exp) <= 1e-14, (i, got[i], exp)
        xs = np.sort(rng.uniform(0, 1, 17)); ys = np.sort(rng.uniform(0, 1, 17))
        eng.set_calibrator("piecewise_linear", xs, ys)
        got = eng.face_probability(logits, boxes).cpu().numpy()
        for i in range(64):
            exp = faceprep.heuristics(np.interp(p_raw[i], xs, ys), boxes[i, 3], boxes[i, 2])
            assert abs(got[i] - exp) <= 1e-15, (i, got[i], exp)
    finally:
        eng.set_calibrator("none")
    assert np.array_equal(eng.face_probability(logits, np.tile(np.array([[0, 0, 100, 100]], np.int32), (64, 1))).cpu().numpy(), p_raw)


def test_detector_tta_and_calibrator_like_the_reference():
    """DeepfakeDetector(use_tta=True): same random draws as the reference (global `random`), probability within the fp32
    gate of the oracle; a LogisticRegression calibrator runs on the device, an opaque one on the host like the reference."""
    from dfd_b200 import deepfake_detection as dd
    sd = synth.make_state_dict()
    dd.load_model_weights(sd)
    det = dd.DeepfakeDetector(use_tta=True, num_tta_augmentations=3, detection_threshold=0.55)
    rng = np.random.RandomState(9)
    for (h, w) in [(220, 180), (64, 90)]:
        crop = synth.make_frame("pink", h, w, rng)
        random.seed(42)
        p, p2, cam = det.analyze_face(crop)
        after = random.random()
        random.seed(42)
        params = tta.draw_params(3)
        assert random.random() == after                                  # consumed exactly the reference's six draws
        preds = [torch.sigmoid(oeff.forward(faceprep.to_input_noclahe(f), sd).squeeze()).item()
                 for f in faceprep.tta_faces(crop, params)]
        exp = float(faceprep.heuristics(np.mean(preds), h, w))
        assert isinstance(p, np.float64) and p == p2 and cam is None
        assert abs(p - exp) <= 1e-4, (p, exp)

        class Opaque:                                                    # any object with predict_proba, as pickled by a user
            def predict_proba(self, X):
                v = X[0][0]
                return [[1 - v * v, v * v]]
        det.set_calibrator(Opaque())
        random.seed(42)
        pc, _, _ = det.analyze_face(crop)
        expc = float(faceprep.heuristics(np.mean(preds) ** 2, h, w))
        assert abs(pc - expc) <= 2e-4
        det.set_calibrator(None)
    det.release()
