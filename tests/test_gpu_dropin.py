"""GPU: the reference-facing Python surface (dfd_b200.{deepfake_detection,frame_analysis,model,backend_server})
behaves like the reference's objects.  Mirrors the reference's own tests (tests/test_functional.py,
test_algorithm.py, test_reliability.py) and adds numeric parity against the oracle."""
import io
import json
import time

import cv2
import numpy as np
import pytest
import torch

import dfd_b200  # noqa: F401
from dfd_b200 import synth
from oracle import effnet as oeff, faceprep as ofp, forensics as ofor, tracker as otr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def weights():
    from dfd_b200 import deepfake_detection as dd
    sd = synth.make_state_dict()
    dd.load_model_weights(sd)
    return sd


# ---- TemporalTracker (reference tests/test_functional.py:223-305, test_algorithm.py:50-155,251-278) ----------
def test_tracker_rules():
    from dfd_b200.deepfake_detection import TemporalTracker
    t = TemporalTracker(window_size=60, voting_window=10, detection_threshold=0.75)
    assert t.get_confidence_level() == "UNCERTAIN"
    for _ in range(9):
        t.update(0.9)
    assert t.get_confidence_level() == "UNCERTAIN"
    t.update(0.9)
    assert t.get_confidence_level() == "FAKE"
    t.reset()
    assert t.get_confidence_level() == "UNCERTAIN" and len(t.frame_classifications) == 0 and len(t.score_history) == 0
    for _ in range(6):
        t.update(0.9)
    for _ in range(4):
        t.update(0.2)
    assert t.get_voting_stats() == {"fake_count": 6, "real_count": 4, "total_frames": 10}
    assert t.get_confidence_level() == "FAKE"
    t.update(None)
    assert len(t.score_history) == 10
    t.release()
    t = TemporalTracker(voting_window=5, detection_threshold=0.75)
    for _ in range(5):
        t.update(0.75)                     # exactly at threshold: strict > -> REAL
    assert t.get_confidence_level() == "REAL"
    t.release()
    t = TemporalTracker(voting_window=10, detection_threshold=0.5)
    for _ in range(5):
        t.update(0.9)
    for _ in range(5):
        t.update(0.1)
    assert t.get_confidence_level() == "REAL"          # tie -> REAL
    for s in (0.1, 0.2, 0.3, 0.4, 0.5):
        t.update(s)
    ref = otr.OracleTemporalTracker(voting_window=10, detection_threshold=0.5)
    for p in [0.9] * 5 + [0.1] * 5 + [0.1, 0.2, 0.3, 0.4, 0.5]:
        ref.update(p)
    assert t.get_temporal_average() == ref.get_temporal_average()
    assert abs(t.get_stability_score() - ref.get_stability_score()) < 1e-12
    t.release()
    t = TemporalTracker(voting_window=10, detection_threshold=0.5)
    for _ in range(30):
        t.update(0.85)
    assert t.get_stability_score() > 0.9
    t.reset()
    for i in range(30):
        t.update(0.1 if i % 2 == 0 else 0.9)
    assert t.get_stability_score() < 0.5
    t.release()


# ---- FrameForensicAnalyzer (reference tests/test_functional.py:164-216, test_algorithm.py:161-205) -----------
def test_forensic_analyzer_surface():
    from dfd_b200.frame_analysis import FrameForensicAnalyzer
    an = FrameForensicAnalyzer(analysis_size=(256, 256))
    rng = np.random.RandomState(0)
    frame = rng.randint(60, 200, (480, 640, 3)).astype(np.uint8)
    r = an.analyze(frame)
    assert set(r["scores"]) == {"frequency", "noise", "ela", "edge", "color", "temporal"}
    assert all(0.0 <= v <= 1.0 for v in r["scores"].values()) and 0.0 <= r["fake_probability"] <= 1.0
    assert r["analysis_type"] == "frame_forensic" and r["frame_number"] == 1
    manual = float(np.clip(sum(r["scores"][k] * an.weights[k] for k in an.weights), 0, 1))
    assert abs(r["fake_probability"] - manual) < 1e-6
    f = an.analyze_fast(frame)
    assert list(f["scores"]) == ["frequency", "temporal", "edge"] and f["analysis_type"] == "frame_forensic_fast"
    assert an.frame_count == 2 and an.prev_frame_gray is not None
    an.reset()
    assert an.frame_count == 0 and an.prev_frame_gray is None and len(an.temporal_diffs) == 0
    # ordinal expectations of the reference's tests
    smooth = cv2.GaussianBlur(np.full((256, 256, 3), 128, np.uint8), (31, 31), 10)
    noisy = rng.randint(60, 200, (256, 256, 3)).astype(np.uint8)
    rs = an.analyze(smooth); an.reset(); rn = an.analyze(noisy); an.reset()
    assert rs["scores"]["frequency"] >= rn["scores"]["frequency"]
    assert an.analyze(np.full((256, 256, 3), 100, np.uint8))["scores"]["color"] >= rn["scores"]["color"]
    # determinism across two fresh analyzers; reference parity on the same frames
    a2 = FrameForensicAnalyzer()
    an.reset()
    o = ofor.OracleForensicAnalyzer()
    for fr in synth.make_sequence("pink", 360, 640, 4, seed=9):
        r1, r2, ro = an.analyze(fr), a2.analyze(fr), o.analyze(fr)
        assert r1 == r2
        assert r1["scores"] == ro["scores"] and r1["fake_probability"] == ro["fake_probability"]
    for shape in ((120, 160), (1080, 1920)):
        an.reset()
        assert 0 <= an.analyze(rng.randint(0, 255, (*shape, 3)).astype(np.uint8))["fake_probability"] <= 1
    with pytest.raises(ValueError):
        an.analyze(np.zeros((10,), np.uint8))
    an.release(); a2.release()


# ---- DeepfakeEfficientNet (reference tests/test_functional.py:93-110, test_reliability.py:123-132) -----------
def test_model_forward_matches_oracle(weights):
    from dfd_b200.model import DeepfakeEfficientNet
    m = DeepfakeEfficientNet(pretrained=False).eval()
    missing, unexpected = m.load_state_dict(weights, strict=False)
    assert not missing and not unexpected
    g = torch.Generator().manual_seed(5)
    x = synth._calib_batch(g, 3).float()
    with torch.no_grad():
        out = m(x)
        out2 = m(x)
    assert out.shape == (3, 1) and torch.equal(out, out2)
    probs = torch.sigmoid(out)
    assert probs.min() >= 0 and probs.max() <= 1
    ref = oeff.forward(x, weights)
    assert float((torch.sigmoid(out.cpu()) - torch.sigmoid(ref)).abs().max()) <= 1e-4
    feats = m.extract_features(x)
    assert feats.shape == (3, 1280)
    assert float((feats.cpu() - oeff.features(x, weights)).abs().max()) < 1e-3
    # a changed parameter is picked up (weights are re-packed when the module's tensors change)
    with torch.no_grad():
        m.net._fc[9].bias.add_(1.0)
        out3 = m(x)
    assert float((out3 - out - 1.0).abs().max()) < 1e-4


# ---- DeepfakeDetector (deepfake_detection.py:292-726; reference tests/test_reliability.py:254-269) -----------
def test_detector_predict_and_reset_against_oracle(weights):
    from dfd_b200.deepfake_detection import DeepfakeDetector
    det = DeepfakeDetector(use_tta=False, num_tta_augmentations=1, detection_threshold=0.55)
    assert det.full_forensic_interval == 3 and det.frame_count == 0 and det.last_frame_forensic_result is None
    frames = synth.make_sequence("pink", 720, 1280, 12, seed=21)
    box = (400, 200, 300, 280)
    o_an, o_tr = ofor.OracleForensicAnalyzer(), otr.OracleTemporalTracker(detection_threshold=0.55)
    count = 0
    for i, f in enumerate(frames):
        f = np.array(f)                                  # predict() annotates the caller's array in place, like the reference
        clean = f.copy()
        out_frame, trig, ff, res = det.predict(f, faces=[box] if i % 4 != 3 else [])
        assert out_frame is f and (f != clean).any()
        f = clean
        count += 1
        exp_f = o_an.analyze(f) if count % 3 == 0 else o_an.analyze_fast(f)       # predict() increments first (:597-600)
        assert res["frame_forensic"]["scores"] == exp_f["scores"]
        assert res["frame_forensic"]["fake_probability"] == exp_f["fake_probability"]
        if i % 4 != 3:
            p = float(torch.sigmoid(oeff.forward(ofp.prepare(f, box), weights)).item())
            p = ofp.heuristics(p, box[3], box[2])
            assert abs(res["face_results"][0]["face_prob"] - p) <= 1e-4
            assert res["analysis_mode"] == "face+frame" and res["faces_detected"] == 1
            o_tr.update(np.float64(res["face_results"][0]["face_prob"]))           # same value -> verdicts must agree exactly
        else:
            assert res["analysis_mode"] == "frame_only" and res["face_results"] == []
            o_tr.update(exp_f["fake_probability"])
        assert res["frame_count"] == count
        if i > 0:
            assert res["confidence_level"] == o_tr.get_confidence_level()
        assert res["temporal_average"] == float(o_tr.get_temporal_average())
        assert abs(res["stability_score"] - float(o_tr.get_stability_score())) < 1e-12
        assert trig is False
    small = frames[0][100:160, 100:170]                                           # < 80 px: +0.10 heuristic
    p_small, _, _ = det.analyze_face(small)
    ref_small = float(torch.sigmoid(oeff.forward(ofp.prepare(frames[0], (100, 100, 70, 60)), weights)).item())
    assert abs(p_small - float(np.clip(ref_small + 0.10, 0, 1))) <= 1e-4
    assert det.analyze_face(np.zeros((0, 0, 3), np.uint8)) == (None, None, None)
    det.reset()
    assert det.frame_count == 0 and det.temporal_tracker.get_confidence_level() == "UNCERTAIN"
    assert len(det.temporal_tracker.score_history) == 0 and det.frame_analyzer.frame_count == 0
    det.release()


# ---- /analyze contract (backend_server.py:82-255; reference tests/test_functional.py:356-423) ---------------
def test_http_contract(weights):
    from dfd_b200 import backend_server as bs
    client = bs.app.test_client()
    r = client.get("/health")
    h = r.get_json()
    assert r.status_code == 200 and h["status"] == "healthy" and "capabilities" in h and h["model_loaded"] is True
    assert client.post("/reset").get_json()["success"] is True
    time.sleep(0.15)
    assert client.post("/analyze", data={}).status_code == 400
    time.sleep(0.15)
    bad = client.post("/analyze", data={"frame": (io.BytesIO(b"not an image"), "x.jpg")}, content_type="multipart/form-data")
    assert bad.status_code == 400 and bad.get_json()["error"] == "Invalid image format"
    rng = np.random.RandomState(4)
    frame = synth.make_frame("pink", 480, 640, rng)
    for ext in (".jpg", ".png", ".bmp"):
        time.sleep(0.15)
        ok, enc = cv2.imencode(ext, frame)
        r = client.post("/analyze", data={"frame": (io.BytesIO(enc.tobytes()), "f" + ext)}, content_type="multipart/form-data")
        j = r.get_json()
        assert r.status_code == 200 and j["success"] and j["analysis_mode"] == "frame_only"
        for k in ("fake_probability", "frame_forensic_probability", "real_probability", "confidence_level",
                  "temporal_average", "stability_score", "frame_count", "processing_time_ms", "faces_detected"):
            assert k in j
        assert "face_probability" not in j
    time.sleep(0.15)
    ok, enc = cv2.imencode(".png", frame)
    r = client.post("/analyze", data={"frame": (io.BytesIO(enc.tobytes()), "f.png"), "faces": json.dumps([[200, 100, 220, 240]])},
                    content_type="multipart/form-data")
    j = r.get_json()
    assert r.status_code == 200 and j["analysis_mode"] == "face+frame" and j["face_bbox"] == {"x": 200, "y": 100, "width": 220, "height": 240}
    p = float(torch.sigmoid(oeff.forward(ofp.prepare(frame, (200, 100, 220, 240)), weights)).item())
    assert abs(j["face_probability"] - p) <= 1e-4 and j["fake_probability"] == j["face_probability"]
    r2 = client.post("/analyze", data={"frame": (io.BytesIO(enc.tobytes()), "f.png")}, content_type="multipart/form-data")
    assert r2.status_code == 429 and "retry_after_ms" in r2.get_json()
    s = client.get("/stats").get_json()
    assert s["frame_count"] == 4 and s["history_length"] == 4 and set(s["voting"]) == {"fake_count", "real_count", "total_frames"}
    client.post("/reset")
    assert client.get("/stats").get_json()["frame_count"] == 0
    assert client.get("/nope").status_code == 404


def test_http_jpeg_upload_is_decoded_on_the_device(weights):
    """/analyze with the extension's wire format (JPEG quality 85, 720 x 405): the upload is decoded on the device and the
    response equals what the reference computes from cv2.imdecode of the same bytes (forensic probability exactly, face
    probability within 1e-4); a progressive JPEG takes the reference's host ingest and gives the same contract."""
    from dfd_b200 import backend_server as bs
    from oracle import forensics as ofor
    client = bs.app.test_client()
    client.post("/reset")
    rng = np.random.RandomState(14)
    frame = synth.make_frame("natural", 405, 720, rng)
    ok, enc = cv2.imencode(".jpg", frame, [cv2.IMWRITE_JPEG_QUALITY, 85])
    decoded = cv2.imdecode(enc, cv2.IMREAD_COLOR)
    box = [150, 60, 260, 300]
    time.sleep(0.15)
    r = client.post("/analyze", data={"frame": (io.BytesIO(enc.tobytes()), "f.jpg"), "faces": json.dumps([box])},
                    content_type="multipart/form-data")
    j = r.get_json()
    assert r.status_code == 200 and j["analysis_mode"] == "face+frame"
    exp = ofor.OracleForensicAnalyzer().analyze(decoded)
    assert j["frame_forensic_probability"] == exp["fake_probability"]
    p = float(torch.sigmoid(oeff.forward(ofp.prepare(decoded, box), weights)).item())
    assert abs(j["face_probability"] - p) <= 1e-4
    # the decode itself: bit-exact with cv2
    det = bs._get_detector()
    dev = det.decode_frame(enc.tobytes())
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), decoded)
    ok, prog = cv2.imencode(".jpg", frame, [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    time.sleep(0.15)
    r = client.post("/analyze", data={"frame": (io.BytesIO(prog.tobytes()), "p.jpg")}, content_type="multipart/form-data")
    assert r.status_code == 200 and r.get_json()["analysis_mode"] == "frame_only"
    client.post("/reset")
