"""Drop-in for the reference's Flask backend (backend_server.py:45-275): same routes, status codes and JSON
fields, served by a small werkzeug WSGI app (Flask is not in this image; ``app.test_client()`` and
``app.run()`` behave like Flask's for the reference's tests).

Face boxes are inputs on this path: ``POST /analyze`` accepts an optional multipart field ``faces`` holding a
JSON list ``[[x, y, w, h], ...]``; without it (and without ``face_detector`` installed) the request is
analysed in ``frame_only`` mode.

Ingest (reference backend_server.py:140-142, ``cv2.imdecode``): a baseline JPEG upload -- the extension's wire format,
``canvas.toDataURL('image/jpeg', 0.85)`` -- is decoded ON THE DEVICE (``DeepfakeDetector.decode_frame`` ->
``dfd_decode_jpeg_batch``, bit-exact with OpenCV), so the frame never exists in host memory and the face crop is taken
from the device copy.  Uploads in any other format (PNG, progressive JPEG, ...) keep the reference's own ingest call,
``cv2.imdecode`` on the host: that is the caller's side of the boundary, not a fallback of the compute path -- every
signal, the classifier and the vote still run in libdfd.
"""
import json
import logging
import time

import numpy as np
import torch
from werkzeug.test import Client
from werkzeug.wrappers import Request, Response

from . import _lib
from .deepfake_detection import DeepfakeDetector, _model_state

logger = logging.getLogger(__name__)

detector = None
face_detector = None            # optional callable(frame) -> [(x, y, w, h)]
_last_request_time = 0
_min_request_interval = 0.1     # backend_server.py:62-63


def _get_detector():
    global detector
    if detector is None:
        detector = DeepfakeDetector(enable_gradcam=False, use_tta=False, num_tta_augmentations=1,
                                    detection_threshold=0.55)                     # backend_server.py:57
    return detector


def _json(obj, status=200):
    r = Response(json.dumps(obj), status=status, mimetype="application/json")
    r.headers["Access-Control-Allow-Origin"] = "*"
    r.headers["Access-Control-Allow-Methods"] = "GET, POST, OPTIONS"
    r.headers["Access-Control-Allow-Headers"] = "Content-Type"
    return r


def health_check(request):
    det = _get_detector()
    gpu = torch.cuda.is_available()
    return _json({"status": "healthy", "model_loaded": True, "device": "cuda:0" if gpu else "cpu",
                  "gpu_name": torch.cuda.get_device_name(0) if gpu else None, "frame_count": det.frame_count,
                  "weights_loaded": bool(_model_state["loaded"]),
                  "capabilities": {"face_detection": face_detector is not None, "frame_forensics": True,
                                   "temporal_tracking": True}})


def reset_detector(request):
    try:
        _get_detector().reset()
        return _json({"success": True, "message": "Detector reset successfully"})
    except Exception as e:
        return _json({"success": False, "error": str(e)}, 500)


def analyze_frame(request):
    global _last_request_time
    now = time.time()
    elapsed = now - _last_request_time
    if elapsed < _min_request_interval:                                         # backend_server.py:66-80
        return _json({"error": "Rate limited", "retry_after_ms": int((_min_request_interval - elapsed) * 1000)}, 429)
    _last_request_time = now
    start = time.time()
    try:
        import cv2
        if "frame" not in request.files:
            return _json({"error": "No frame provided"}, 400)
        raw = request.files["frame"].read()
        det = _get_detector()
        frame = None
        if len(raw) > 3 and raw[:2] == b"\xff\xd8":                              # JPEG: decode on the device
            try:
                frame = det.decode_frame(raw)
            except _lib.DfdError:
                frame = None                                                     # not a baseline stream: reference ingest below
        if frame is None:
            data = np.frombuffer(raw, np.uint8)
            frame = cv2.imdecode(data, cv2.IMREAD_COLOR) if data.size else None
        if frame is None:
            return _json({"error": "Invalid image format"}, 400)
        ff = det.analyze_frame_forensics(frame)                                  # frame_count read BEFORE increment (:148,156)
        ff_prob = ff["fake_probability"]
        if "faces" in request.form:
            faces = [tuple(int(v) for v in b) for b in json.loads(request.form["faces"])]
        else:
            faces = (face_detector(frame.cpu().numpy() if torch.is_tensor(frame) else frame)
                     if face_detector is not None else [])
        det.frame_count += 1
        tr = det.temporal_tracker
        if len(faces) > 0:
            x, y, w, h = faces[0]
            if torch.is_tensor(frame):
                fake_prob, _, _ = det.analyze_face_box(frame, (x, y, w, h))
            else:
                fake_prob, _, _ = det.analyze_face(frame[y:y + h, x:x + w])
            if fake_prob is not None:
                tr.update(fake_prob)
                ms = (time.time() - start) * 1000
                return _json({
                    "success": True, "analysis_mode": "face+frame", "faces_detected": len(faces),
                    "fake_probability": float(fake_prob), "face_probability": float(fake_prob),
                    "frame_forensic_probability": float(ff_prob), "real_probability": float(1 - fake_prob),
                    "confidence_level": tr.get_confidence_level(), "temporal_average": float(tr.get_temporal_average()),
                    "stability_score": float(tr.get_stability_score()), "frame_count": det.frame_count,
                    "processing_time_ms": round(ms, 1),
                    "face_bbox": {"x": int(x), "y": int(y), "width": int(w), "height": int(h)}})
        tr.update(ff_prob)
        ms = (time.time() - start) * 1000
        return _json({
            "success": True, "analysis_mode": "frame_only", "faces_detected": len(faces),
            "fake_probability": float(ff_prob), "frame_forensic_probability": float(ff_prob),
            "real_probability": float(1 - ff_prob), "confidence_level": tr.get_confidence_level(),
            "temporal_average": float(tr.get_temporal_average()), "stability_score": float(tr.get_stability_score()),
            "frame_count": det.frame_count, "processing_time_ms": round(ms, 1)})
    except Exception as e:
        logger.error("Error analyzing frame: %s", e)
        return _json({"error": str(e)}, 500)


def get_stats(request):
    try:
        det = _get_detector()
        tr = det.temporal_tracker
        return _json({"frame_count": det.frame_count, "temporal_average": float(tr.get_temporal_average()),
                      "stability_score": float(tr.get_stability_score()), "confidence_level": tr.get_confidence_level(),
                      "history_length": len(tr.score_history), "voting": tr.get_voting_stats(),
                      "device": "cuda:0" if torch.cuda.is_available() else "cpu"})
    except Exception as e:
        return _json({"error": str(e)}, 500)


ROUTES = {("GET", "/health"): health_check, ("POST", "/reset"): reset_detector,
          ("POST", "/analyze"): analyze_frame, ("GET", "/stats"): get_stats}


class _TestClient(Client):
    """werkzeug's test client with Flask's context-manager form (``with app.test_client() as c:``); responses are werkzeug
    TestResponse objects (``status_code``, ``data``, ``get_json()``, ``json``) like Flask's."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class App:
    """Minimal WSGI application with the slice of Flask's surface the reference's tests use."""

    def __init__(self):
        self.config = {}            # Flask's app.config (the reference's tests set TESTING)

    def __call__(self, environ, start_response):
        request = Request(environ)
        if request.method == "OPTIONS":
            resp = _json({}, 200)
        else:
            fn = ROUTES.get((request.method, request.path))
            if fn is None:
                known = any(path == request.path for (_, path) in ROUTES)
                resp = _json({"error": "Method Not Allowed" if known else "Not Found"}, 405 if known else 404)
            else:
                resp = fn(request)
        return resp(environ, start_response)

    def test_client(self):
        return _TestClient(self)

    def run(self, host="0.0.0.0", port=5000, debug=False, threaded=True):
        from werkzeug.serving import run_simple
        run_simple(host, port, self, threaded=False)      # one context, one host thread (include/dfd.h)


app = App()

if __name__ == "__main__":
    app.run()
