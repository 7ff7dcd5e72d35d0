"""Drop-in for the reference's ``deepfake_detection`` module (deepfake_detection.py:93-747):
``TemporalTracker`` and ``DeepfakeDetector`` with the reference's constructor arguments, attributes and
result shapes.  All numerics run in libdfd's CUDA kernels; these classes move one frame at a time across
the C-ABI (the batched multi-stream entry is ``dfd_b200.engine.Engine.analyze_batch``).

Test-time augmentation (``use_tta``, deepfake_detection.py:408-443) runs on the device: the random draws come from Python's
global ``random`` in the reference's order (``dfd_b200.tta``), flip / brightness / rotation / resize / classifier / mean are
libdfd kernels.  A calibrator (``weights/calibrator.pkl`` or ``set_calibrator``) that is a scikit-learn
``LogisticRegression`` is applied on the device as well; any other object with ``predict_proba`` is called on the host
exactly as the reference does (it is opaque user code).

Out of scope: face detection (boxes are inputs -- pass ``faces=[(x,y,w,h),...]`` or
install ``detector.face_detector``), MTCNN re-detection, GradCAM.
"""
import os
import pickle
import time
import warnings
from collections import deque

import numpy as np
import torch

from . import _lib, overlay as _overlay, runtime, tta as _tta
from .frame_analysis import FrameForensicAnalyzer

_model_state = {"sd": None, "loaded": False}
DEVICE = "cuda:0" if torch.cuda.is_available() else "cpu"      # deepfake_detection.py:21 (module attribute, a string)


def load_model_weights(state_dict_or_path):
    """Install the classifier weights every DeepfakeDetector of this process uses (the reference loads
    weights/best_model.pth into a module-global model at import time, deepfake_detection.py:30-90)."""
    from . import weights
    _model_state["sd"] = weights.extract_state_dict(state_dict_or_path)
    _model_state["loaded"] = True
    for eng in list(runtime._engines.values()):
        eng.load_state_dict(_model_state["sd"])


def _ensure_weights(eng):
    if eng.has_weights:
        return
    if _model_state["sd"] is None:
        from . import synth
        warnings.warn("no checkpoint installed (weights/best_model.pth is not shipped): using the fixed-seed "
                      "random-init state_dict; call load_model_weights(path)", stacklevel=3)
        _model_state["sd"] = synth.make_state_dict()
    eng.load_state_dict(_model_state["sd"])


class TemporalTracker:
    """Voting-based temporal classification (deepfake_detection.py:93-289) backed by a device ring buffer."""

    def __init__(self, window_size=60, high_confidence_threshold=0.6, voting_window=10, detection_threshold=0.5,
                 *, device=None, _engine=None, _slot=None):
        self.window_size = window_size
        self.high_confidence_threshold = high_confidence_threshold
        self.voting_window = voting_window
        self.detection_threshold = detection_threshold
        self._eng = _engine if _engine is not None else runtime.get_engine(device)
        self._own_slot = _slot is None
        self._slot = runtime.alloc_slot(self._eng) if _slot is None else _slot
        self._eng.configure_stream(self._slot, window_size, voting_window, detection_threshold)
        self.score_history = deque(maxlen=window_size)
        self.variance_history = deque(maxlen=30)
        self.frame_classifications = deque(maxlen=voting_window)
        self.current_verdict = None
        self.last_alert_time = 0
        self.alert_cooldown = 5
        self._avg, self._stab = 0.0, 0.0
        self._counts = (0, 0)

    # records produced by the device are mirrored into the public deques the reference exposes
    def _absorb(self, rec, p):
        if int(rec["last_vote"]) >= 0:
            self.score_history.append(p)
            if len(self.score_history) >= 5:
                self.variance_history.append(np.var(list(self.score_history)[-5:]))
            self.frame_classifications.append("FAKE" if int(rec["last_vote"]) == 1 else "REAL")
        v = int(rec["verdict"])
        self.current_verdict = None if v == _lib.UNCERTAIN else _lib.VERDICT_NAMES[v]
        self._avg, self._stab = float(rec["temporal_average"]), float(rec["stability_score"])
        self._counts = (int(rec["fake_count"]), int(rec["real_count"]))

    def update(self, fake_probability):
        if fake_probability is None:
            return
        is_np = isinstance(fake_probability, np.generic)        # selects how Python's sum() would round (vote.cu)
        rec = self._eng.records_to_numpy(self._eng.vote_update([self._slot], [float(fake_probability)], [int(is_np)]))[0]
        self._absorb(rec, fake_probability)

    def get_temporal_average(self):
        return self._avg if len(self.score_history) else 0.0

    def get_weighted_average(self):
        if len(self.score_history) == 0:
            return 0.0
        scores = list(self.score_history)
        w = np.linspace(0.5, 1.0, len(scores))
        return sum(s * x for s, x in zip(scores, w)) / sum(w)

    def get_stability_score(self):
        return self._stab if len(self.score_history) >= 10 else 0.0

    def detect_anomalies(self):
        if len(self.variance_history) < 10:
            return 0.0
        return min(np.mean(list(self.variance_history)) * 10, 1.0)

    def should_trigger_forensic_analysis(self):
        if len(self.score_history) < self.window_size // 2:
            return False
        now = time.time()
        if (self.get_temporal_average() > self.high_confidence_threshold and self.get_stability_score() > 0.7
                and now - self.last_alert_time > self.alert_cooldown):
            self.last_alert_time = now
            return True
        return False

    def get_confidence_level(self):
        return "UNCERTAIN" if self.current_verdict is None else self.current_verdict

    def get_voting_stats(self):
        return {"fake_count": self._counts[0], "real_count": self._counts[1],
                "total_frames": len(self.frame_classifications)}

    def reset(self):
        self._eng.reset_tracker(self._slot)
        self.score_history.clear()
        self.variance_history.clear()
        self.frame_classifications.clear()
        self.current_verdict = None
        self.last_alert_time = 0
        self._avg, self._stab, self._counts = 0.0, 0.0, (0, 0)

    def release(self):
        if self._own_slot:
            runtime.free_slot(self._eng, self._slot)
            self._slot = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class DeepfakeDetector:
    """Multi-signal detector (deepfake_detection.py:292-726): frame forensics always, face model when a box
    is supplied, 10-frame vote.  ``dtype`` selects the classifier precision ("fp32" accuracy mode or "bf16")."""

    def __init__(self, enable_gradcam=False, use_tta=True, num_tta_augmentations=3, detection_threshold=0.5,
                 face_weight=0.70, forensic_weight=0.30, *, device=None, dtype="fp32", face_detector=None):
        self.enable_gradcam = enable_gradcam
        self.use_tta = use_tta
        self.num_tta_augmentations = num_tta_augmentations
        self.detection_threshold = detection_threshold
        self.face_weight = face_weight          # stored and unused, exactly like the reference
        self.forensic_weight = forensic_weight
        self.dtype = dtype
        self.face_detector = face_detector
        self._eng = runtime.get_engine(device)
        self._slot = runtime.alloc_slot(self._eng)
        self._eng.reset(self._slot)
        self.temporal_tracker = TemporalTracker(window_size=60, high_confidence_threshold=0.6, voting_window=10,
                                                detection_threshold=detection_threshold, _engine=self._eng, _slot=self._slot)
        self.frame_count = 0
        self.frame_analyzer = FrameForensicAnalyzer(analysis_size=(256, 256), _engine=self._eng, _slot=self._slot)
        self.full_forensic_interval = 3
        self.last_frame_forensic_result = None
        self.calibrator = None
        self._device_calibrator = False
        calibrator_path = os.path.join(os.path.dirname(__file__), "weights", "calibrator.pkl")     # deepfake_detection.py:333-342
        if os.path.exists(calibrator_path):
            try:
                with open(calibrator_path, "rb") as f:
                    self.set_calibrator(pickle.load(f))
                print("✓ Probability calibrator loaded")
            except Exception:
                print("⚠️ Could not load calibrator")

    def reset(self):
        self.temporal_tracker.reset()
        self.frame_count = 0
        self.frame_analyzer.reset()
        self.last_frame_forensic_result = None

    def release(self):
        runtime.free_slot(self._eng, self._slot)
        self._slot = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    # -- Layer 0: frame forensics -----------------------------------------------------
    def analyze_frame_forensics(self, frame):
        """Full analysis every 3rd frame_count, fast otherwise (deepfake_detection.py:504-515)."""
        if self.frame_count % self.full_forensic_interval == 0:
            result = self.frame_analyzer.analyze(frame)
        else:
            result = self.frame_analyzer.analyze_fast(frame)
        self.last_frame_forensic_result = result
        return result

    # -- Layer 1: face model ----------------------------------------------------------------
    def set_calibrator(self, calibrator):
        """Install ``self.calibrator`` (deepfake_detection.py:333-342).  A fitted scikit-learn ``LogisticRegression`` on the
        single raw-probability feature (Platt scaling) becomes a device-side map (``dfd_set_calibrator``); ``None`` removes it;
        any other object is kept and called on the host through ``predict_proba`` as the reference does."""
        self.calibrator = calibrator
        self._device_calibrator = False
        self._eng.set_calibrator("none")
        if calibrator is None:
            return
        coef, icpt = getattr(calibrator, "coef_", None), getattr(calibrator, "intercept_", None)
        if (type(calibrator).__name__ == "LogisticRegression" and coef is not None and np.shape(coef) == (1, 1)
                and np.shape(icpt) == (1,)):
            self._eng.set_calibrator("logistic", [float(coef[0][0])], [float(icpt[0])])
            self._device_calibrator = True

    def apply_calibration(self, raw_prob):
        """deepfake_detection.py:445-455 (host form, for callers that use it directly)."""
        if self.calibrator is None:
            return raw_prob
        try:
            return self.calibrator.predict_proba([[raw_prob]])[0][1]
        except Exception:
            return raw_prob

    def apply_heuristics(self, fake_prob, face_region):
        h, w = face_region.shape[:2]
        return np.clip(fake_prob + (0.10 if (h < 80 or w < 80) else 0.0), 0, 1)

    def _face_probability(self, frames, box, cw, ch):
        """prep (+ TTA) -> classifier -> sigmoid / mean / calibration / heuristics for ONE box of frames[0]; cw x ch is the
        crop numpy slicing would give.  np.float64 like the reference's np.clip result, NaN if the box was rejected."""
        n_pred = self.num_tta_augmentations if self.use_tta else 1
        bx = np.asarray([box], np.int32)
        if n_pred > 1:
            params = [_tta.draw_params(n_pred)]                    # consumes `random` exactly like deepfake_detection.py:418-430
            xin = self._eng.face_prep_tta(frames, bx, [0], params, self.dtype)
        else:
            xin = self._eng.face_prep_batch(frames, bx, [0], self.dtype)
        logits = self._eng.effnet_forward(xin)
        host_calib = self.calibrator is not None and not self._device_calibrator
        # an opaque host calibrator sits between the mean and the heuristics: take the raw mean from the device (a >= 80 px
        # dummy size switches the heuristic off), call the object, apply the reference's heuristic line on the host
        size = np.array([[0, 0, 80, 80]] if host_calib else [[0, 0, cw, ch]], np.int32)
        p = self._eng.face_probability_tta(logits, size, n_pred) if n_pred > 1 else self._eng.face_probability(logits, size)
        p = np.float64(p.cpu().numpy()[0])
        if host_calib and not np.isnan(p):
            p = np.clip(self.apply_calibration(float(p)) + (0.10 if (ch < 80 or cw < 80) else 0.0), 0, 1)
        return p

    def analyze_face(self, face_region):
        """(p, p, None) for a BGR face crop, or (None, None, None) on failure (deepfake_detection.py:517-550)."""
        try:
            face = np.asarray(face_region)
            if face.ndim != 3 or face.shape[2] != 3 or face.dtype != np.uint8 or face.shape[0] < 1 or face.shape[1] < 1:
                raise ValueError("face_region must be a non-empty (h, w, 3) uint8 BGR crop")
            _ensure_weights(self._eng)
            h, w = face.shape[:2]
            ft = torch.from_numpy(np.ascontiguousarray(face)).to(self._eng.device).unsqueeze(0)
            p = self._face_probability(ft, [0, 0, w, h], w, h)
            if np.isnan(p):                                 # box rejected on the device (k_box_sanitize)
                raise ValueError(f"face crop {w}x{h} exceeds the engine's max_crop ({self._eng.cfg.max_crop})")
            return p, p, None
        except Exception as e:                              # the reference swallows and reports (:548-550)
            print(f"Face analysis error: {e}")
            return None, None, None

    def analyze_face_box(self, frame_dev, box):
        """``analyze_face(frame[y:y+h, x:x+w])`` for a frame that is already on the device (an (H, W, 3) uint8 CUDA tensor,
        e.g. from ``decode_frame``): the crop never visits the host.  Same return convention as ``analyze_face``."""
        try:
            _ensure_weights(self._eng)
            x, y, w, h = (int(v) for v in box)
            H, W = int(frame_dev.shape[0]), int(frame_dev.shape[1])
            # heuristics / rotation centre see the crop numpy slicing would give (clamped to the frame), like face_region.shape
            cw, ch = max(min(x + w, W) - max(x, 0), 0), max(min(y + h, H) - max(y, 0), 0)
            p = self._face_probability(frame_dev.contiguous().unsqueeze(0), [x, y, w, h], cw, ch)
            if np.isnan(p):
                raise ValueError(f"face box {box} is empty inside the {W}x{H} frame or exceeds the engine's max_crop")
            return p, p, None
        except Exception as e:
            print(f"Face analysis error: {e}")
            return None, None, None

    def decode_frame(self, image_bytes):
        """``cv2.imdecode(np.frombuffer(image_bytes, np.uint8), cv2.IMREAD_COLOR)`` (backend_server.py:140-142) on the
        DEVICE for baseline JPEG uploads (the extension's wire format): returns an (H, W, 3) uint8 BGR CUDA tensor, bit-exact
        with OpenCV.  Raises DfdError for streams the device decoder does not cover (the caller decides what to do)."""
        H, W = self._eng.jpeg_info(image_bytes)[:2]
        packed, offsets = self._eng.pack_jpegs([image_bytes])
        frames, status = self._eng.decode_jpeg_batch(packed, offsets, H, W)
        if int(status.cpu()[0]) != 0:
            raise _lib.DfdError("corrupt JPEG stream (entropy-coded data incomplete)")
        return frames[0]

    # -- result annotation (deepfake_detection.py:552-586, 688-726) ----------------------------------
    def get_box_color(self, confidence_level):
        return _overlay.get_box_color(confidence_level)

    def _as_device_frame(self, frame):
        if torch.is_tensor(frame) and frame.is_cuda:
            return frame if frame.is_contiguous() else frame.contiguous()
        return torch.from_numpy(np.ascontiguousarray(frame)).to(self._eng.device)

    def _finish_drawing(self, frame, fdev):
        """The reference draws into the caller's array and returns it: copy the annotated device frame back into it."""
        if torch.is_tensor(frame) and frame.is_cuda:
            if fdev is not frame:
                frame.copy_(fdev)
            return frame
        out = fdev.cpu().numpy()
        if isinstance(frame, np.ndarray) and frame.flags.writeable and frame.shape == out.shape:
            frame[...] = out
            return frame
        return out

    def draw_detection_overlay(self, frame, x, y, w, h, fake_prob, confidence_level):
        """Box, verdict label and vote counts drawn on the device copy of the frame (bit-exact with the reference's OpenCV
        calls); accepts a numpy frame (annotated in place and returned, like the reference) or a CUDA tensor."""
        fdev = self._as_device_frame(frame)
        cl = _overlay.CommandList(fdev.shape[0], fdev.shape[1])
        _overlay.detection_overlay(cl, int(x), int(y), int(w), int(h), fake_prob, confidence_level,
                                   self.temporal_tracker.get_voting_stats())
        self._eng.draw_overlay(fdev, cl)
        return self._finish_drawing(frame, fdev)

    def _draw_frame_analysis_overlay(self, frame, fake_prob, confidence_level, forensic_result):
        fdev = self._as_device_frame(frame)
        cl = _overlay.CommandList(fdev.shape[0], fdev.shape[1])
        _overlay.frame_analysis_overlay(cl, fake_prob, confidence_level, forensic_result)
        self._eng.draw_overlay(fdev, cl)
        return self._finish_drawing(frame, fdev)

    # -- whole frame ---------------------------------------------------------------------------
    def predict(self, frame, faces=None, draw=True):
        """(frame, trigger_forensic, forensic_frame, result_data) as deepfake_detection.py:588-686.  ``faces`` =
        [(x, y, w, h), ...] (face detection is the caller's).  The frame is uploaded once; forensics, every face crop and the
        annotation (``draw=True``, the reference's behaviour: the overlay of face k is part of the pixels face k + 1 is cropped
        from, :611-634) work on the device copy, and the annotated frame is written back into the caller's array."""
        self.frame_count += 1
        is_np = not (torch.is_tensor(frame) and frame.is_cuda)
        if is_np:
            f = np.asarray(frame)
            if f.ndim != 3 or f.shape[2] != 3 or f.dtype != np.uint8:
                raise ValueError("frame must be an (H, W, 3) uint8 BGR image")
        fdev = self._as_device_frame(frame)
        H, W = int(fdev.shape[0]), int(fdev.shape[1])
        frame_forensic = self.analyze_frame_forensics(fdev)
        if faces is None:
            faces = self.face_detector(frame if is_np else fdev.cpu().numpy()) if self.face_detector is not None else []
        trigger, forensic_frame, face_results = False, None, []
        confidence_level = "UNCERTAIN"
        drawn = False
        if len(faces) > 0:
            for (x, y, w, h) in faces:
                fake_prob, _, _ = self.analyze_face_box(fdev, (x, y, w, h))
                if fake_prob is None:
                    continue
                self.temporal_tracker.update(fake_prob)
                confidence_level = self.temporal_tracker.get_confidence_level()
                if self.temporal_tracker.should_trigger_forensic_analysis():
                    trigger, forensic_frame = True, fdev.cpu().numpy()
                if draw:
                    cl = _overlay.CommandList(H, W)
                    _overlay.detection_overlay(cl, int(x), int(y), int(w), int(h), fake_prob, confidence_level,
                                               self.temporal_tracker.get_voting_stats())
                    self._eng.draw_overlay(fdev, cl)
                    drawn = True
                face_results.append({"face_prob": float(fake_prob), "combined_prob": float(fake_prob),
                                     "bbox": {"x": int(x), "y": int(y), "w": int(w), "h": int(h)}})
            if not face_results:
                # the reference raises NameError here (confidence_level unbound); report the tracker state
                confidence_level = self.temporal_tracker.get_confidence_level()
        else:
            frame_fake_prob = frame_forensic["fake_probability"]
            self.temporal_tracker.update(frame_fake_prob)
            confidence_level = self.temporal_tracker.get_confidence_level()
            if self.temporal_tracker.should_trigger_forensic_analysis():
                trigger, forensic_frame = True, fdev.cpu().numpy()
            if draw:
                cl = _overlay.CommandList(H, W)
                _overlay.frame_analysis_overlay(cl, frame_fake_prob, confidence_level, frame_forensic)
                self._eng.draw_overlay(fdev, cl)
                drawn = True
        if drawn:
            frame = self._finish_drawing(frame, fdev)
        result_data = {
            "frame_count": self.frame_count,
            "faces_detected": len(faces),
            "face_results": face_results,
            "frame_forensic": frame_forensic,
            "confidence_level": confidence_level if len(faces) or self.frame_count > 1 else "UNCERTAIN",
            "temporal_average": float(self.temporal_tracker.get_temporal_average()),
            "stability_score": float(self.temporal_tracker.get_stability_score()),
            "analysis_mode": "face+frame" if len(faces) > 0 else "frame_only",
        }
        return frame, trigger, forensic_frame, result_data

    analyze = predict      # alias named by the project README / north_star; not present in the reference class


_detector = None


def _global_detector():
    global _detector
    if _detector is None:
        _detector = DeepfakeDetector(use_tta=False, num_tta_augmentations=1, detection_threshold=0.5,
                                     face_weight=0.70, forensic_weight=0.30)      # deepfake_detection.py:730-736
    return _detector


def predict(frame, faces=None):
    return _global_detector().predict(frame, faces)[0]


def predict_with_forensics(frame, faces=None):
    return _global_detector().predict(frame, faces)
