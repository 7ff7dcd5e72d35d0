"""Drop-in for the reference's ``frame_analysis.FrameForensicAnalyzer`` (frame_analysis.py:22-395):
same constructor, attributes and result dicts; the six signals run in libdfd's CUDA kernels."""
from collections import deque

import numpy as np
import torch

from . import runtime

SIGNALS = ("frequency", "noise", "ela", "edge", "color", "temporal")


class FrameForensicAnalyzer:
    """Analyzes video frames for deepfake artifacts (works with or without faces).

    Differences from the reference object, all outside its documented surface: ``prev_frame_gray`` and
    ``temporal_diffs`` live on the device (``prev_frame_gray`` is ``None`` before the first frame and a
    placeholder marker afterwards); ``analysis_size`` must be (256, 256), the only size the reference uses.
    """

    def __init__(self, analysis_size=(256, 256), *, device=None, _engine=None, _slot=None):
        if tuple(analysis_size) != (256, 256):
            raise ValueError("the B200 path implements the reference's analysis_size=(256, 256) only")
        self.analysis_size = analysis_size
        self._eng = _engine if _engine is not None else runtime.get_engine(device)
        self._own_slot = _slot is None
        self._slot = runtime.alloc_slot(self._eng) if _slot is None else _slot
        self.prev_frame_gray = None
        self.temporal_diffs = deque(maxlen=30)
        self.frame_count = 0
        self.weights = {"frequency": 0.25, "noise": 0.20, "ela": 0.20, "edge": 0.15, "color": 0.10, "temporal": 0.10}
        self.last_raw = None
        if self._own_slot:
            self._eng.reset(self._slot)

    def _run(self, frame, full):
        if torch.is_tensor(frame) and frame.is_cuda:      # a frame already resident on the device (device JPEG ingest)
            if frame.dim() != 3 or frame.shape[2] != 3 or frame.dtype != torch.uint8:
                raise ValueError("frame must be an (H, W, 3) uint8 BGR image")
            ft = frame.contiguous().unsqueeze(0)
        else:
            frame = np.asarray(frame)
            if frame.ndim != 3 or frame.shape[2] != 3 or frame.dtype != np.uint8:
                raise ValueError("frame must be an (H, W, 3) uint8 BGR image")   # cv2.resize would raise too
            ft = torch.from_numpy(np.ascontiguousarray(frame)).to(self._eng.device).unsqueeze(0)
        res = self._eng.forensic_to_numpy(self._eng.forensics_batch(ft, [self._slot], [1 if full else 0]))[0]
        self.frame_count = int(res["frame_number"])
        self.prev_frame_gray = "device"
        raw = np.array(res["raw"])
        self.last_raw = raw
        if not np.isnan(raw[13]):
            self.temporal_diffs.append(np.float32(raw[13]))
        order = SIGNALS if full else ("frequency", "temporal", "edge")
        scores = {k: float(res["scores"][SIGNALS.index(k)]) for k in order}
        return {"scores": scores, "fake_probability": float(res["fake_probability"]),
                "analysis_type": "frame_forensic" if full else "frame_forensic_fast", "frame_number": self.frame_count}

    def analyze(self, frame):
        """All six signals (frame_analysis.py:58-101)."""
        return self._run(frame, True)

    def analyze_fast(self, frame):
        """frequency / temporal / edge only (frame_analysis.py:103-126)."""
        return self._run(frame, False)

    def reset(self):
        """frame_analysis.py:391-395."""
        self._eng.lib.dfd_reset_stream  # noqa: B018  (symbol must exist)
        self._eng.reset_forensics(self._slot)
        self.prev_frame_gray = None
        self.temporal_diffs.clear()
        self.frame_count = 0

    def release(self):
        if self._own_slot:
            runtime.free_slot(self._eng, self._slot)
            self._slot = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
