"""Oracle (CPU, test infrastructure): the 10-frame temporal vote.

Restates ``TemporalTracker`` (reference deepfake_detection.py:93-289) without
its per-frame prints.  All arithmetic is Python float (IEEE double) in deque
insertion order, which is what the device-side ring-buffer kernel must match
bit for bit (verdicts, counts) -- SURVEY.md §3.4.
"""
from collections import deque

import numpy as np


class OracleTemporalTracker:
    def __init__(self, window_size=60, high_confidence_threshold=0.6, voting_window=10,
                 detection_threshold=0.5):
        self.window_size = window_size
        self.high_confidence_threshold = high_confidence_threshold
        self.voting_window = voting_window
        self.detection_threshold = detection_threshold
        self.score_history = deque(maxlen=window_size)          # :111
        self.variance_history = deque(maxlen=30)                # :112
        self.frame_classifications = deque(maxlen=voting_window)  # :117
        self.current_verdict = None

    def update(self, p):
        if p is None:                                            # :123-124
            return
        self.score_history.append(p)
        if len(self.score_history) >= 5:                         # :129-132
            self.variance_history.append(np.var(list(self.score_history)[-5:]))
        self.frame_classifications.append("FAKE" if p > self.detection_threshold else "REAL")  # :135 strict >
        n = len(self.frame_classifications)
        if n == 0 or n < self.voting_window:                     # :152-160
            self.current_verdict = None
            return
        fake = sum(1 for c in self.frame_classifications if c == "FAKE")
        self.current_verdict = "FAKE" if fake > n - fake else "REAL"   # :175-178 tie -> REAL

    def get_confidence_level(self):
        return "UNCERTAIN" if self.current_verdict is None else self.current_verdict   # :252-258

    def get_voting_stats(self):
        fake = sum(1 for c in self.frame_classifications if c == "FAKE")
        return {"fake_count": fake, "real_count": len(self.frame_classifications) - fake,
                "total_frames": len(self.frame_classifications)}

    def get_temporal_average(self):
        if not self.score_history:
            return 0.0
        return sum(self.score_history) / len(self.score_history)     # :198-202

    def get_stability_score(self):
        if len(self.score_history) < 10:                              # :216-221
            return 0.0
        s = list(self.score_history)
        mean = sum(s) / len(s)
        var = sum((x - mean) ** 2 for x in s) / len(s)
        return 1.0 - min(var * 4, 1.0)

    def reset(self):
        self.score_history.clear()
        self.variance_history.clear()
        self.frame_classifications.clear()
        self.current_verdict = None
