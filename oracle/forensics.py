"""Oracle (CPU, test infrastructure): the six frame-forensic signals.

Restates ``FrameForensicAnalyzer`` (reference frame_analysis.py:22-395) with
the same third-party calls (cv2 / numpy) but split into *raw statistics* and
*score tables*, so a CUDA kernel can be compared at both levels:

    raw statistics  -> relative tolerance 1e-4   (north_star)
    step scores     -> identical (they are sums of constants)

Raw statistic order (``RAW_NAMES``) is the layout of ``raw_stats_out`` in
``include/dfd.h``.
"""
from collections import deque

import cv2
import numpy as np

RAW_NAMES = (
    "freq_high_ratio", "freq_mid_ratio", "freq_mid_cv",      # frame_analysis.py:156-170
    "noise_cv", "noise_mean",                                # :208-209
    "ela_cv", "ela_mean",                                    # :259-260
    "edge_density", "lap_var",                               # :289-293
    "sat_std", "val_std", "unique_hues",                     # :321-341
    "temporal_cv", "temporal_last_diff", "temporal_n",       # :364-373
    "reserved",
)
RAW = {n: i for i, n in enumerate(RAW_NAMES)}
N_RAW = len(RAW_NAMES)

SIGNALS = ("frequency", "noise", "ela", "edge", "color", "temporal")
FULL_WEIGHTS = {"frequency": 0.25, "noise": 0.20, "ela": 0.20,
                "edge": 0.15, "color": 0.10, "temporal": 0.10}      # :49-56
FAST_ORDER = ("frequency", "temporal", "edge")                         # :114-116
FAST_WEIGHTS = {"frequency": 0.45, "temporal": 0.25, "edge": 0.30}    # :118


def _clip01(x):
    return float(np.clip(x, 0.0, 1.0))


# --------------------------------------------------------------------------
# raw statistics
# --------------------------------------------------------------------------
def radial_masks(h, w):
    """Band masks of frame_analysis.py:40-46,146-148."""
    cy, cx = h // 2, w // 2
    yy, xx = np.ogrid[:h, :w]
    dist = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2)
    r_in, r_mid, r_out = min(h, w) // 8, min(h, w) // 4, min(h, w) // 2
    low = dist <= r_in
    mid = (dist > r_in) & (dist <= r_mid)
    high = (dist > r_mid) & (dist <= r_out)
    return low, mid, high


def frequency_stats(tile_bgr, masks):
    """frame_analysis.py:136-170.  fft2 of a float32 array is complex64 under
    NumPy 2.x."""
    gray = cv2.cvtColor(tile_bgr, cv2.COLOR_BGR2GRAY).astype(np.float32)
    mag = np.log1p(np.abs(np.fft.fftshift(np.fft.fft2(gray))))
    low_m, mid_m, high_m = masks
    low = mag[low_m].mean() if np.any(low_m) else 0
    mid = mag[mid_m].mean() if np.any(mid_m) else 0
    high = mag[high_m].mean() if np.any(high_m) else 0
    total = low + mid + high + 1e-10
    mid_vals = mag[mid_m]
    mid_cv = (np.std(mid_vals) / (np.mean(mid_vals) + 1e-10)) if len(mid_vals) > 10 else None
    return high / total, mid / total, mid_cv


def _block_reduce(img, fn, block=32):
    h, w = img.shape
    out = []
    for i in range(0, h - block + 1, block):
        for j in range(0, w - block + 1, block):
            out.append(fn(img[i:i + block, j:j + block]))
    return out


def noise_stats(tile_bgr):
    """frame_analysis.py:188-209."""
    gray = cv2.cvtColor(tile_bgr, cv2.COLOR_BGR2GRAY).astype(np.float32)
    resid = gray - cv2.GaussianBlur(gray, (5, 5), 0)
    stds = _block_reduce(resid, np.std)
    if len(stds) < 4:
        return None
    stds = np.array(stds)
    mean_noise = np.mean(stds)
    return np.std(stds) / (mean_noise + 1e-10), mean_noise


def jpeg_q90_roundtrip(tile_bgr):
    """frame_analysis.py:234-236."""
    ok, enc = cv2.imencode(".jpg", tile_bgr, [int(cv2.IMWRITE_JPEG_QUALITY), 90])
    return cv2.imdecode(enc, cv2.IMREAD_COLOR)


def ela_stats(tile_bgr):
    """frame_analysis.py:234-260."""
    rec = jpeg_q90_roundtrip(tile_bgr)
    if rec is None:
        return None
    d = cv2.cvtColor(cv2.absdiff(tile_bgr, rec), cv2.COLOR_BGR2GRAY).astype(np.float32)
    means = _block_reduce(d, np.mean)
    if len(means) < 4:
        return None
    means = np.array(means)
    ela_mean = np.mean(means)
    return np.std(means) / (ela_mean + 1e-10), ela_mean


def edge_stats(tile_bgr):
    """frame_analysis.py:285-293."""
    gray = cv2.cvtColor(tile_bgr, cv2.COLOR_BGR2GRAY)
    edges = cv2.Canny(gray, 50, 150)
    density = np.sum(edges > 0) / edges.size
    lap_var = np.var(cv2.Laplacian(gray, cv2.CV_64F))
    return density, lap_var


def color_stats(tile_bgr):
    """frame_analysis.py:318-341."""
    hsv = cv2.cvtColor(tile_bgr, cv2.COLOR_BGR2HSV)
    sat_std = np.std(hsv[:, :, 1].astype(np.float32))
    val_std = np.std(hsv[:, :, 2].astype(np.float32))
    return sat_std, val_std, len(np.unique(hsv[:, :, 0]))


# --------------------------------------------------------------------------
# step-score tables
# --------------------------------------------------------------------------
def frequency_score(high_ratio, mid_ratio, mid_cv):
    """frame_analysis.py:159-180."""
    s = 0.0
    if high_ratio < 0.18:
        s += 0.4
    elif high_ratio < 0.22:
        s += 0.2
    if mid_cv is not None:
        if mid_cv > 0.6:
            s += 0.25
        elif mid_cv > 0.45:
            s += 0.1
    if mid_ratio > 0.45 and high_ratio < 0.2:
        s += 0.15
    return _clip01(s)


def noise_score(st):
    """frame_analysis.py:204-225."""
    if st is None:
        return 0.0
    cv, mean = st
    s = 0.0
    if cv > 0.7:
        s += 0.5
    elif cv > 0.5:
        s += 0.25
    if mean < 1.0:
        s += 0.3
    elif mean < 2.0:
        s += 0.1
    return _clip01(s)


def ela_score(st):
    """frame_analysis.py:255-276."""
    if st is None:
        return 0.0
    cv, mean = st
    s = 0.0
    if cv > 0.9:
        s += 0.5
    elif cv > 0.6:
        s += 0.2
    if mean > 15:
        s += 0.2
    elif mean > 10:
        s += 0.1
    return _clip01(s)


def edge_score(density, lap_var):
    """frame_analysis.py:295-309."""
    s = 0.0
    if density < 0.02:
        s += 0.35
    elif density < 0.04:
        s += 0.15
    if lap_var < 50:
        s += 0.3
    elif lap_var < 100:
        s += 0.1
    return _clip01(s)


def color_score(sat_std, val_std, hues):
    """frame_analysis.py:326-347."""
    s = 0.0
    if sat_std < 15:
        s += 0.3
    elif sat_std < 25:
        s += 0.1
    if val_std < 15:
        s += 0.25
    elif val_std < 25:
        s += 0.1
    if hues < 30:
        s += 0.25
    elif hues < 50:
        s += 0.1
    return _clip01(s)


class OracleForensicAnalyzer:
    """Stateful analyzer with the reference's surface (frame_analysis.py:22-56,
    58-126, 391-395) that additionally records ``last_raw`` (np.float64[N_RAW],
    NaN where a statistic was not computed)."""

    def __init__(self, analysis_size=(256, 256)):
        self.analysis_size = analysis_size
        self.prev_frame_gray = None
        self.temporal_diffs = deque(maxlen=30)
        self.frame_count = 0
        h, w = analysis_size
        self._masks = radial_masks(h, w)
        self.weights = dict(FULL_WEIGHTS)
        self.last_raw = np.full(N_RAW, np.nan)

    # -- temporal signal keeps state: frame_analysis.py:349-389 --------------
    def _temporal(self, tile):
        raw = self.last_raw
        gray = cv2.cvtColor(tile, cv2.COLOR_BGR2GRAY).astype(np.float32)
        if self.prev_frame_gray is None:
            self.prev_frame_gray = gray
            raw[RAW["temporal_n"]] = 0
            return 0.0
        mean_diff = np.mean(cv2.absdiff(gray, self.prev_frame_gray))
        self.temporal_diffs.append(mean_diff)
        self.prev_frame_gray = gray
        raw[RAW["temporal_last_diff"]] = mean_diff
        raw[RAW["temporal_n"]] = len(self.temporal_diffs)
        if len(self.temporal_diffs) < 5:
            return 0.0
        diffs = np.array(self.temporal_diffs)
        cv = np.std(diffs) / (np.mean(diffs) + 1e-10)
        raw[RAW["temporal_cv"]] = cv
        s = 0.0
        if cv > 1.5:
            s += 0.4
        elif cv > 1.0:
            s += 0.2
        if mean_diff < 0.3 and self.frame_count > 10:
            s += 0.3
        elif mean_diff < 0.8 and self.frame_count > 10:
            s += 0.1
        return _clip01(s)

    def _frequency(self, tile):
        hr, mr, cv = frequency_stats(tile, self._masks)
        self.last_raw[RAW["freq_high_ratio"]] = hr
        self.last_raw[RAW["freq_mid_ratio"]] = mr
        self.last_raw[RAW["freq_mid_cv"]] = np.nan if cv is None else cv
        return frequency_score(hr, mr, cv)

    def _edges(self, tile):
        d, lv = edge_stats(tile)
        self.last_raw[RAW["edge_density"]] = d
        self.last_raw[RAW["lap_var"]] = lv
        return edge_score(d, lv)

    def _resize(self, frame):
        return cv2.resize(frame, self.analysis_size, interpolation=cv2.INTER_LINEAR)  # :71

    def analyze(self, frame):
        self.frame_count += 1
        self.last_raw = np.full(N_RAW, np.nan)
        tile = self._resize(frame)
        raw = self.last_raw
        scores = {}
        scores["frequency"] = self._frequency(tile)
        st = noise_stats(tile)
        if st is not None:
            raw[RAW["noise_cv"]], raw[RAW["noise_mean"]] = st
        scores["noise"] = noise_score(st)
        st = ela_stats(tile)
        if st is not None:
            raw[RAW["ela_cv"]], raw[RAW["ela_mean"]] = st
        scores["ela"] = ela_score(st)
        scores["edge"] = self._edges(tile)
        ss, vs, nh = color_stats(tile)
        raw[RAW["sat_std"]], raw[RAW["val_std"]], raw[RAW["unique_hues"]] = ss, vs, nh
        scores["color"] = color_score(ss, vs, nh)
        scores["temporal"] = self._temporal(tile)
        combined = sum(scores[k] * self.weights[k] for k in self.weights)   # :94
        return {"scores": scores, "fake_probability": _clip01(combined),
                "analysis_type": "frame_forensic", "frame_number": self.frame_count}

    def analyze_fast(self, frame):
        self.frame_count += 1
        self.last_raw = np.full(N_RAW, np.nan)
        tile = self._resize(frame)
        scores = {}
        scores["frequency"] = self._frequency(tile)
        scores["temporal"] = self._temporal(tile)
        scores["edge"] = self._edges(tile)
        combined = sum(scores[k] * FAST_WEIGHTS[k] for k in FAST_WEIGHTS)   # :119
        return {"scores": scores, "fake_probability": _clip01(combined),
                "analysis_type": "frame_forensic_fast", "frame_number": self.frame_count}

    def reset(self):
        self.prev_frame_gray = None
        self.temporal_diffs.clear()
        self.frame_count = 0
