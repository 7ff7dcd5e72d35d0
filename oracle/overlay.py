"""Oracle (CPU, test infrastructure): result annotation, restated from the reference with the OpenCV calls it makes
(deepfake_detection.py:552-586 ``draw_detection_overlay``, :688-726 ``_draw_frame_analysis_overlay``).  Pinned to the
unmodified reference by tests/golden/overlay.json (tests/test_oracle_overlay.py)."""
import cv2


def draw_detection_overlay(frame, x, y, w, h, fake_prob, confidence_level, voting_stats):
    color = (0, 0, 255) if confidence_level == "FAKE" else (0, 255, 0)
    cv2.rectangle(frame, (x, y), (x + w, y + h), color, 3)
    if confidence_level == "FAKE":
        label = f"FAKE (Frame: {fake_prob*100:.0f}%)"
    else:
        label = f"REAL (Frame: {(1-fake_prob)*100:.0f}%)"
    label_size, _ = cv2.getTextSize(label, cv2.FONT_HERSHEY_SIMPLEX, 0.7, 2)
    cv2.rectangle(frame, (x, y - 30), (x + label_size[0] + 10, y), color, -1)
    cv2.putText(frame, label, (x + 5, y - 10), cv2.FONT_HERSHEY_SIMPLEX, 0.7, (255, 255, 255), 2)
    if voting_stats["total_frames"] > 0:
        info = (f"Votes: F:{voting_stats['fake_count']} R:{voting_stats['real_count']} "
                f"(Last {voting_stats['total_frames']} frames)")
        cv2.putText(frame, info, (x, y + h + 20), cv2.FONT_HERSHEY_SIMPLEX, 0.5, color, 1)
    return frame


def draw_frame_analysis_overlay(frame, fake_prob, confidence_level, forensic_result):
    h, w = frame.shape[:2]
    if confidence_level == "FAKE":
        color, label = (0, 0, 255), f"SUSPICIOUS ({fake_prob*100:.0f}%)"
    elif confidence_level == "REAL":
        color, label = (0, 255, 0), f"AUTHENTIC ({(1-fake_prob)*100:.0f}%)"
    else:
        color, label = (0, 200, 255), f"ANALYZING ({fake_prob*100:.0f}%)"
    cv2.rectangle(frame, (2, 2), (w - 2, h - 2), color, 2)
    overlay = frame.copy()
    cv2.rectangle(overlay, (0, 0), (w, 30), color, -1)
    cv2.addWeighted(overlay, 0.6, frame, 0.4, 0, frame)
    cv2.putText(frame, f"[Frame Analysis] {label}", (10, 20), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (255, 255, 255), 1)
    scores = forensic_result.get("scores", {})
    signals = [f"FFT:{scores.get('frequency',0)*100:.0f}", f"Noise:{scores.get('noise',0)*100:.0f}",
               f"ELA:{scores.get('ela',0)*100:.0f}", f"Edge:{scores.get('edge',0)*100:.0f}"]
    cv2.putText(frame, " | ".join(signals), (10, h - 15), cv2.FONT_HERSHEY_SIMPLEX, 0.35, color, 1)
    return frame
