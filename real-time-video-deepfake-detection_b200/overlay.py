"""Host side of the result annotation (reference deepfake_detection.py:552-586 ``get_box_color`` /
``draw_detection_overlay`` and :688-726 ``_draw_frame_analysis_overlay``): turns a verdict into the list of draw commands
``dfd_draw_overlay`` composites on the device copy of the frame.

Label formatting is the reference's f-strings.  Text: the string is rasterised once with OpenCV's Hershey font into a small
stroke mask (``cv2.putText`` on a zero canvas -- glyph outlines are font data; integer translations of the origin translate the
strokes exactly) and stamped by the kernel; rectangles, the blended status bar and every frame pixel are device work."""
import cv2
import numpy as np

CMD_DTYPE = np.dtype([("op", "<i4"), ("x0", "<i4"), ("y0", "<i4"), ("x1", "<i4"), ("y1", "<i4"), ("thickness", "<i4"),
                      ("color", "u1", (4,)), ("alpha", "<f4"), ("beta", "<f4"), ("mask_off", "<i4"), ("mask_w", "<i4"),
                      ("mask_h", "<i4"), ("reserved", "<i4")])
assert CMD_DTYPE.itemsize == 52          # sizeof(dfd_draw_cmd)
OUTLINE, FILL, BLEND, MASK = 1, 2, 3, 4
_PAD = 8
_mask_cache = {}


def text_mask(text, font_scale, thickness, org=None, frame_hw=None, font=cv2.FONT_HERSHEY_SIMPLEX):
    """(mask u8 [h, w], dx, dy): putText(img, text, (x, y), ...) inks exactly img[y + dy + r, x + dx + c] where mask[r, c] != 0.

    Away from the frame borders the strokes only translate with the origin, so the mask is cached per string.  OpenCV clips
    every stroke to the image BEFORE rasterising it, which moves a pixel or two of a stroke that crosses a border: a string
    whose padded box crosses a border is rasterised on a canvas whose edges coincide with the frame's on the crossed sides."""
    (tw, th), base = cv2.getTextSize(text, font, font_scale, thickness)
    pad = _PAD + thickness
    if org is not None and frame_hw is not None:
        H, W = frame_hw
        x0, y0, x1, y1 = org[0] - pad, org[1] - th - pad, org[0] + tw + pad, org[1] + base + pad
        if x0 < 0 or y0 < 0 or x1 > W or y1 > H:
            cx0, cy0, cx1, cy1 = max(x0, 0), max(y0, 0), min(x1, W), min(y1, H)
            if cx1 <= cx0 or cy1 <= cy0:
                return np.zeros((1, 1), np.uint8), 0, 0
            canvas = np.zeros((cy1 - cy0, cx1 - cx0), np.uint8)
            cv2.putText(canvas, text, (org[0] - cx0, org[1] - cy0), font, font_scale, 255, thickness)
            return canvas, cx0 - org[0], cy0 - org[1]
    key = (text, font_scale, thickness, font)
    hit = _mask_cache.get(key)
    if hit is not None:
        return hit
    canvas = np.zeros((th + base + 2 * pad, tw + 2 * pad), np.uint8)
    cv2.putText(canvas, text, (pad, pad + th), font, font_scale, 255, thickness)
    ys, xs = np.nonzero(canvas)
    if ys.size == 0:
        out = (np.zeros((1, 1), np.uint8), 0, 0)
    else:
        y0, y1, x0, x1 = ys.min(), ys.max() + 1, xs.min(), xs.max() + 1
        out = (np.ascontiguousarray(canvas[y0:y1, x0:x1]), int(x0) - pad, int(y0) - pad - th)
    if len(_mask_cache) > 4096:
        _mask_cache.clear()
    _mask_cache[key] = out
    return out


class CommandList:
    def __init__(self, frame_h, frame_w):
        self.frame_hw = (int(frame_h), int(frame_w))
        self.cmds, self.masks, self.mask_bytes = [], [], 0

    def _add(self, op, x0, y0, x1, y1, color, thickness=0, alpha=0.0, beta=0.0, mask=None):
        c = np.zeros((), CMD_DTYPE)
        c["op"], c["x0"], c["y0"], c["x1"], c["y1"], c["thickness"] = op, x0, y0, x1, y1, thickness
        c["color"][:3] = color
        c["alpha"], c["beta"] = alpha, beta
        if mask is not None:
            c["mask_off"], c["mask_w"], c["mask_h"] = self.mask_bytes, mask.shape[1], mask.shape[0]
            self.masks.append(mask.reshape(-1))
            self.mask_bytes += mask.size
        self.cmds.append(c)

    def rectangle(self, p0, p1, color, thickness):            # cv2.rectangle
        if thickness < 0:
            self._add(FILL, p0[0], p0[1], p1[0], p1[1], color)
        else:
            self._add(OUTLINE, p0[0], p0[1], p1[0], p1[1], color, thickness=thickness)

    def blended_rectangle(self, p0, p1, color, alpha, beta):   # overlay copy + filled rectangle + cv2.addWeighted
        self._add(BLEND, p0[0], p0[1], p1[0], p1[1], color, alpha=np.float32(alpha), beta=np.float32(beta))

    def put_text(self, text, org, font_scale, color, thickness):    # cv2.putText(FONT_HERSHEY_SIMPLEX)
        mask, dx, dy = text_mask(text, font_scale, thickness, org, self.frame_hw)
        self._add(MASK, org[0] + dx, org[1] + dy, 0, 0, color, mask=mask)

    def pack(self):
        cmds = np.array(self.cmds, CMD_DTYPE) if self.cmds else np.zeros(0, CMD_DTYPE)
        masks = np.concatenate(self.masks) if self.masks else np.zeros(0, np.uint8)
        return cmds, np.ascontiguousarray(masks, np.uint8)


def get_box_color(confidence_level):
    return (0, 0, 255) if confidence_level == "FAKE" else (0, 255, 0)          # deepfake_detection.py:552-557


def detection_overlay(cl, x, y, w, h, fake_prob, confidence_level, voting_stats):
    """draw_detection_overlay (deepfake_detection.py:559-586)."""
    color = get_box_color(confidence_level)
    cl.rectangle((x, y), (x + w, y + h), color, 3)
    if confidence_level == "FAKE":
        label = f"FAKE (Frame: {fake_prob*100:.0f}%)"
    else:
        label = f"REAL (Frame: {(1-fake_prob)*100:.0f}%)"
    label_size, _ = cv2.getTextSize(label, cv2.FONT_HERSHEY_SIMPLEX, 0.7, 2)
    cl.rectangle((x, y - 30), (x + label_size[0] + 10, y), color, -1)
    cl.put_text(label, (x + 5, y - 10), 0.7, (255, 255, 255), 2)
    if voting_stats["total_frames"] > 0:
        voting_info = (f"Votes: F:{voting_stats['fake_count']} R:{voting_stats['real_count']} "
                       f"(Last {voting_stats['total_frames']} frames)")
        cl.put_text(voting_info, (x, y + h + 20), 0.5, color, 1)


def frame_analysis_overlay(cl, fake_prob, confidence_level, forensic_result):
    """_draw_frame_analysis_overlay (deepfake_detection.py:688-726)."""
    h, w = cl.frame_hw
    if confidence_level == "FAKE":
        color, label = (0, 0, 255), f"SUSPICIOUS ({fake_prob*100:.0f}%)"
    elif confidence_level == "REAL":
        color, label = (0, 255, 0), f"AUTHENTIC ({(1-fake_prob)*100:.0f}%)"
    else:
        color, label = (0, 200, 255), f"ANALYZING ({fake_prob*100:.0f}%)"
    cl.rectangle((2, 2), (w - 2, h - 2), color, 2)
    cl.blended_rectangle((0, 0), (w, 30), color, 0.6, 0.4)
    cl.put_text(f"[Frame Analysis] {label}", (10, 20), 0.5, (255, 255, 255), 1)
    scores = forensic_result.get("scores", {})
    signals = [f"FFT:{scores.get('frequency',0)*100:.0f}", f"Noise:{scores.get('noise',0)*100:.0f}",
               f"ELA:{scores.get('ela',0)*100:.0f}", f"Edge:{scores.get('edge',0)*100:.0f}"]
    cl.put_text(" | ".join(signals), (10, h - 15), 0.35, color, 1)
