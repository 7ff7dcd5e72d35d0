"""Oracle (CPU, test infrastructure): face-crop preparation ("Oracle-A").

Reference order (deepfake_detection.py:357-389), with the MTCNN *detector*
replaced by the supplied box because face boxes are inputs (north_star;
SURVEY.md §8c): MTCNN's ``extract_face`` step -- crop, PIL BILINEAR resize to
160x160, un-normalised float32 CHW (``post_process=False``, reference
deepfake_detection.py:24-28) -- is applied to the whole crop.

    crop = frame[y:y+h, x:x+w]
    BGR->LAB, CLAHE(2.0, 8x8) on L, LAB->BGR            (:357-370)
    BGR->RGB -> PIL                                      (:376)
    PIL resize (160,160) BILINEAR -> float32 CHW 0..255  (MTCNN extract_face, :377)
    F.interpolate 224x224 bilinear, align_corners=False  (:382-383)
    /255, ImageNet mean/std                              (:384-389)

facenet-pytorch (>=2.5.2 in the reference's requirements.txt) is absent from
this image; its resize behaviour is restated from its published source
(``crop_resize``: ``img.crop(box).copy().resize((size,size), Image.BILINEAR)``).
"""
import cv2
import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def clahe_lab(face_bgr):
    """deepfake_detection.py:357-370 (``preprocess_face_quality``)."""
    lab = cv2.cvtColor(face_bgr.copy(), cv2.COLOR_BGR2LAB)
    l, a, b = cv2.split(lab)
    l = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(l)
    return cv2.cvtColor(cv2.merge([l, a, b]), cv2.COLOR_LAB2BGR)


def resize160(face_bgr_u8):
    rgb = Image.fromarray(cv2.cvtColor(face_bgr_u8, cv2.COLOR_BGR2RGB))
    return np.asarray(rgb.resize((160, 160), Image.BILINEAR))        # (160,160,3) u8 RGB


def to_input(rgb160_u8):
    """(160,160,3) u8 RGB -> (1,3,224,224) float32 normalised."""
    t = torch.from_numpy(np.ascontiguousarray(rgb160_u8)).permute(2, 0, 1).float().unsqueeze(0)
    t = F.interpolate(t, size=(224, 224), mode="bilinear", align_corners=False)
    t = t.to(torch.float32) / 255.0
    mean = torch.tensor(MEAN).view(1, 3, 1, 1)
    std = torch.tensor(STD).view(1, 3, 1, 1)
    return (t - mean) / std


def prepare(frame_bgr, box):
    x, y, w, h = (int(v) for v in box)
    crop = frame_bgr[y:y + h, x:x + w]
    return to_input(resize160(clahe_lab(crop)))


def heuristics(p, crop_h, crop_w):
    """deepfake_detection.py:489-502: +0.10 for crops under 80 px; clip [0,1]
    (np.float64 result)."""
    adj = 0.10 if (crop_h < 80 or crop_w < 80) else 0.0
    return np.clip(p + adj, 0, 1)


def tta_augment(face_bgr, flip, brightness, angle):
    """One augmentation of ``analyze_face_with_tta`` (deepfake_detection.py:417-434) with the random draws made explicit."""
    aug = face_bgr.copy()
    if flip:
        aug = cv2.flip(aug, 1)
    aug = cv2.convertScaleAbs(aug, alpha=brightness, beta=0)
    h, w = aug.shape[:2]
    M = cv2.getRotationMatrix2D((w / 2, h / 2), angle, 1.0)
    return cv2.warpAffine(aug, M, (w, h))


def tta_faces(face_bgr, params):
    """[preprocessed, augmented...]: the images ``_single_prediction`` sees under TTA (deepfake_detection.py:520-526, 408-438):
    CLAHE first, augmentations of the CLAHE'd crop."""
    pre = clahe_lab(face_bgr)
    return [pre] + [tta_augment(pre, *p) for p in params]


def to_input_noclahe(face_bgr_u8):
    return to_input(resize160(face_bgr_u8))
