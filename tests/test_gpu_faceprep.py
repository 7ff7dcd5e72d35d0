"""GPU parity: face-crop preparation vs Oracle-A (oracle/faceprep.py; CLAHE stage pinned to the reference)."""
import numpy as np
import pytest
import torch

import dfd_b200  # noqa: F401
from dfd_b200 import synth
from oracle import faceprep

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from dfd_b200.engine import Engine
    e = Engine(device=0, max_streams=8, max_batch=16, max_crop=1280)
    yield e
    e.close()


BOXES = [(100, 50, 300, 300), (0, 0, 97, 83), (400, 300, 50, 71), (640, 10, 400, 96), (7, 600, 64, 64),
         (800, 200, 223, 410), (20, 20, 900, 700), (1200, 650, 40, 40), (333, 111, 161, 159), (5, 5, 160, 160),
         (0, 0, 1280, 720), (1000, 400, 8, 8)]


@pytest.mark.parametrize("family", ["uniform", "pink", "gradient"])
def test_face_prep_stages_and_tensor(eng, family):
    rng = np.random.RandomState(11)
    frame = synth.make_frame(family, 720, 1280, rng)
    ft = torch.from_numpy(frame).cuda().unsqueeze(0)
    boxes = np.array(BOXES, np.int32)
    fidx = np.zeros(len(BOXES), np.int32)
    out = eng.face_prep_batch(ft, boxes, fidx, "fp32")
    out_bf = eng.face_prep_batch(ft, boxes, fidx, "bf16")
    for i, (x, y, w, h) in enumerate(BOXES):
        crop = frame[y:y + h, x:x + w]
        ref_clahe = faceprep.clahe_lab(crop)
        got_clahe = eng.dbg_face_clahe(ft, boxes, fidx, i).cpu().numpy()
        assert np.array_equal(got_clahe, ref_clahe), (i, "CLAHE/LAB stage not bit-exact")
        ref160 = faceprep.resize160(ref_clahe)
        assert np.array_equal(eng.dbg_face160(i).cpu().numpy(), ref160), (i, "PIL 160 stage not bit-exact")
        ref = faceprep.to_input(ref160)[0].permute(1, 2, 0).numpy()        # HWC
        got = out[i].cpu().numpy()
        assert np.abs(got - ref).max() < 2e-6, (i, np.abs(got - ref).max())
        gb = out_bf[i].float().cpu().numpy()
        assert np.array_equal(gb, torch.from_numpy(got).bfloat16().float().numpy())


def test_face_prep_multi_frame_batch(eng):
    rng = np.random.RandomState(3)
    frames = np.stack([synth.make_frame("pink", 360, 640, rng) for _ in range(4)])
    ft = torch.from_numpy(frames).cuda()
    boxes = synth.make_boxes(8, 360, 640, rng, lo=48, hi=300)
    fidx = np.array([0, 1, 2, 3, 3, 2, 1, 0], np.int32)
    out = eng.face_prep_batch(ft, boxes, fidx, "fp32").cpu().numpy()
    for i in range(8):
        ref = faceprep.prepare(frames[fidx[i]], boxes[i])[0].permute(1, 2, 0).numpy()
        assert np.abs(out[i] - ref).max() < 2e-6
