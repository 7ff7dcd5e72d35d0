"""GPU parity: result annotation (reference deepfake_detection.py:552-586, 688-726) -- the device compositor
against the oracle (oracle/overlay.py, pinned to the unmodified reference by tests/golden/overlay.json)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import dfd_b200  # noqa: F401
from dfd_b200 import overlay, synth
from oracle import effnet as oeff, faceprep as ofp, forensics as ofor, overlay as oov, tracker as otr

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FRES = {"scores": {"frequency": 0.25, "noise": 0.5, "ela": 0.15, "edge": 0.65, "color": 0.1, "temporal": 0.0}}


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def eng():
    from dfd_b200.engine import Engine
    e = Engine(device=0, max_streams=4, max_batch=8, max_crop=512)
    yield e
    e.close()


def test_overlay_matches_reference_golden(eng):
    with open(os.path.join(HERE, "golden", "overlay.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        frame = synth.make_frame("pink", c["h"], c["w"], np.random.RandomState(c["seed"]))
        votes = {"fake_count": c["votes"][0], "real_count": c["votes"][1], "total_frames": c["votes"][2]}
        cl = overlay.CommandList(c["h"], c["w"])
        overlay.detection_overlay(cl, *c["box"], c["fake_prob"], c["verdict"], votes)
        got = eng.draw_overlay(torch.from_numpy(frame).cuda(), cl).cpu().numpy()
        assert sha(got) == c["detection_sha1"]
        cl = overlay.CommandList(c["h"], c["w"])
        overlay.frame_analysis_overlay(cl, c["fake_prob"], c["verdict"], FRES)
        got = eng.draw_overlay(torch.from_numpy(frame).cuda(), cl).cpu().numpy()
        assert sha(got) == c["frame_sha1"]


def test_overlay_random_boxes_and_pitched_frames(eng):
    rng = np.random.RandomState(17)
    for it in range(60):
        H, W = int(rng.randint(100, 800)), int(rng.randint(160, 1300))
        frame = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
        box = (int(rng.randint(-40, W)), int(rng.randint(-40, H)), int(rng.randint(1, 400)), int(rng.randint(1, 400)))
        p = float(rng.uniform(0, 1))
        verdict = ["FAKE", "REAL", "UNCERTAIN"][it % 3]
        votes = {"fake_count": int(rng.randint(0, 11)), "real_count": int(rng.randint(0, 11)), "total_frames": int(rng.randint(0, 11))}
        # a frame that is a view into a wider buffer (row pitch > 3 * W)
        wide = torch.zeros((H, W + 16, 3), dtype=torch.uint8, device="cuda")
        view = wide[:, :W]
        view.copy_(torch.from_numpy(frame))
        cl = overlay.CommandList(H, W)
        if it % 2:
            overlay.detection_overlay(cl, *box, p, verdict, votes)
            ref = oov.draw_detection_overlay(frame.copy(), *box, p, verdict, votes)
        else:
            overlay.frame_analysis_overlay(cl, p, verdict, FRES)
            ref = oov.draw_frame_analysis_overlay(frame.copy(), p, verdict, FRES)
        cmds, masks = cl.pack()
        import ctypes as C
        rc = eng.lib.dfd_draw_overlay(eng.h, C.c_void_p(view.data_ptr()), H, W, view.stride(0), cmds.ctypes.data_as(C.c_void_p),
                                      int(cmds.size), masks.ctypes.data_as(C.c_void_p), int(masks.size), eng._stream())
        assert rc == 0
        assert np.array_equal(view.cpu().numpy(), ref), (it, box, verdict)
        assert int(wide[:, W:].max()) == 0                      # nothing written outside the frame


def test_predict_returns_the_reference_annotation():
    """DeepfakeDetector.predict: two faces per frame (the second crop contains the first face's overlay, as in the reference's
    in-place drawing, deepfake_detection.py:611-634) and a face-less frame; the annotated frame equals the oracle's."""
    from dfd_b200 import deepfake_detection as dd
    sd = synth.make_state_dict()
    dd.load_model_weights(sd)
    det = dd.DeepfakeDetector(use_tta=False, num_tta_augmentations=1, detection_threshold=0.55)
    frames = synth.make_sequence("pink", 360, 640, 13, seed=31)
    boxes = [(200, 100, 150, 160), (230, 60, 200, 120)]        # the second box overlaps the first one's label
    o_an, o_tr = ofor.OracleForensicAnalyzer(), otr.OracleTemporalTracker(detection_threshold=0.55)
    for i, f in enumerate(frames):
        f = np.array(f)
        exp_frame = f.copy()
        faces = boxes if i % 5 != 4 else []
        out, trig, ff, res = det.predict(f, faces=faces)
        count = i + 1
        exp_f = o_an.analyze(exp_frame) if count % 3 == 0 else o_an.analyze_fast(exp_frame)
        if faces:
            for k, box in enumerate(faces):
                x, y, w, h = box
                # the device path's own probability drives the label; its parity with the oracle classifier is the fp32 gate
                p_dev = res["face_results"][k]["face_prob"]
                p_ref = float(ofp.heuristics(torch.sigmoid(oeff.forward(ofp.prepare(exp_frame, box), sd)).item(), h, w))
                assert abs(p_dev - p_ref) <= 1e-4, (i, k)
                o_tr.update(np.float64(p_dev))
                exp_frame = oov.draw_detection_overlay(exp_frame, x, y, w, h, p_dev, o_tr.get_confidence_level(),
                                                       o_tr.get_voting_stats())
        else:
            o_tr.update(exp_f["fake_probability"])
            exp_frame = oov.draw_frame_analysis_overlay(exp_frame, exp_f["fake_probability"], o_tr.get_confidence_level(), exp_f)
        assert out is f
        assert np.array_equal(out, exp_frame), (i, int((out != exp_frame).any(axis=2).sum()))
    # draw=False leaves the frame alone; a CUDA tensor frame is annotated in place
    f = np.array(frames[0]); keep = f.copy()
    out, _, _, _ = det.predict(f, faces=[boxes[0]], draw=False)
    assert np.array_equal(out, keep)
    ft = torch.from_numpy(keep).cuda()
    out, _, _, res = det.predict(ft, faces=[])
    assert out is ft and not np.array_equal(ft.cpu().numpy(), keep)
    det.release()
