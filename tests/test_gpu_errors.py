"""GPU: error behaviour of the C-ABI (include/dfd.h): every entry returns a negative status with a message instead of
running on bad input, and the product has no silent fallback."""
import numpy as np
import pytest
import torch

import dfd_b200  # noqa: F401
from dfd_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def test_classifier_requires_weights_and_capacity():
    from dfd_b200.engine import Engine
    e = Engine(device=0, max_streams=4, max_batch=4, max_crop=64)
    try:
        x = torch.zeros((2, 224, 224, 3), dtype=torch.bfloat16, device="cuda")
        with pytest.raises(_lib.DfdError, match="load_weights"):
            e.effnet_forward(x)
        e.load_state_dict(synth.make_state_dict())
        e.effnet_forward(x)
        with pytest.raises(_lib.DfdError, match="max_batch"):
            e.effnet_forward(torch.zeros((5, 224, 224, 3), dtype=torch.bfloat16, device="cuda"))
    finally:
        e.close()


def test_frame_without_face_uses_forensic_probability():
    """analyze_batch with no boxes: the vote input is the forensic probability (deepfake_detection.py:652-655)."""
    from dfd_b200.engine import Engine
    from oracle import forensics as ofor
    e = Engine(device=0, max_streams=2, max_batch=2, max_crop=64, detection_threshold=0.55)
    try:
        e.load_state_dict(synth.make_state_dict())
        rng = np.random.RandomState(3)
        frame = synth.make_frame("blur", 360, 640, rng)
        rec, fres, fprob = e.analyze_batch(torch.from_numpy(frame).cuda().unsqueeze(0), [0], [1], None, None, dtype="bf16",
                                           want_forensic=True)
        torch.cuda.synchronize()
        r = e.records_to_numpy(rec)[0]
        want = ofor.OracleForensicAnalyzer().analyze(frame)["fake_probability"]
        assert abs(r["forensic_probability"] - want) < 1e-12
        assert abs(r["vote_input"] - want) < 1e-12
        assert np.isnan(r["face_probability"])
    finally:
        e.close()


def test_bad_configuration_is_rejected():
    from dfd_b200.engine import Engine
    with pytest.raises(_lib.DfdError):
        Engine(device=0, max_streams=0, max_batch=4, max_crop=64)
    with pytest.raises(_lib.DfdError):
        Engine(device=0, max_streams=4, max_batch=4, max_crop=64, voting_window=1000)


def test_oversized_and_out_of_frame_boxes_are_clamped_or_rejected():
    """boxes are clamped to the frame like the reference's numpy slicing; a box larger than max_crop or empty
    after clamping is rejected (probability NaN -> the vote falls back to the forensic probability) and nothing faults."""
    from dfd_b200.engine import Engine
    from oracle import effnet as oeff, faceprep as ofp
    sd = synth.make_state_dict()
    e = Engine(device=0, max_streams=4, max_batch=8, max_crop=256, detection_threshold=0.55)
    try:
        e.load_state_dict(sd)
        rng = np.random.RandomState(11)
        frame = synth.make_frame("pink", 480, 640, rng)
        boxes = np.array([[100, 50, 200, 180],        # valid
                          [500, 400, 300, 200],       # sticks out right/bottom: clamped to 140 x 80
                          [-20, -30, 120, 130],       # sticks out left/top: clamped to 100 x 100 at (0, 0)
                          [0, 0, 300, 100],           # wider than max_crop: rejected
                          [700, 10, 50, 50],          # entirely outside: empty -> rejected
                          [10, 10, 0, 40]], np.int32)  # zero width: rejected
        ft = torch.from_numpy(frame).cuda().unsqueeze(0)
        rec, fres, fprob = e.analyze_batch(ft, [0], [1], boxes, np.zeros(len(boxes), np.int32), dtype="fp32", want_forensic=True)
        torch.cuda.synchronize()
        got = fprob.cpu().numpy()
        assert np.isnan(got[3]) and np.isnan(got[4]) and np.isnan(got[5])
        for i, clamped in ((0, (100, 50, 200, 180)), (1, (500, 400, 140, 80)), (2, (0, 0, 100, 100))):
            x, y, w, h = clamped
            p = float(torch.sigmoid(oeff.forward(ofp.prepare(frame, np.array(clamped)), sd)).item())
            p = float(ofp.heuristics(p, h, w))
            assert abs(got[i] - p) <= 1e-4, (i, got[i], p)
        # a frame whose FIRST box is rejected votes with the forensic probability, like a frame without a face
        rec2, fres2, _ = e.analyze_batch(ft, [1], [1], boxes[3:4], np.zeros(1, np.int32), dtype="fp32", want_forensic=True)
        r = e.records_to_numpy(rec2)[0]
        assert np.isnan(r["face_probability"]) and r["vote_input"] == e.forensic_to_numpy(fres2)[0]["fake_probability"]
    finally:
        e.close()


def test_stream_id_out_of_range_is_flagged_not_faulted():
    from dfd_b200.engine import Engine
    e = Engine(device=0, max_streams=4, max_batch=4, max_crop=64)
    try:
        rng = np.random.RandomState(12)
        fr = np.stack([synth.make_frame("blur", 120, 160, rng) for _ in range(3)])
        res = e.forensic_to_numpy(e.forensics_batch(torch.from_numpy(fr).cuda(), [0, 4, -1], [1, 1, 1]))
        assert res[0]["frame_number"] == 1 and not np.isnan(res[0]["fake_probability"])
        assert res[1]["frame_number"] == -1 and np.isnan(res[1]["fake_probability"])
        assert res[2]["frame_number"] == -1 and np.isnan(res[2]["fake_probability"])
        rec = e.records_to_numpy(e.vote_update([0, 7, -3], [0.9, 0.9, 0.9]))
        assert rec[0]["verdict"] == 0 and rec[0]["history_len"] == 1
        assert rec[1]["verdict"] == -1 and rec[2]["verdict"] == -1
        # the engine is still healthy
        res = e.forensic_to_numpy(e.forensics_batch(torch.from_numpy(fr[:1]).cuda(), [0], [0]))
        assert res[0]["frame_number"] == 2
    finally:
        e.close()


def test_two_engines_on_two_devices_in_one_process():
    """function attributes are per device and every entry point selects its context's device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from dfd_b200.engine import Engine
    sd = synth.make_state_dict()
    g = torch.Generator().manual_seed(5)
    x = synth._calib_batch(g, 4).float().permute(0, 2, 3, 1).contiguous()
    outs = []
    engines = [Engine(device=d, max_streams=4, max_batch=8, max_crop=1024) for d in (0, 1)]
    try:
        for d, e in enumerate(engines):
            e.load_state_dict(sd)
        for d, e in enumerate(engines):          # device 0 stays "current" for torch while engine 1 runs
            xd = x.to(f"cuda:{d}")
            outs.append((e.effnet_forward(xd.bfloat16()).cpu(), e.effnet_forward(xd).cpu()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        rng = np.random.RandomState(2)
        frame = synth.make_frame("pink", 720, 1280, rng)
        box = np.array([[300, 200, 700, 450]], np.int32)       # 700 px wide: exercises the > 48 KB k_clahe_hpass launch
        ps = []
        for d, e in enumerate(engines):
            ft = torch.from_numpy(frame).to(f"cuda:{d}").unsqueeze(0)
            rec, _, fp = e.analyze_batch(ft, [0], [1], box, [0], dtype="bf16")
            ps.append(float(fp.cpu()[0]))
        assert ps[0] == ps[1]
    finally:
        for e in engines:
            e.close()


def test_two_engines_with_different_workspaces_on_one_device():
    """The dynamic shared-memory limit of a kernel belongs to the (device, function) pair: a second engine with a smaller
    max_crop must not lower it under the first one (regression: 'invalid argument' at k_clahe_hpass)."""
    from dfd_b200.engine import Engine
    from oracle import faceprep
    big = Engine(device=0, max_streams=2, max_batch=4, max_crop=2176)
    small = Engine(device=0, max_streams=2, max_batch=4, max_crop=256)
    try:
        rng = np.random.RandomState(4)
        frame = synth.make_frame("pink", 720, 1280, rng)
        ft = torch.from_numpy(frame).cuda().unsqueeze(0)
        box = np.array([[100, 100, 200, 180]], np.int32)
        for e in (big, small, big, small, big):
            out = e.face_prep_batch(ft, box, [0], "fp32")
            torch.cuda.synchronize()
            ref = faceprep.prepare(frame, box[0])[0].permute(1, 2, 0).numpy()
            assert np.abs(out[0].cpu().numpy() - ref).max() < 2e-6
    finally:
        big.close()
        small.close()
