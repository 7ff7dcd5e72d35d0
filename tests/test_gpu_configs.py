"""GPU parity at the BASELINE.json configurations that are not the bench workload:
config 3 (six forensic signals, 1080p, batch 64), config 4 (many streams through dfd_analyze_batch, fused
vote) and config 5 (4K frames, up to 8 variable-size boxes, fp32 accuracy mode)."""
import numpy as np
import pytest
import torch

import dfd_b200  # noqa: F401
from dfd_b200 import synth
from oracle import effnet as oeff, faceprep as ofp, forensics as ofor, tracker as otr

pytestmark = pytest.mark.gpu
NAMES = {0: "UNCERTAIN", 1: "REAL", 2: "FAKE"}


@pytest.fixture(scope="module")
def sd():
    return synth.make_state_dict()


def test_config3_forensics_1080p_batch64():
    from dfd_b200.engine import Engine
    eng = Engine(device=0, max_streams=64, max_batch=64, max_crop=64)
    try:
        rng = np.random.RandomState(3)
        bases = [synth.make_frame(f, 1080, 1920, rng) for f in synth.FAMILIES]
        steps = []
        for t in range(3):
            fr = np.stack([np.clip(np.roll(bases[s % 5], (s, 2 * s), (0, 1)).astype(np.int16) + (t * ((s % 3) - 1)), 0, 255).astype(np.uint8)
                           for s in range(64)])
            steps.append(fr)
        oracles = {s: ofor.OracleForensicAnalyzer() for s in (0, 1, 2, 3, 4, 37, 63)}
        for t, fr in enumerate(steps):
            full = int(t % 3 == 0)
            res = eng.forensic_to_numpy(eng.forensics_batch(torch.from_numpy(fr).cuda(), np.arange(64), [full] * 64))
            assert res.shape == (64,) and np.all(res["frame_number"] == t + 1)
            for s, o in oracles.items():
                exp = o.analyze(fr[s]) if full else o.analyze_fast(fr[s])
                raw = o.last_raw
                for k in range(15):
                    if not np.isnan(raw[k]):
                        assert abs(res[s]["raw"][k] - raw[k]) <= 1e-4 * max(abs(raw[k]), 1e-12), (t, s, ofor.RAW_NAMES[k])
                assert abs(res[s]["fake_probability"] - exp["fake_probability"]) < 0.26      # at most one step-score branch apart
            # identical frames in different streams give identical records (streams 0 and 5 differ only by a roll)
            assert np.all((res["fake_probability"] >= 0) & (res["fake_probability"] <= 1))
    finally:
        eng.close()


def test_config5_4k_variable_boxes_fp32(sd):
    from dfd_b200.engine import Engine
    eng = Engine(device=0, max_streams=4, max_batch=8, max_crop=1280)
    try:
        eng.load_state_dict(sd)
        rng = np.random.RandomState(5)
        frame = synth.make_frame("pink", 2160, 3840, rng)
        boxes = np.array([[100, 100, 1200, 1100], [2000, 300, 48, 48], [1500, 900, 641, 333], [3000, 1500, 79, 300],
                          [10, 1700, 400, 400], [2500, 50, 1000, 1200], [3700, 2000, 96, 120], [800, 1300, 257, 255]], np.int32)
        ft = torch.from_numpy(frame).cuda().unsqueeze(0)
        rec, fres, fprob = eng.analyze_batch(ft, [0], [1], boxes, np.zeros(8, np.int32), dtype="fp32", want_forensic=True)
        got = fprob.cpu().numpy()
        for i, b in enumerate(boxes):
            p = float(torch.sigmoid(oeff.forward(ofp.prepare(frame, b), sd)).item())
            p = float(ofp.heuristics(p, b[3], b[2]))
            assert abs(got[i] - p) <= 1e-4, (i, got[i], p)                     # north_star fp32 gate
        r = eng.records_to_numpy(rec)[0]
        assert abs(r["face_probability"] - got[0]) == 0 and r["vote_input"] == got[0]          # faces[0] feeds the vote
        exp = ofor.OracleForensicAnalyzer().analyze(frame)
        assert eng.forensic_to_numpy(fres)[0]["fake_probability"] == exp["fake_probability"]
    finally:
        eng.close()


def test_config4_many_streams_fused_vote(sd):
    """dfd_analyze_batch over 32 streams x 12 steps (bf16 classifier): the fused vote is bit-identical with the
    reference tracker fed the same per-frame probabilities, mixed face / no-face frames."""
    from dfd_b200.engine import Engine
    n, steps = 32, 12
    eng = Engine(device=0, max_streams=n, max_batch=n, max_crop=512, detection_threshold=0.55)
    try:
        eng.load_state_dict(sd)
        rng = np.random.RandomState(8)
        bases = [synth.make_frame(synth.FAMILIES[s % 5], 360, 640, rng) for s in range(n)]
        trackers = [otr.OracleTemporalTracker(detection_threshold=0.55) for _ in range(n)]
        for t in range(steps):
            fr = np.stack([np.clip(b.astype(np.int16) + rng.randint(-2, 3, b.shape[:2] + (1,)), 0, 255).astype(np.uint8) for b in bases])
            has_face = [(s + t) % 4 != 0 for s in range(n)]
            bf = np.array([s for s in range(n) if has_face[s]], np.int32)
            bx = synth.make_boxes(len(bf), 360, 640, rng, lo=60, hi=300)
            rec, fres, fprob = eng.analyze_batch(torch.from_numpy(fr).cuda(), np.arange(n), [int(t % 3 == 0)] * n, bx, bf,
                                                 dtype="bf16", want_forensic=True)
            rec = eng.records_to_numpy(rec)
            fres = eng.forensic_to_numpy(fres)
            fp = fprob.cpu().numpy()
            j = 0
            for s in range(n):
                if has_face[s]:
                    p = np.float64(fp[j]); j += 1
                    assert rec[s]["face_probability"] == p and rec[s]["vote_input"] == p
                else:
                    p = float(fres[s]["fake_probability"])
                    assert np.isnan(rec[s]["face_probability"]) and rec[s]["vote_input"] == p
                trackers[s].update(p)
                assert NAMES[int(rec[s]["verdict"])] == trackers[s].get_confidence_level(), (t, s)
                vs = trackers[s].get_voting_stats()
                assert (int(rec[s]["fake_count"]), int(rec[s]["real_count"])) == (vs["fake_count"], vs["real_count"])
                assert rec[s]["temporal_average"] == trackers[s].get_temporal_average()
                assert abs(rec[s]["stability_score"] - trackers[s].get_stability_score()) < 1e-12
                assert int(rec[s]["frame_count"]) == t + 1 and int(rec[s]["stream_id"]) == s
    finally:
        eng.close()
