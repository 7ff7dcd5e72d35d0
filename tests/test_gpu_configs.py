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


def test_config2_classifier_batch256(sd):
    """BASELINE config 2 at its full size: 256 crops in one call, both precisions.  16 sampled logits against the oracle
    (fp32: north_star's 1e-4 gate; bf16: the regression bound of test_bf16_logits) and, for bf16, bit-identity of the
    first 24 logits with a batch-24 call (an image's result does not depend on the batch it is in)."""
    from dfd_b200.engine import Engine
    eng = Engine(device=0, max_streams=4, max_batch=256, max_crop=64)
    try:
        eng.load_state_dict(sd)
        g = torch.Generator().manual_seed(256)
        x = synth._calib_batch(g, 256).float()
        xn = x.permute(0, 2, 3, 1).contiguous().cuda()
        idx = list(range(0, 256, 17))[:16]
        ref = oeff.forward(x[idx], sd).flatten()
        z32 = eng.effnet_forward(xn).cpu()
        dp = (torch.sigmoid(z32[idx]) - torch.sigmoid(ref)).abs()
        print("config 2 fp32 b256: max |dp|", float(dp.max()))
        assert float(dp.max()) <= 1e-4
        zb = eng.effnet_forward(xn.bfloat16()).cpu()
        dpb = (torch.sigmoid(zb[idx]) - torch.sigmoid(ref)).abs()
        print("config 2 bf16 b256: max |dp|", float(dpb.max()), "mean", float(dpb.mean()))
        assert float(dpb.max()) <= 0.08
        z24 = eng.effnet_forward(xn[:24].contiguous().bfloat16()).cpu()
        assert torch.equal(z24, zb[:24])
        z24f = eng.effnet_forward(xn[:24].contiguous()).cpu()
        assert torch.equal(z24f, z32[:24])
    finally:
        eng.close()


def test_readme_blend_mode(sd):
    """north_star (4) / SURVEY a15: the README's 70/30 blend (opt-in, DFD_BLEND_README).  The vote input must be
    face_weight * face + forensic_weight * forensic in Python float arithmetic (two rounded products, one rounded sum),
    the face / forensic probabilities must match the oracle, and the vote must be the reference tracker's on those inputs."""
    from dfd_b200.engine import Engine
    n, steps = 8, 12
    eng = Engine(device=0, max_streams=n, max_batch=n, max_crop=512, detection_threshold=0.5, face_weight=0.70,
                 forensic_weight=0.30, blend_mode="readme")
    try:
        eng.load_state_dict(sd)
        rng = np.random.RandomState(15)
        bases = [synth.make_frame(synth.FAMILIES[s % 5], 360, 640, rng) for s in range(n)]
        trackers = [otr.OracleTemporalTracker(detection_threshold=0.5) for _ in range(n)]
        analyzers = [ofor.OracleForensicAnalyzer() for _ in range(n)]
        for t in range(steps):
            fr = np.stack([np.clip(b.astype(np.int16) + rng.randint(-2, 3, b.shape[:2] + (1,)), 0, 255).astype(np.uint8) for b in bases])
            has_face = [(s + t) % 5 != 0 for s in range(n)]
            bf = np.array([s for s in range(n) if has_face[s]], np.int32)
            bx = synth.make_boxes(len(bf), 360, 640, rng, lo=60, hi=300)
            full = int(t % 3 == 0)
            rec, fres, fprob = eng.analyze_batch(torch.from_numpy(fr).cuda(), np.arange(n), [full] * n, bx, bf, dtype="fp32",
                                                 want_forensic=True)
            rec = eng.records_to_numpy(rec)
            j = 0
            for s in range(n):
                exp = analyzers[s].analyze(fr[s]) if full else analyzers[s].analyze_fast(fr[s])
                forensic = exp["fake_probability"]
                assert rec[s]["forensic_probability"] == forensic
                if has_face[s]:
                    b = bx[j]; j += 1
                    p_or = float(torch.sigmoid(oeff.forward(ofp.prepare(fr[s], b), sd)).item())
                    p_or = float(ofp.heuristics(p_or, b[3], b[2]))
                    face = np.float64(rec[s]["face_probability"])
                    assert abs(face - p_or) <= 1e-4
                    want = 0.70 * face + 0.30 * forensic                    # Python / NumPy float64 arithmetic, no FMA
                    assert rec[s]["vote_input"] == want, (t, s, rec[s]["vote_input"], want)
                    trackers[s].update(want)
                else:
                    assert rec[s]["vote_input"] == forensic
                    trackers[s].update(forensic)
                assert NAMES[int(rec[s]["verdict"])] == trackers[s].get_confidence_level(), (t, s)
                vs = trackers[s].get_voting_stats()
                assert (int(rec[s]["fake_count"]), int(rec[s]["real_count"])) == (vs["fake_count"], vs["real_count"])
                assert rec[s]["temporal_average"] == trackers[s].get_temporal_average()
    finally:
        eng.close()


def _verdict_run(sd, dtype, n=64, steps=30):
    """Shared driver: n streams x steps frames through dfd_analyze_batch in `dtype`, against the ORACLE path (oracle
    forensics + Oracle-A face prep + fp32 oracle network + reference tracker fed ORACLE probabilities -- never the
    GPU's own).  Returns (frame-level verdict disagreements, vote disagreements, total, max |dp|)."""
    from dfd_b200.engine import Engine
    eng = Engine(device=0, max_streams=n, max_batch=n, max_crop=512, detection_threshold=0.55)
    try:
        eng.load_state_dict(sd)
        rng = np.random.RandomState(77)
        bases = [synth.make_frame(synth.FAMILIES[1 + s % 4], 240, 320, rng) for s in range(n)]
        boxes0 = synth.make_boxes(n, 240, 320, rng, lo=60, hi=200)
        trackers = [otr.OracleTemporalTracker(detection_threshold=0.55) for _ in range(n)]
        bad_verdict = bad_vote = total = 0
        near_threshold = [0]
        max_dp = 0.0
        for t in range(steps):
            fr = np.stack([np.clip(b.astype(np.int16) + rng.randint(-3, 4, b.shape), 0, 255).astype(np.uint8) for b in bases])
            bx = boxes0.copy()
            bx[:, 0] = np.clip(bx[:, 0] + rng.randint(-2, 3, n), 0, 320 - bx[:, 2])
            bx[:, 1] = np.clip(bx[:, 1] + rng.randint(-2, 3, n), 0, 240 - bx[:, 3])
            rec, _, fprob = eng.analyze_batch(torch.from_numpy(fr).cuda(), np.arange(n), [int(t % 3 == 0)] * n, bx,
                                              np.arange(n, dtype=np.int32), dtype=dtype)
            rec = eng.records_to_numpy(rec)
            xin = torch.cat([ofp.prepare(fr[s], bx[s]) for s in range(n)])
            p_or = torch.sigmoid(oeff.forward(xin, sd).flatten()).double().numpy()
            for s in range(n):
                p = np.float64(ofp.heuristics(float(np.float32(p_or[s])), bx[s][3], bx[s][2]))
                trackers[s].update(p)
                max_dp = max(max_dp, abs(float(rec[s]["face_probability"]) - float(p)))
                if int(rec[s]["last_vote"]) != int(p > 0.55):
                    bad_vote += 1
                    if abs(float(p) - 0.55) <= 1e-4:
                        near_threshold[0] += 1
                bad_verdict += int(NAMES[int(rec[s]["verdict"])] != trackers[s].get_confidence_level())
                total += 1
        return bad_verdict, bad_vote, total, max_dp, near_threshold[0]
    finally:
        eng.close()


def test_verdict_disagreement_rate_fp32(sd):
    """fp32 mode (the parity-green mode): verdicts against the ORACLE path fed ORACLE probabilities, 64 streams x 30
    frames.  |dp| <= 1e-4; a vote can only differ when the oracle probability lies within 1e-4 of the threshold."""
    bad_verdict, bad_vote, total, max_dp, near = _verdict_run(sd, "fp32")
    print(f"fp32: {bad_verdict} / {total} frame verdicts and {bad_vote} votes differ from the oracle path "
          f"({near} of them with the oracle probability within 1e-4 of the threshold); max |dp| {max_dp:.2e}")
    assert max_dp <= 1e-4
    assert bad_vote == near                      # a vote may differ only where 1e-4 straddles the threshold
    if near == 0:
        assert bad_verdict == 0


def test_verdict_disagreement_rate_bf16(sd):
    """bf16 mode: the same run.  The bf16 gate (5e-3) is not met on the synthetic weights (see test_bf16_logits), so
    verdicts CAN differ from the reference's; this test measures how often and bounds it (non-circular: the oracle
    tracker is fed oracle probabilities)."""
    bad_verdict, bad_vote, total, max_dp, _ = _verdict_run(sd, "bf16")
    print(f"bf16: {bad_verdict} / {total} frame verdicts ({100.0 * bad_verdict / total:.2f} %) and {bad_vote} votes "
          f"({100.0 * bad_vote / total:.2f} %) differ from the oracle path; max |dp| {max_dp:.4f}")
    assert max_dp <= 0.08
    assert bad_verdict <= 0.10 * total and bad_vote <= 0.10 * total
