"""GPU parity: device-side vote ring buffers vs the reference's TemporalTracker traces (golden)."""
import json
import os

import numpy as np
import pytest
import torch

import dfd_b200  # noqa: F401
from oracle.tracker import OracleTemporalTracker

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
NAMES = {0: "UNCERTAIN", 1: "REAL", 2: "FAKE"}


def _cases():
    with open(os.path.join(G, "tracker.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"vw{c['voting_window']}_t{c['threshold']}_{len(c['probs'])}")
def test_vote_matches_reference_golden(case):
    from dfd_b200.engine import Engine
    eng = Engine(device=0, max_streams=8, max_batch=8, max_crop=64, voting_window=case["voting_window"],
                 detection_threshold=case["threshold"])
    try:
        for p, exp in zip(case["probs"], case["steps"]):
            r = eng.records_to_numpy(eng.vote_update([2], [p]))[0]
            assert NAMES[int(r["verdict"])] == exp["verdict"]
            assert (int(r["fake_count"]), int(r["real_count"])) == (exp["fake"], exp["real"])
            assert r["temporal_average"] == exp["avg"]
            assert abs(r["stability_score"] - exp["stab"]) <= 1e-12
    finally:
        eng.close()


def test_vote_many_streams_random_vs_oracle():
    from dfd_b200.engine import Engine
    rng = np.random.RandomState(5)
    n_streams, steps = 64, 80
    eng = Engine(device=0, max_streams=n_streams, max_batch=n_streams, max_crop=64, detection_threshold=0.55)
    trackers = [OracleTemporalTracker(detection_threshold=0.55) for _ in range(n_streams)]
    try:
        for t in range(steps):
            p = 0.55 + rng.choice([-1e-7, 0.0, 1e-7, -0.3, 0.3, 0.1], n_streams) * rng.choice([0, 1, 1], n_streams)
            p = np.clip(p, 0, 1)
            skip = rng.rand(n_streams) < 0.1
            vin = np.where(skip, np.nan, p)
            rec = eng.records_to_numpy(eng.vote_update(np.arange(n_streams), vin))
            for s in range(n_streams):
                trackers[s].update(None if skip[s] else float(p[s]))
                assert NAMES[int(rec[s]["verdict"])] == trackers[s].get_confidence_level(), (t, s)
                vs = trackers[s].get_voting_stats()
                assert int(rec[s]["fake_count"]) == vs["fake_count"] and int(rec[s]["real_count"]) == vs["real_count"]
                assert rec[s]["temporal_average"] == trackers[s].get_temporal_average()
        eng.reset(3)
        rec = eng.records_to_numpy(eng.vote_update([3], [np.nan]))[0]
        assert int(rec["verdict"]) == 0 and int(rec["history_len"]) == 0
    finally:
        eng.close()
