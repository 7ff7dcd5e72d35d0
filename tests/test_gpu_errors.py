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
