"""GPU: repeatability as a race detector.  compute-sanitizer is disabled on this GPU pool (profiles/sanitizer_r02.txt), so the
warp-specialised mbarrier / TMEM / TMA kernels and the self-synchronising JPEG decoder are launched many times on the same
input and every output must be bit-identical: a data race, a missing barrier or an uninitialised read shows up as nondeterminism."""
import cv2
import numpy as np
import pytest
import torch

import dfd_b200  # noqa: F401
from dfd_b200 import synth

pytestmark = pytest.mark.gpu


def test_classifier_repeatable_both_modes():
    from dfd_b200.engine import Engine
    e = Engine(device=0, max_streams=4, max_batch=64, max_crop=64)
    try:
        e.load_state_dict(synth.make_state_dict())
        g = torch.Generator().manual_seed(3)
        x = synth._calib_batch(g, 37).float().permute(0, 2, 3, 1).contiguous().cuda()      # ragged batch: partial tiles / clusters
        for xin in (x, x.bfloat16()):
            ref = e.effnet_forward(xin).clone()
            for _ in range(40):
                assert torch.equal(e.effnet_forward(xin), ref)
        # interleaved precisions and batch sizes reuse the same workspaces
        a = e.effnet_forward(x[:5].contiguous()).clone()
        for _ in range(10):
            e.effnet_forward(x.bfloat16())
            assert torch.equal(e.effnet_forward(x[:5].contiguous()), a)
    finally:
        e.close()


def test_jpeg_decode_repeatable_and_overlapped():
    """20 decodes of the same batch (bit-identical), then decodes on a side stream while the classifier runs on the main one."""
    from dfd_b200.engine import Engine
    e = Engine(device=0, max_streams=4, max_batch=16, max_crop=64)
    try:
        e.load_state_dict(synth.make_state_dict())
        rng = np.random.RandomState(31)
        imgs = [synth.make_frame(f, 405, 720, rng) for f in ("uniform", "pink", "natural", "gradient", "blur", "flat")]
        streams = [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 85])[1].tobytes() for im in imgs]
        refs = np.stack([cv2.imdecode(np.frombuffer(s, np.uint8), cv2.IMREAD_COLOR) for s in streams])
        packed, off = e.pack_jpegs(streams)
        for _ in range(20):
            frames, status = e.decode_jpeg_batch(packed, off, 405, 720)
            assert status.cpu().tolist() == [0] * 6 and np.array_equal(frames.cpu().numpy(), refs)
        g = torch.Generator().manual_seed(4)
        x = synth._calib_batch(g, 16).float().permute(0, 2, 3, 1).contiguous().cuda()
        z0 = e.effnet_forward(x).clone()
        side = torch.cuda.Stream()
        out = torch.empty((6, 405, 720, 3), dtype=torch.uint8, device="cuda")
        for _ in range(10):
            with torch.cuda.stream(side):
                e.decode_jpeg_batch(packed, off, 405, 720, out=out)
            z = e.effnet_forward(x)
            torch.cuda.synchronize()
            assert torch.equal(z, z0) and np.array_equal(out.cpu().numpy(), refs)
    finally:
        e.close()
