"""GPU parity: device JPEG ingest (csrc/jpegdec.cu, dfd_decode_jpeg_batch) against cv2.imdecode -- the reference's frame
ingest (backend_server.py:140-142) on the /analyze wire format (JPEG quality 0.85, <= 720 px; extension/content.js:86-109).
Bar: bit-exact (integer pipeline)."""
import cv2
import numpy as np
import pytest
import torch

import dfd_b200  # noqa: F401
from dfd_b200 import _lib, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from dfd_b200.engine import Engine
    e = Engine(device=0, max_streams=8, max_batch=16, max_crop=512)
    yield e
    e.close()


def _roundtrip(eng, imgs, params):
    streams, refs = [], []
    for im in imgs:
        ok, enc = cv2.imencode(".jpg", im, params)
        assert ok
        streams.append(enc.tobytes())
        refs.append(cv2.imdecode(enc, cv2.IMREAD_COLOR))
    H, W = refs[0].shape[:2]
    packed, offsets = eng.pack_jpegs(streams)
    frames, status = eng.decode_jpeg_batch(packed, offsets, H, W)
    torch.cuda.synchronize()
    assert status.cpu().tolist() == [0] * len(imgs)
    got = frames.cpu().numpy()
    for i, r in enumerate(refs):
        assert np.array_equal(got[i], r), (i, params, int((got[i] != r).sum()))
    return sum(len(s) for s in streams)


@pytest.mark.parametrize("hw", [(720, 1280), (405, 720), (480, 640), (97, 83), (241, 319), (16, 16), (8, 8), (1, 1), (17, 1), (1080, 1920)])
def test_wire_format_q85_420(eng, hw):
    """The extension's format: quality 85, 4:2:0, standard tables; every frame family in one batch; odd sizes exercise the
    partial-MCU edges of the fancy up-sampling (720 x 405 is what the extension actually sends for 720p video)."""
    rng = np.random.RandomState(hw[0] * 7 + hw[1])
    imgs = [synth.make_frame(f, hw[0], hw[1], rng) for f in synth.FAMILIES + ("natural",)]
    _roundtrip(eng, imgs, [cv2.IMWRITE_JPEG_QUALITY, 85])


@pytest.mark.parametrize("params", [
    [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_OPTIMIZE, 1],           # optimised Huffman tables (Chrome / Skia's encoder)
    [cv2.IMWRITE_JPEG_QUALITY, 30],
    [cv2.IMWRITE_JPEG_QUALITY, 100],
    [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444],
    [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422],
    [cv2.IMWRITE_JPEG_QUALITY, 75, cv2.IMWRITE_JPEG_LUMA_QUALITY, 60, cv2.IMWRITE_JPEG_CHROMA_QUALITY, 40],
])
def test_encoder_variants(eng, params):
    rng = np.random.RandomState(5)
    for hw in ((360, 640), (203, 301)):
        imgs = [synth.make_frame(f, hw[0], hw[1], rng) for f in ("pink", "gradient", "natural", "uniform")]
        _roundtrip(eng, imgs, params)


def test_grayscale_stream(eng):
    rng = np.random.RandomState(6)
    g = cv2.cvtColor(synth.make_frame("natural", 300, 444, rng), cv2.COLOR_BGR2GRAY)
    ok, enc = cv2.imencode(".jpg", g, [cv2.IMWRITE_JPEG_QUALITY, 85])
    ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
    packed, offsets = eng.pack_jpegs([enc.tobytes()])
    frames, status = eng.decode_jpeg_batch(packed, offsets, 300, 444)
    assert status.cpu().tolist() == [0] and np.array_equal(frames.cpu().numpy()[0], ref)


def test_unsupported_and_invalid_streams_fail_loudly(eng):
    """No CPU fallback: what the device decoder does not cover is an error, not a silent host decode."""
    rng = np.random.RandomState(7)
    im = synth.make_frame("natural", 128, 160, rng)
    ok, prog = cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    ok, rst = cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_RST_INTERVAL, 4])
    ok, base = cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 85])
    for bad in (prog.tobytes(), rst.tobytes()):
        packed, offsets = eng.pack_jpegs([bad])
        with pytest.raises(_lib.DfdError, match="baseline"):
            eng.decode_jpeg_batch(packed, offsets, 128, 160)
        with pytest.raises(_lib.DfdError):
            eng.jpeg_info(bad)
    packed, offsets = eng.pack_jpegs([b"not a jpeg at all"])
    with pytest.raises(_lib.DfdError, match="valid JPEG"):
        eng.decode_jpeg_batch(packed, offsets, 128, 160)
    packed, offsets = eng.pack_jpegs([base.tobytes()])
    with pytest.raises(_lib.DfdError, match="batch is"):
        eng.decode_jpeg_batch(packed, offsets, 100, 160)
    assert eng.jpeg_info(base.tobytes())[:3] == (128, 160, 3)
    # a stream cut in the middle of its entropy-coded data decodes to too few blocks: flagged per frame
    cut = base.tobytes()[: len(base) // 2] + b"\xff\xd9"
    packed, offsets = eng.pack_jpegs([base.tobytes(), cut])
    frames, status = eng.decode_jpeg_batch(packed, offsets, 128, 160)
    st = status.cpu().tolist()
    assert st[0] == 0 and st[1] == -4
    assert np.array_equal(frames.cpu().numpy()[0], cv2.imdecode(base, cv2.IMREAD_COLOR))


def test_decoded_frames_feed_the_path(eng):
    """JPEG bytes -> device decode -> dfd_forensics_batch gives exactly what the reference computes from cv2.imdecode."""
    from oracle import forensics as ofor
    rng = np.random.RandomState(8)
    imgs = [synth.make_frame(f, 405, 720, rng) for f in ("pink", "blur", "gradient", "natural")]
    streams = [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 85])[1].tobytes() for im in imgs]
    packed, offsets = eng.pack_jpegs(streams)
    frames, status = eng.decode_jpeg_batch(packed, offsets, 405, 720)
    res = eng.forensic_to_numpy(eng.forensics_batch(frames, [0, 1, 2, 3], [1, 1, 1, 1]))
    for i, s in enumerate(streams):
        ref = cv2.imdecode(np.frombuffer(s, np.uint8), cv2.IMREAD_COLOR)
        exp = ofor.OracleForensicAnalyzer().analyze(ref)
        assert res[i]["fake_probability"] == exp["fake_probability"]
