"""The product's __host__ __device__ pixel math (csrc/px_*.h), compiled for the
CPU by tests/hostcheck, must be bit-identical to the third-party calls the
reference makes (cv2 / PIL) -- checked here without a GPU so that the CUDA
kernels, which inline the same functions, only add indexing on top."""
import ctypes
import os
import subprocess

import cv2
import numpy as np
import pytest
import torch
import torch.nn.functional as F
from PIL import Image

import dfd_b200  # noqa: F401
from dfd_b200 import synth
from oracle import faceprep

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "real-time-video-deepfake-detection_b200", "csrc")


@pytest.fixture(scope="module")
def hc():
    so = os.path.join(HERE, "hostcheck", "libhostcheck.so")
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", CSRC, src, "-o", so])
    lib = ctypes.CDLL(so)
    lib.hc_init(0, 4096)
    lib.hc_np_mean.restype = ctypes.c_float
    lib.hc_np_std.restype = ctypes.c_float
    return lib


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def frames():
    rng = np.random.RandomState(5)
    for fam in synth.FAMILIES:
        yield fam, synth.make_frame(fam, 360, 640, rng)


def test_colour_conversions_exhaustive(hc):
    allc = np.stack(np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij"), -1)
    allc = np.ascontiguousarray(allc.reshape(-1, 1, 3).astype(np.uint8))
    n = allc.shape[0]
    for fn, code, ch in (("hc_bgr2gray", cv2.COLOR_BGR2GRAY, 1), ("hc_bgr2hsv", cv2.COLOR_BGR2HSV, 3),
                         ("hc_bgr2lab", cv2.COLOR_BGR2LAB, 3), ("hc_lab2bgr", cv2.COLOR_LAB2BGR, 3)):
        dst = np.zeros((n, 1, ch) if ch > 1 else (n, 1), np.uint8)
        getattr(hc, fn)(ptr(allc), ptr(dst), ctypes.c_long(n))
        assert np.array_equal(dst, cv2.cvtColor(allc, code)), fn


@pytest.mark.parametrize("shape", [(720, 1280), (1080, 1920), (2160, 3840), (480, 640), (120, 160), (333, 517), (256, 256), (97, 1001)])
def test_cv_resize_256(hc, shape):
    rng = np.random.RandomState(shape[0])
    for fam in ("uniform", "pink"):
        src = synth.make_frame(fam, shape[0], shape[1], rng)
        dst = np.zeros((256, 256, 3), np.uint8)
        hc.hc_cvresize(ptr(src), shape[0], shape[1], ptr(dst), 256, 256)
        assert np.array_equal(dst, cv2.resize(src, (256, 256), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("shape", [(300, 300), (97, 83), (50, 71), (400, 96), (64, 64), (223, 410), (8, 8), (161, 159), (96, 100), (17, 9)])
def test_clahe(hc, shape):
    rng = np.random.RandomState(shape[1])
    for fam in ("uniform", "pink", "blur", "gradient"):
        img = np.ascontiguousarray(synth.make_frame(fam, shape[0], shape[1], rng)[:, :, 1])
        dst = np.zeros_like(img)
        hc.hc_clahe(ptr(img), shape[0], shape[1], ptr(dst))
        ref = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(img)
        assert np.array_equal(dst, ref), (fam, int((dst != ref).sum()))


@pytest.mark.parametrize("shape", [(300, 300), (97, 83), (50, 71), (400, 96), (160, 160), (900, 700), (161, 159), (40, 40), (1200, 1000)])
def test_pil_resize_160(hc, shape):
    rng = np.random.RandomState(shape[0] + 1)
    src = synth.make_frame("uniform", shape[0], shape[1], rng)
    dst = np.zeros((160, 160, 3), np.uint8)
    hc.hc_pil_resize(ptr(src), shape[0], shape[1], ptr(dst), 160)
    ref = np.asarray(Image.fromarray(src).resize((160, 160), Image.BILINEAR))
    assert np.array_equal(dst, ref)


def test_torch_upsample_normalise(hc):
    rng = np.random.RandomState(3)
    src = rng.randint(0, 256, (160, 160, 3)).astype(np.uint8)
    dst = np.zeros((3, 224, 224), np.float32)
    hc.hc_torch_up_norm(ptr(src), 160, ptr(dst), 224)
    ref = faceprep.to_input(src)[0].numpy()
    assert np.abs(dst - ref).max() < 2e-6      # float path: rounding-order noise only


def test_jpeg_q90_roundtrip(hc):
    for fam, f in frames():
        tile = cv2.resize(f, (256, 256), interpolation=cv2.INTER_LINEAR)
        dst = np.zeros_like(tile)
        hc.hc_jpeg_roundtrip(ptr(tile), 256, 256, ptr(dst))
        ok, enc = cv2.imencode(".jpg", tile, [int(cv2.IMWRITE_JPEG_QUALITY), 90])
        ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
        assert np.array_equal(dst, ref), (fam, int((dst != ref).sum()))


def test_canny_laplacian_gauss(hc):
    for fam, f in frames():
        tile = cv2.resize(f, (256, 256), interpolation=cv2.INTER_LINEAR)
        g = cv2.cvtColor(tile, cv2.COLOR_BGR2GRAY)
        e = np.zeros_like(g)
        hc.hc_canny(ptr(g), 256, 256, ptr(e))
        assert np.array_equal(e, cv2.Canny(g, 50, 150)), fam
        lap = np.zeros((256, 256), np.float64)
        hc.hc_laplacian(ptr(g), 256, 256, ptr(lap))
        assert np.array_equal(lap, cv2.Laplacian(g, cv2.CV_64F)), fam
        res = np.zeros((256, 256), np.float32)
        hc.hc_gauss_resid(ptr(g), 256, 256, ptr(res))
        gf = g.astype(np.float32)
        assert np.array_equal(res, gf - cv2.GaussianBlur(gf, (5, 5), 0)), fam


def test_numpy_order_reductions(hc):
    rng = np.random.RandomState(9)
    for n in (3, 5, 7, 8, 9, 15, 16, 29, 30, 63, 64):
        a = (rng.rand(n) * 10).astype(np.float32)
        assert hc.hc_np_mean(ptr(a), n) == np.mean(a)
        assert hc.hc_np_std(ptr(a), n) == np.std(a)


def _reachable(*groups):
    """Every value a score function can return: score = 0.0; score += a; score += b; ... ; float(np.clip(score, 0, 1))
    (frame_analysis.py:154-180 and the other five), in the reference's order of additions."""
    import itertools
    out = set()
    for combo in itertools.product(*groups):
        s = 0.0
        for v in combo:
            if v:
                s += v
        out.add(float(np.clip(s, 0.0, 1.0)))
    return sorted(out)


def test_combined_score_is_pythons_compensated_sum(hc):
    """The combined forensic probability is Python's sum() over floats (Neumaier-compensated since CPython 3.12) of
    scores[k] * weights[k] in dict order (frame_analysis.py:94,119).  Enumerates EVERY reachable combination of the six
    (three) step scores and compares dfd_py_sum_products -- the function k_finalize inlines -- with sum() bit for bit; a plain
    running sum fails this test (it flips the strict `p > 0.5` vote for some hundred combinations)."""
    import itertools
    freq = _reachable((0.4, 0.2, 0), (0.25, 0.1, 0), (0.15, 0))
    noise = _reachable((0.5, 0.25, 0), (0.3, 0.1, 0))
    ela = _reachable((0.5, 0.2, 0), (0.2, 0.1, 0))
    edge = _reachable((0.35, 0.15, 0), (0.3, 0.1, 0))
    color = _reachable((0.3, 0.1, 0), (0.25, 0.1, 0), (0.25, 0.1, 0))
    temporal = _reachable((0.4, 0.2, 0), (0.3, 0.1, 0))
    hc.hc_py_sum_products_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_void_p]
    for groups, weights in (((freq, noise, ela, edge, color, temporal), (0.25, 0.20, 0.20, 0.15, 0.10, 0.10)),
                            ((freq, temporal, edge), (0.45, 0.25, 0.30))):
        combos = np.array(list(itertools.product(*groups)), np.float64)
        w = np.array(weights, np.float64)
        got = np.zeros(len(combos), np.float64)
        hc.hc_py_sum_products_batch(ptr(combos), ptr(w), len(weights), len(combos), ptr(got))
        wl = list(weights)
        want = np.array([sum(s * k for s, k in zip(row, wl)) for row in combos.tolist()], np.float64)
        assert np.array_equal(got, want), int((got != want).sum())
        plain = np.zeros(len(combos))
        for j in range(len(wl)):
            plain = plain + combos[:, j] * wl[j]
        print(len(combos), "combinations; a plain running sum differs on", int((plain != want).sum()),
              "and flips p > 0.5 on", int(((plain > 0.5) != (want > 0.5)).sum()))


def test_constant_division_is_exact(tmp_path):
    """dfd_div_const (3-instruction division by 255 / the ImageNet std constants in k_vpass_up_norm) equals the IEEE
    quotient: sampled run (every 97th float of the operand ranges) of the exhaustive checker; the full run is in its header."""
    import subprocess
    src = os.path.join(HERE, "hostcheck", "divconst_check.c")
    exe = str(tmp_path / "divconst_check")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", src, "-o", exe, "-lm"])
    for which in "0123":
        out = subprocess.run([exe, which, "97"], capture_output=True, text=True)
        assert out.returncode == 0 and " 0 mismatches" in out.stdout, out.stdout


@pytest.mark.parametrize("sub_bits", [0, 1024])
def test_jpeg_decoder_functions_match_cv2_imdecode(hc, sub_bits):
    """csrc/px_jpegdec.h (header parser, Huffman state machine, IDCT, fancy up-sampling, colour conversion -- the functions the
    device decoder jpegdec.cu inlines) on the CPU against cv2.imdecode: sequentially (sub_bits = 0) and as a simulation of the
    kernel's self-synchronising parallel schedule with 1024-bit subsequences.  Bit-exact, odd sizes included."""
    rng = np.random.RandomState(21)
    cases = [((405, 720), [cv2.IMWRITE_JPEG_QUALITY, 85]), ((97, 83), [cv2.IMWRITE_JPEG_QUALITY, 85]),
             ((64, 48), [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_OPTIMIZE, 1]),
             ((131, 77), [cv2.IMWRITE_JPEG_QUALITY, 92, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]),
             ((131, 77), [cv2.IMWRITE_JPEG_QUALITY, 60, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422]),
             ((1, 1), [cv2.IMWRITE_JPEG_QUALITY, 85])]
    for (h, w), params in cases:
        for fam in ("pink", "gradient", "natural"):
            img = synth.make_frame(fam, h, w, rng)
            ok, enc = cv2.imencode(".jpg", img, params)
            ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
            out = np.zeros_like(ref)
            rounds = ctypes.c_int(0)
            buf = np.ascontiguousarray(enc)
            rc = hc.hc_jpeg_decode(ptr(buf), ctypes.c_long(buf.size), ptr(out), sub_bits, ctypes.byref(rounds))
            assert rc == 0 and np.array_equal(out, ref), (h, w, params, fam, rc)
    # unsupported / invalid streams are recognised by the parser
    img = synth.make_frame("pink", 64, 64, rng)
    info = (ctypes.c_int * 7)()
    ok, prog = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    assert hc.hc_jpeg_info(ptr(np.ascontiguousarray(prog)), ctypes.c_long(prog.size), info) == -2
    ok, rst = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_RST_INTERVAL, 2])
    assert hc.hc_jpeg_info(ptr(np.ascontiguousarray(rst)), ctypes.c_long(rst.size), info) == -2
    junk = np.frombuffer(b"\x89PNG\r\n\x1a\n" + bytes(64), np.uint8)
    assert hc.hc_jpeg_info(ptr(junk), ctypes.c_long(junk.size), info) == -1
