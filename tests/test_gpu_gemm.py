"""GPU: tcgen05/TMEM/TMA GEMM (csrc/gemm_tcgen05.cu) vs a CUDA-core fp32-accumulate reference on the
same bf16 operands, for every 1x1-conv shape of EfficientNet-B0 and ragged M."""
import pytest

import dfd_b200  # noqa: F401

pytestmark = pytest.mark.gpu

# (N, K) of every expand / project / head conv
LAYERS = [(16, 32), (96, 16), (24, 96), (144, 24), (24, 144), (40, 144), (240, 40), (40, 240), (80, 240), (480, 80),
          (80, 480), (112, 480), (672, 112), (112, 672), (192, 672), (1152, 192), (192, 1152), (320, 1152), (1280, 320)]


@pytest.fixture(scope="module")
def eng():
    from dfd_b200.engine import Engine
    e = Engine(device=0, max_streams=4, max_batch=4, max_crop=64)
    yield e
    e.close()


@pytest.mark.parametrize("N,K", LAYERS)
def test_gemm_layer_shapes(eng, N, K):
    # res bit 0: residual add; bit 1: SE-gated A operand staged by the producer warps (A_SCALE, 49 rows / image)
    for M, act, res in ((12544, 1, 0), (49 * 5, 0, 1), (128 * 148 * 2 + 77, 1, 1), (49 * 37, 0, 2), (12544, 0, 3)):
        err = eng.gemm_selftest(M, N, K, act, res)
        assert 0 <= err < 2e-2, (M, N, K, act, res, err)      # bf16 output rounding: <= 2^-8 relative


def test_gemm_tiny_and_ragged(eng):
    for M in (1, 7, 127, 128, 129, 300):
        err = eng.gemm_selftest(M, 96, 16, 1, 0)
        assert 0 <= err < 2e-2, (M, err)


@pytest.mark.parametrize("N,K", LAYERS + [(32, 32)])
def test_gemm_tf32x3_layer_shapes(eng, N, K):
    """3xTF32 tensor-core GEMM (csrc/gemm_tf32x3.cu, the fp32 accuracy mode) against an fp64-accumulated reference on the
    same fp32 operands: the error must be fp32-rounding sized (a single tf32 product would be ~5e-4), for every 1x1-conv
    shape, ragged M, residual and SE-gated operands.  Bound 6e-6 relative: the largest measured value is 4.1e-6 (N = 192,
    K = 1152, chunks of six k-blocks; 3.3e-6 with chunks of four); the gate that counts, |dp| <= 1e-4 on the whole network, is
    asserted in test_gpu_effnet.py / test_gpu_configs.py (measured 5.5e-6)."""
    for M, act, mode in ((12544, 1, 0), (49 * 5, 0, 1), (128 * 148 * 2 + 77, 1, 1), (49 * 37, 0, 2), (12544, 0, 3)):
        err, _ = eng.gemm_tf32_selftest(M, N, K, act, mode)
        assert 0 <= err < 6e-6, (M, N, K, act, mode, err)


def test_gemm_tf32x3_tiny_and_ragged(eng):
    for M in (1, 7, 127, 128, 129, 300):
        err, _ = eng.gemm_tf32_selftest(M, 96, 16, 1, 0)
        assert 0 <= err < 4e-6, (M, err)
