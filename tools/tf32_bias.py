"""GPU probe: is the 3xTF32 GEMM's error a systematic scale bias?  python tools/tf32_bias.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfd_b200  # noqa
from dfd_b200.engine import Engine
e = Engine(device=0, max_streams=4, max_batch=4, max_crop=64)
for (N, K) in [(96, 16), (16, 32), (144, 24), (240, 40), (24, 96), (40, 144), (80, 240), (112, 480), (112, 672), (192, 1152), (1280, 320)]:
    rms, bias = e.gemm_tf32_selftest(12544, N, K, 0, 4)
    print(f"N={N:5d} K={K:5d} rms rel err {rms:.3e}  scale bias {bias:+.3e}", flush=True)
e.close()
