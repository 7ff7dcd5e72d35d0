"""B200-native per-frame deepfake-detection hot path (drop-in for the
reference's DeepfakeDetector / FrameForensicAnalyzer / TemporalTracker /
DeepfakeEfficientNet surface).  Import as ``dfd_b200`` (see /dfd_b200.py).

Compute runs in hand-written sm_100a CUDA kernels behind the C-ABI declared in
``include/dfd.h`` (``csrc/`` -> ``libdfd.so``).  There is no CPU fallback: any
compute entry raises if the library is missing or no B200 is visible.
"""
__all__ = ["arch", "synth"]
