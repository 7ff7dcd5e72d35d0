"""CPU oracle for the per-frame detection hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  It restates, on the CPU, the
algorithm the reference executes for the hot path so that the
CUDA path can be checked against it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package
(``real-time-video-deepfake-detection_b200/``) never does and fails loudly when
its CUDA library is missing.

Third-party arithmetic.  The reference's numerics live in libraries whose
sources are not under /root/reference: OpenCV (``opencv-python>=4.5.3`` in the
reference's requirements.txt; 4.13.0 in this image), NumPy (unpinned; 2.3.x
here), Pillow (``>=8.0.0``; 12.x here), PyTorch CPU, and
``efficientnet_pytorch`` (lukemelas; NOT in requirements.txt, not installable
here).  The oracle calls the first four exactly as the reference's call sites
do (cited per function) and restates the fifth in plain ``torch.nn.functional``
(``oracle/effnet.py``).

Pinning status:
  * forensic signals, TemporalTracker, CLAHE pre-processing, heuristics:
    PINNED -- ``tests/golden/make_golden.py`` imports the unmodified reference
    modules from /root/reference in the build container and the committed
    fixtures hold the reference's own outputs; ``tests/test_oracle_golden.py``
    replays them.
  * EfficientNet-B0 forward: the reference's tests hold no numeric vectors and
    ``efficientnet_pytorch`` cannot be imported, so it is pinned against an
    independent implementation (HF ``transformers`` EfficientNet, a port of
    the official TF model) with mapped weights -- "parity pinned to an
    independent port, not to the reference's dependency itself".
"""
