"""CPU: the C-ABI library builds, loads and exports every symbol include/dfd.h declares (no compute calls)."""
import ctypes
import os
import re

import dfd_b200  # noqa: F401
from dfd_b200 import _lib, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "dfd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dfd.h but not exported by libdfd.so"
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)


def test_struct_sizes_and_blob_layout_agree_with_the_library():
    lib = _lib.load()
    assert lib.dfd_abi_version() == 2
    assert _lib.FORENSIC_BYTES == 192 and _lib.RECORD_BYTES == 72
    _, total = weights.blob_layout()
    assert total == lib.dfd_weights_blob_floats()
    cfg = _lib.Config()
    lib.dfd_default_config(ctypes.byref(cfg))
    assert (cfg.window_size, cfg.voting_window, cfg.detection_threshold) == (60, 10, 0.5)
    assert (cfg.face_weight, cfg.forensic_weight, cfg.blend_mode) == (0.70, 0.30, 0)


def test_no_cpu_fallback():
    """Without a GPU every compute entry must fail loudly."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dfd_b200.engine import Engine
    with pytest.raises(_lib.DfdError):
        Engine(device=0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "real-time-video-deepfake-detection_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
