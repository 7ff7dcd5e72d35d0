"""Process-wide engine registry: one libdfd context per GPU, stream-slot allocation for the
single-stream reference-style objects (DeepfakeDetector / FrameForensicAnalyzer / TemporalTracker)."""
import threading

import torch

from . import _lib

_engines = {}
_lock = threading.Lock()
DEFAULT_SLOTS = 64


def default_device():
    if not torch.cuda.is_available():
        raise _lib.DfdError("no CUDA device visible: the B200 path has no CPU fallback")
    return torch.cuda.current_device()


def get_engine(device=None):
    """Shared engine for reference-style single-stream objects (small batches, many slots)."""
    from .engine import Engine
    dev = default_device() if device is None else (device.index if isinstance(device, torch.device) else int(device))
    with _lock:
        e = _engines.get(dev)
        if e is None:
            e = Engine(device=dev, max_streams=DEFAULT_SLOTS, max_batch=8, max_crop=2176)
            e._free = list(range(DEFAULT_SLOTS - 1, -1, -1))
            _engines[dev] = e
        return e


def alloc_slot(engine):
    with _lock:
        if not engine._free:
            raise _lib.DfdError("all per-stream state slots of the shared engine are in use; release() unused detectors")
        return engine._free.pop()


def free_slot(engine, slot):
    with _lock:
        if slot is not None and slot not in engine._free:
            engine._free.append(slot)
