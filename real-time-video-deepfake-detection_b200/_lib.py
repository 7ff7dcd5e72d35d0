"""ctypes binding of libdfd.so (include/dfd.h).  Fails loudly: no CPU fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdfd.so")

N_RAW, N_SIGNALS = 16, 6
F32, BF16 = 0, 1
UNCERTAIN, REAL, FAKE = 0, 1, 2
VERDICT_NAMES = {UNCERTAIN: "UNCERTAIN", REAL: "REAL", FAKE: "FAKE"}


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_streams", C.c_int32), ("max_batch", C.c_int32), ("max_crop", C.c_int32),
                ("window_size", C.c_int32), ("voting_window", C.c_int32), ("detection_threshold", C.c_double),
                ("face_weight", C.c_double), ("forensic_weight", C.c_double), ("blend_mode", C.c_int32),
                ("reserved", C.c_int32)]


class ForensicResult(C.Structure):
    _fields_ = [("raw", C.c_double * N_RAW), ("scores", C.c_double * N_SIGNALS), ("fake_probability", C.c_double),
                ("frame_number", C.c_int32), ("full", C.c_int32)]


class VoteRecord(C.Structure):
    _fields_ = [("stream_id", C.c_int32), ("verdict", C.c_int32), ("fake_count", C.c_int32), ("real_count", C.c_int32),
                ("history_len", C.c_int32), ("frame_count", C.c_int32), ("last_vote", C.c_int32), ("reserved", C.c_int32),
                ("vote_input", C.c_double),
                ("temporal_average", C.c_double), ("stability_score", C.c_double), ("face_probability", C.c_double),
                ("forensic_probability", C.c_double)]


FORENSIC_BYTES = C.sizeof(ForensicResult)     # 192
RECORD_BYTES = C.sizeof(VoteRecord)           # 72

# name -> (restype, argtypes); every symbol include/dfd.h declares
_P, _I, _S = C.c_void_p, C.c_int, C.c_size_t
SYMBOLS = {
    "dfd_default_config": (None, [C.POINTER(Config)]),
    "dfd_abi_version": (_I, []),
    "dfd_create": (_I, [C.POINTER(Config), C.POINTER(_P)]),
    "dfd_destroy": (None, [_P]),
    "dfd_last_error": (C.c_char_p, [_P]),
    "dfd_weights_blob_floats": (_S, []),
    "dfd_load_weights": (_I, [_P, _P, _S]),
    "dfd_forensics_batch": (_I, [_P, _P, _I, _I, _I, _S, _I, _P, _P, _P, _P]),
    "dfd_face_prep_batch": (_I, [_P, _P, _I, _I, _I, _S, _I, _P, _P, _I, _P, _I, _P]),
    "dfd_effnet_forward": (_I, [_P, _P, _I, _I, _P, _P]),
    "dfd_face_probability": (_I, [_P, _P, _P, _I, _P, _P]),
    "dfd_face_prep_tta": (_I, [_P, _P, _I, _I, _I, _S, _I, _P, _P, _I, _I, _P, _P, _I, _P]),
    "dfd_face_probability_tta": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "dfd_set_calibrator": (_I, [_P, _I, _I, _P, _P, _P]),
    "dfd_draw_overlay": (_I, [_P, _P, _I, _I, _I, _P, _I, _P, _S, _P]),
    "dfd_vote_update": (_I, [_P, _P, _P, _P, _I, _P, _P]),
    "dfd_analyze_batch": (_I, [_P, _P, _I, _I, _I, _S, _I, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "dfd_decode_jpeg_batch": (_I, [_P, _P, _P, _I, _I, _I, _P, _S, _I, _P, _P]),
    "dfd_jpeg_info": (_I, [_P, _S, _P]),
    "dfd_reset_stream": (_I, [_P, _I, _P]),
    "dfd_reset_stream_part": (_I, [_P, _I, _I, _P]),
    "dfd_configure_stream": (_I, [_P, _I, _I, _I, C.c_double, _P]),
    "dfd_flight_report": (_I, [_P, C.c_char_p, _S]),
    "dfd_launch_count": (C.c_int64, [_P]),
    "dfd_profile_start": (_I, [_P, _P]),
    "dfd_profile_stop": (_I, [_P, C.c_char_p, _S, _P]),
    "dfd_dbg_tiles": (_I, [_P, _P, _P, _I, _P]),
    "dfd_dbg_jpeg_roundtrip": (_I, [_P, _P, _P, _I, _P]),
    "dfd_dbg_canny": (_I, [_P, _P, _P, _I, _P]),
    "dfd_dbg_face160": (_I, [_P, _I, _P, _P]),
    "dfd_dbg_face_clahe": (_I, [_P, _P, _I, _I, _S, _I, _P, _P, _I, _P, _P]),
    "dfd_dbg_set_tap": (_I, [_P, C.c_char_p]),
    "dfd_dbg_activation": (C.c_int64, [_P, C.c_char_p, _P, C.c_int64, _P]),
    "dfd_dbg_set_option": (_I, [_P, C.c_char_p, _I]),
    "dfd_gemm_selftest": (_I, [_P, _I, _I, _I, _I, _I, C.POINTER(C.c_double), _P]),
    "dfd_gemm_bench": (_I, [_P, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_double), _P]),
    "dfd_gemm_tf32_selftest": (_I, [_P, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_double), C.POINTER(C.c_double), _P]),
}

_lib = None


class DfdError(RuntimeError):
    pass


def load():
    """Load libdfd.so and bind every declared symbol (raises if anything is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DfdError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the export is missing
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib
