"""ctypes binding of include/poccala_b200.h (the C ABI of the sm_100a engine).

The product path fails loudly when the shared library is missing or no B200 is present: there is
no CPU fallback and nothing here imports the test oracle.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POCCALA_B200_LIB") or os.path.join(_HERE, "_lib", "libpoccala_b200.so")  # (override: A/B runs)

OK = 0
ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM = -1, -2, -3, -4
DIM_MAX, XS, KA, EMIT, STATES, TRANS_SLOTS = 39, 40, 80, 3, 5, 9

_lib = None


class NativeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("poccala_b200 native error %d: %s" % (code, msg))
        self.code = code


class UnsupportedError(NativeError):
    pass


def _p(x):
    """Device/host pointer argument: torch tensor, numpy array, int or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    if hasattr(x, "ctypes"):
        return C.c_void_p(x.ctypes.data)
    raise TypeError("cannot pass %r as a pointer" % type(x))


_SIGS = {
    "pc_abi_version": (C.c_int, []),
    "pc_last_error": (C.c_char_p, []),
    "pc_rows_bytes": (C.c_int64, [C.c_int64]),
    "pc_gmm_bytes": (C.c_int64, [C.c_int64]),
    "pc_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "pc_destroy": (C.c_int, [C.c_void_p]),
    "pc_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "pc_get_option": (C.c_int64, [C.c_void_p, C.c_char_p]),
    "pc_corpus_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                   C.POINTER(C.c_void_p)]),
    "pc_corpus_destroy": (C.c_int, [C.c_void_p]),
    "pc_corpus_total_frames": (C.c_int64, [C.c_void_p]),
    "pc_corpus_total_tiles": (C.c_int64, [C.c_void_p]),
    "pc_corpus_active_tiles": (C.c_int64, [C.c_void_p]),
    "pc_corpus_emission_floats": (C.c_int64, [C.c_void_p]),
    "pc_corpus_total_pairs": (C.c_int64, [C.c_void_p]),
    "pc_corpus_total_states": (C.c_int64, [C.c_void_p]),
    "pc_corpus_frames_bytes": (C.c_int64, [C.c_void_p]),
    "pc_corpus_offsets": (C.c_int, [C.c_void_p] * 5),
    "pc_pack_gmm": (C.c_int, [C.c_void_p] * 6 + [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "pc_prepare_frames_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 4),
    "pc_prepare_frames_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 4),
    "pc_prepare_rows_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 4),
    "pc_prepare_rows_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 4),
    "pc_gmm_score": (C.c_int, [C.c_void_p] * 4 + [C.c_int32, C.c_void_p, C.c_void_p]),
    "pc_gmm_score_dense": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p]),
    "pc_forward_backward": (C.c_int, [C.c_void_p] * 10),
    "pc_log_bands": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pc_accumulate": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 4),
    "pc_transitions_max": (C.c_int, [C.c_void_p] * 6),
    "pc_transitions_sum": (C.c_int, [C.c_void_p] * 7),
    "pc_update_params": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 5
                         + [C.c_double, C.c_int32] + [C.c_void_p] * 5),
    "pc_viterbi": (C.c_int, [C.c_void_p] * 12),
    "pc_kmeans_workspace_bytes": (C.c_int64, [C.c_int32, C.c_void_p, C.c_int32]),
    "pc_kmeans_run": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
                      + [C.c_void_p] * 7 + [C.c_int64, C.c_void_p]),
    "pc_kmeans_finish": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
                         + [C.c_void_p] * 7),
    "pc_segment_keys": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 4),
    "pc_group_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int32]),
    "pc_group_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 4),
    "pc_gather_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 3),
    "pc_frame_moments_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 3),
    "pc_set_reduce_hook": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "pc_mfcc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_int32] * 4 + [C.c_void_p] + [C.c_int32] * 4
                + [C.c_void_p, C.c_void_p]),
    "pc_vad_distance": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 3 + [C.c_double, C.c_double]
                        + [C.c_void_p] * 3),
    "pc_peer_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p]),
    "pc_peer_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pc_peer_buffers": (C.c_int, [C.c_void_p, C.c_int32] + [C.POINTER(C.c_void_p)] * 3),
    "pc_update_params_peer": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_int32]
                              + [C.c_void_p] * 5),
    "pc_peer_destroy": (C.c_int, [C.c_void_p]),
    "pc_em_iteration_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
                             + [C.c_void_p] * 6 + [C.c_double, C.c_int32, C.c_void_p, C.c_void_p]),
}

# int hook(void *user, int32_t op, void *stream)  (pc_reduce_hook)
REDUCE_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_void_p)

EXPORTS = tuple(_SIGS)


def lib():
    """Load the shared library (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "poccala_b200: %s is missing - build it with `python -m poccala_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.pc_abi_version() != 3:
            raise RuntimeError("poccala_b200: ABI version mismatch")
        _lib = l
    return _lib


_CONSTS = {"PC_MAX_PEERS": 16}


def lib_const(name):
    """Constants of include/poccala_b200.h the host side needs."""
    return _CONSTS[name]


def check(code):
    if code == OK:
        return
    msg = lib().pc_last_error().decode("utf-8", "replace")
    if code == ERR_UNSUPPORTED:
        raise UnsupportedError(code, msg)
    raise NativeError(code, msg)


def call(name, *args):
    check(getattr(lib(), name)(*args))
