"""ctypes binding of include/drs_b200.h.

There is deliberately no fallback: if libdrs_b200.so is missing the import of
the library raises, and every call that returns a non-zero status raises
DrsError with the library's message.
"""
from __future__ import absolute_import, division, print_function

import ctypes as C
import os

from . import build as _build

# enums of drs_b200.h
OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_INDEX, ERR_COMM, ERR_HOST = range(7)
MODEL_DLRM, MODEL_WND, MODEL_NCF, MODEL_DIN, MODEL_MTWND, MODEL_DIEN = range(6)
INTERACT_CAT, INTERACT_DOT = 0, 1
PREC_FP32, PREC_TF32X3, PREC_TF32 = 0, 1, 2
MLP_BOT, MLP_TOP, MLP_FINAL, MLP_ATTN_BASE, MLP_TASK_BASE, MLP_RNN_BASE = 0, 1, 2, 16, 4096, 8192
(STAGE_POOLED, STAGE_BOT, STAGE_INTERACT, STAGE_TOP, STAGE_OUT, STAGE_OFFSETS, STAGE_TASK,
 STAGE_IDS) = range(8)
SHARD_TABLE, SHARD_ROW = 0, 1
ABI_VERSION = 7
SHARD_BLOB_BYTES = 192

PRECISIONS = {"fp32": PREC_FP32, "tf32x3": PREC_TF32X3, "tf32": PREC_TF32}
MODEL_TYPES = {"dlrm": MODEL_DLRM, "wnd": MODEL_WND, "ncf": MODEL_NCF, "din": MODEL_DIN,
               "mtwnd": MODEL_MTWND, "dien": MODEL_DIEN}


class DrsError(RuntimeError):
    def __init__(self, code, message):
        super(DrsError, self).__init__("drs_b200 error %d: %s" % (code, message))
        self.code = code


class Spec(C.Structure):
    _fields_ = [
        ("model_type", C.c_int32), ("n_tables", C.c_int32),
        ("table_rows", C.POINTER(C.c_int64)), ("emb_dim", C.c_int32),
        ("n_bot", C.c_int32), ("ln_bot", C.POINTER(C.c_int32)),
        ("n_top", C.c_int32), ("ln_top", C.POINTER(C.c_int32)),
        ("interaction_op", C.c_int32), ("interaction_itself", C.c_int32),
        ("max_batch", C.c_int32), ("max_lookups_per_table", C.c_int32),
        ("precision", C.c_int32), ("n_lanes", C.c_int32),
        ("n_tasks", C.c_int32), ("n_task", C.c_int32), ("ln_task", C.POINTER(C.c_int32)),
        ("task_sigmoid_layer", C.c_int32),
        ("hidden_size", C.c_int32),
        ("shard_mode", C.c_int32), ("shard_world", C.c_int32),
    ]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("gpu_ms", C.c_float),
                ("d2h_ms", C.c_float), ("total_ms", C.c_float)]


_PP = C.POINTER(C.c_void_p)
_SIGNATURES = {
    "drs_last_error": (C.c_char_p, []),
    "drs_abi_version": (C.c_int, []),
    "drs_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "drs_device_numa_node": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "drs_create": (C.c_int, [C.POINTER(Spec), C.c_int, _PP]),
    "drs_destroy": (C.c_int, [C.c_void_p]),
    "drs_out_dim": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "drs_load_table": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int]),
    "drs_load_fc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                              C.c_int, C.c_int]),
    "drs_stage_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, _PP, C.c_int, _PP, C.c_int]),
    "drs_infer_staged": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.POINTER(Timing)]),
    "drs_infer": (C.c_int, [C.c_void_p, C.c_void_p, _PP, C.c_int, _PP, C.c_int, C.c_void_p,
                            C.POINTER(Timing)]),
    "drs_enqueue_staged": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "drs_enqueue_host": (C.c_int, [C.c_void_p, C.c_void_p, _PP, C.c_int, _PP, C.c_int,
                                   C.c_void_p]),
    "drs_sync": (C.c_int, [C.c_void_p]),
    "drs_packed_words": (C.c_int64, [C.c_int64, C.c_int]),
    "drs_pack_ids": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "drs_wire_parse_requests": (C.c_int, [C.c_void_p, C.c_int64, C.c_char_p, C.c_void_p, C.c_int,
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                          C.POINTER(C.c_int32)]),
    "drs_wire_emit_responses": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p,
                                          C.c_int64, C.c_void_p, C.c_int64,
                                          C.POINTER(C.c_int64)]),
    "drs_hostpack_words": (C.c_int64, [C.c_int64, C.c_int]),
    "drs_hostpack_ids": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                   C.POINTER(C.c_int)]),
    "drs_enqueue_host_packed": (C.c_int, [C.c_void_p, C.c_void_p, _PP, C.POINTER(C.c_int32), _PP,
                                          C.c_int, C.c_void_p]),
    "drs_submit_staged": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                    C.POINTER(C.c_int64)]),
    "drs_submit_host": (C.c_int, [C.c_void_p, C.c_void_p, _PP, C.c_int, _PP, C.c_int, C.c_void_p,
                                  C.POINTER(C.c_int64)]),
    "drs_submit_many": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.c_void_p, C.POINTER(C.c_int64)]),
    "drs_wait": (C.c_int, [C.c_void_p, C.c_int64]),
    "drs_lane_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "drs_timer_start": (C.c_int, [C.c_void_p]),
    "drs_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "drs_sparse_lengths_sum": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64,
                                         C.c_void_p, C.c_int, C.c_void_p]),
    "drs_fetch_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                  C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "drs_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "drs_hbm_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "drs_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "drs_profile_begin": (C.c_int, [C.c_void_p, C.c_int]),
    "drs_profile_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float),
                                  C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "drs_host_alloc": (C.c_int, [_PP, C.c_int64]),
    "drs_host_free": (C.c_int, [C.c_void_p]),
    "drs_host_register": (C.c_int, [C.c_void_p, C.c_int64]),
    "drs_host_unregister": (C.c_int, [C.c_void_p]),
    "drs_shard_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "drs_shard_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "drs_exchange_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _PP, _PP,
                                     C.POINTER(C.c_int64), _PP]),
    "drs_exchange_end": (C.c_int, [C.c_void_p]),
}

_lib = None


def lib_path():
    # DRS_LIB_PATH: tuning experiments load an alternative build of the same library
    return os.environ.get("DRS_LIB_PATH") or _build.LIB_PATH


def load():
    """dlopen libdrs_b200.so (built in-tree). Raises if it is not there."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError(
                "libdrs_b200.so not found at %s: build it with "
                "`python -m deeprecsys_b200.build` (no CPU fallback exists)" % path)
        lib = C.CDLL(path)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if a symbol is missing
            fn.restype, fn.argtypes = restype, argtypes
        if lib.drs_abi_version() != ABI_VERSION:
            raise ImportError("libdrs_b200.so has ABI version %d, this binding needs %d: rebuild "
                              "with `python -m deeprecsys_b200.build`"
                              % (lib.drs_abi_version(), ABI_VERSION))
        _lib = lib
    return _lib


def check(code):
    if code != OK:
        raise DrsError(code, load().drs_last_error().decode("utf-8", "replace"))


def exported_symbols():
    return sorted(_SIGNATURES)
