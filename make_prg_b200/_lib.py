"""ctypes binding of libmprg.so (the C ABI declared in include/mprg.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible when a
context is created, the caller gets a loud MprgError.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("MPRG_LIB", PKG / "libmprg.so"))

MPRG_OK = 0
ERRORS = {-1: "no CUDA device", -2: "CUDA error", -3: "bad argument", -4: "partitioning error",
          -5: "internal error"}
IV_MATCH, IV_NONMATCH = 0, 1
NODE_LEAF, NODE_INTERVAL, NODE_CLUSTER = 0, 1, 2
LOCUS_OK, LOCUS_CURATION_ERROR = 0, 1


class MprgError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libmprg error {code} ({ERRORS.get(code, '?')}): {message}")
        self.code = code
        self.message = message

    def __reduce__(self):  # crosses process boundaries (shards of a multi-GPU run)
        return (MprgError, (self.code, self.message))


class Task(C.Structure):
    _fields_ = [("locus", C.c_int32), ("rows_off", C.c_int32), ("n_rows", C.c_int32),
                ("c0", C.c_int32), ("c1", C.c_int32)]


class Interval(C.Structure):
    _fields_ = [("start", C.c_int32), ("stop", C.c_int32), ("type", C.c_int32)]


TASK_DTYPE = np.dtype([("locus", "<i4"), ("rows_off", "<i4"), ("n_rows", "<i4"), ("c0", "<i4"),
                       ("c1", "<i4")])
INTERVAL_DTYPE = np.dtype([("start", "<i4"), ("stop", "<i4"), ("type", "<i4")])

_lib = None

P = C.c_void_p
I32, I64 = C.c_int32, C.c_int64

# name -> (restype, argtypes); every symbol include/mprg.h declares
SIGNATURES = {
    "mprg_create": (C.c_int, [C.c_int, C.POINTER(P)]),
    "mprg_destroy": (None, [P]),
    "mprg_last_error": (C.c_char_p, [P]),
    "mprg_device_info": (C.c_int, [P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mprg_launch_count": (I64, [P]),
    "mprg_scan_stats": (C.c_int, [P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(I64), C.c_int]),
    "mprg_path_counts": (C.c_int, [P, P, C.c_int]),
    "mprg_kmeans_stats": (C.c_int, [P, C.POINTER(C.c_double), C.POINTER(I64), C.POINTER(I64), C.c_int]),
    "mprg_set_workers": (C.c_int, [P, I32]),
    "mprg_set_wait_mode": (C.c_int, [P, I32]),
    "mprg_copy_stats": (C.c_int, [P, C.POINTER(I64), C.POINTER(I64), C.c_int]),
    "mprg_timer": (C.c_int, [P, C.c_int, C.POINTER(C.c_double)]),
    "mprg_scan_log": (C.c_int, [P, P, P, I32, C.POINTER(I32), C.c_int]),
    "mprg_read_yardstick": (C.c_int, [P, P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mprg_batch_upload": (C.c_int, [P, P, P, P, P, I32, C.POINTER(P)]),
    "mprg_batch_free": (None, [P, P]),
    "mprg_batch_flags": (C.c_int, [P, P, P]),
    "mprg_batch_download_packed": (C.c_int, [P, P, I32, P, I64, C.POINTER(I32)]),
    "mprg_scan_tasks": (C.c_int, [P, P, P, I32, P, I64, P, P, P]),
    "mprg_partition_tasks": (C.c_int, [P, P, P, I32, P, I64, I32, P, P, P]),
    "mprg_partition_consensus": (C.c_int, [P, P, P, I32, I32, P, I32, C.POINTER(I32)]),
    "mprg_dedupe_rows": (C.c_int, [P, P, P, I32, P, I64, P, P, P, P, P]),
    "mprg_kmer_counts": (C.c_int, [P, P, P, P, I32, C.POINTER(I32), C.POINTER(I32), P, I64]),
    "mprg_kmeans": (C.c_int, [P, P, I32, I32, I32, P, C.POINTER(C.c_double)]),
    "mprg_kmeans_mode": (C.c_int, [P, P, I32, I32, I32, P, C.POINTER(C.c_double), I32]),
    "mprg_one_ref_like": (C.c_int, [P, P, P, P, P, I32, P]),
    "mprg_cluster_tasks": (C.c_int, [P, P, P, I32, P, I64, I32, P, P, P]),
    "mprg_build": (C.c_int, [P, P, I32, I32, C.POINTER(P)]),
    "mprg_build_sub": (C.c_int, [P, P, I32, I32, P, C.POINTER(P)]),
    "mprg_build_ascii": (C.c_int, [P, P, P, P, P, I32, I32, I32, C.POINTER(P), C.POINTER(P)]),
    "mprg_build_packed": (C.c_int, [P, P, P, P, P, P, I32, I32, I32, C.POINTER(P), C.POINTER(P)]),
    "mprg_result_free": (None, [P]),
    "mprg_result_n_loci": (I32, [P]),
    "mprg_result_status": (I32, [P, I32]),
    "mprg_result_statuses": (C.c_int, [P, P, P]),
    "mprg_result_prg": (P, [P, I32, C.POINTER(I64)]),
    "mprg_result_n_nodes": (I32, [P, I32]),
    "mprg_result_n_sites": (I32, [P, I32]),
    "mprg_result_nodes": (C.c_int, [P, I32, P, P, P, P, P, P, P, P]),
    "mprg_result_row_pool_size": (I64, [P, I32]),
    "mprg_result_row_pool": (C.c_int, [P, I32, P]),
    "mprg_result_from_prgs": (C.c_int, [P, P, I32, C.POINTER(P)]),
    "mprg_replace_n": (C.c_int, [P, I32, I32]),
    "mprg_fasta_load": (C.c_int, [P, I32, I32, I32, C.POINTER(P)]),
    "mprg_fasta_free": (None, [P]),
    "mprg_fasta_packed": (C.c_int, [P, C.POINTER(P), C.POINTER(I64), C.POINTER(P), C.POINTER(P)]),
    "mprg_pack_rows": (C.c_int, [P, I32, I32, P, I64, C.POINTER(I32)]),
    "mprg_fasta_info": (C.c_int, [P, C.POINTER(I32), C.POINTER(P), C.POINTER(I64), C.POINTER(P), C.POINTER(P),
                                  C.POINTER(P), C.POINTER(P), C.POINTER(P)]),
    "mprg_fasta_titles": (P, [P, I32, C.POINTER(I64)]),
    "mprg_encode_prg": (C.c_int, [P, I64, P, I64, C.POINTER(I64)]),
    "mprg_prg_to_gfa": (C.c_int, [P, I64, P, I64, C.POINTER(I64)]),
    "mprg_writer_open": (C.c_int, [C.c_char_p, I32, C.POINTER(P)]),
    "mprg_writer_add": (C.c_int, [P, P, P, P, I32, I32]),
    "mprg_writer_add_ds": (C.c_int, [P, P, P, P, P, I32, I32, I32, I32]),
    "mprg_writer_close": (C.c_int, [P, C.POINTER(I64), C.POINTER(I64)]),
    "mprg_writer_abort": (None, [P]),
    "mprg_writer_error": (C.c_char_p, [P]),
    "mprg_merge_outputs": (C.c_int, [P, I32, C.c_char_p, I32, C.POINTER(I64), C.c_char_p, I64]),
}


def load():
    """dlopen libmprg.so and declare the prototypes.  Raises MprgError when the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MprgError(-5, f"{LIB_PATH} not found: build it with `python -m make_prg_b200.build` "
                            "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def ptr(a):
    """Raw pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(P)
