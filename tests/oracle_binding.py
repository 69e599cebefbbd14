"""ctypes binding of the CPU oracle (oracle/liboracle_junc.so) — TEST INFRASTRUCTURE ONLY.

Nothing under portcullis_b200/ imports this module; only tests/, __graft_entry__.smoke() and the
cpu_baseline / `--impl reference` legs of bench.py do.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from portcullis_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle_junc.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "portcullis_ref")
BAMTOOL = os.path.join(ORACLE_DIR, "_ref", "bamtool")

_o = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def load():
    global _o
    if _o is None:
        if not os.path.exists(ORACLE_LIB):
            build_oracle()
        o = C.CDLL(ORACLE_LIB)
        P = C.c_void_p
        o.oj_run.restype = C.c_int
        o.oj_run.argtypes = [C.POINTER(L.PjBatch), C.c_int32, P, P, P, C.c_int32, C.POINTER(P), C.POINTER(C.c_int64), P]
        o.oj_free.argtypes = [P]
        o.oj_last_error.restype = C.c_char_p
        o.oj_finalize.restype = C.c_int
        o.oj_finalize.argtypes = [P, C.c_int64, C.c_double]
        o.oj_extra.restype = C.c_int
        o.oj_extra.argtypes = [C.POINTER(L.PjBatch), C.c_int32, P, P, C.c_int64, C.c_int32, P, C.POINTER(C.c_int64)]
        o.oj_padded_query.restype = C.c_int
        o.oj_padded_query.argtypes = [C.c_int32, P, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_char_p,
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        o.oj_padded_genome.restype = C.c_int
        o.oj_padded_genome.argtypes = [C.c_int32, P, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_char_p]
        o.oj_entropy.restype = C.c_double
        o.oj_entropy.argtypes = [P, C.c_int64]
        o.oj_hamming.restype = C.c_int
        o.oj_hamming.argtypes = [C.c_char_p, C.c_char_p, C.c_int32]
        o.oj_revcomp.argtypes = [C.c_char_p, C.c_int32, C.c_char_p]
        o.oj_splice_motif.restype = C.c_int
        o.oj_splice_motif.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int32)]
        o.oj_min_anchor.restype = C.c_int32
        o.oj_min_anchor.argtypes = [C.c_int32] * 4
        _o = o
    return _o


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("[oracle %d] %s" % (code, msg))
        self.code = code


def run(cols, target_len, genomes, orientation=L.ORIENT["UNKNOWN"]):
    """cols: dict of numpy columns (see portcullis_b200.columnar); genomes: list of bytes per target.
    Returns (rows, stats) with rows a structured array (JUNCTION_DTYPE) before host finalize."""
    from portcullis_b200.columnar import batch_struct
    o = load()
    b, keep = batch_struct(cols)
    tl = np.ascontiguousarray(target_len, dtype=np.int32)
    cat = b"".join(genomes)
    off = np.zeros(len(genomes) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(g) for g in genomes])
    gbuf = C.create_string_buffer(cat, len(cat) + 1)
    rows_p = C.c_void_p()
    n = C.c_int64()
    stats = (L.PjTargetStats * len(tl))()
    rc = o.oj_run(C.byref(b), len(tl), tl.ctypes.data, C.addressof(gbuf), off.ctypes.data, orientation,
                  C.byref(rows_p), C.byref(n), C.addressof(stats))
    if rc:
        raise OracleError(rc, o.oj_last_error().decode())
    rows = np.empty(n.value, dtype=L.JUNCTION_DTYPE)
    if n.value:
        C.memmove(rows.ctypes.data, rows_p.value, n.value * L.JUNCTION_DTYPE.itemsize)
    o.oj_free(rows_p)
    st = np.array([(s.spliced_count, s.unspliced_count, s.sum_query_lengths, s.min_query_length, s.max_query_length)
                   for s in stats], dtype=[("spliced", "u8"), ("unspliced", "u8"), ("sumq", "u8"), ("minq", "i4"), ("maxq", "i4")])
    del keep
    return rows, st


def finalize(rows, mean_query_length):
    o = load()
    rows = np.ascontiguousarray(rows)
    rc = o.oj_finalize(rows.ctypes.data, len(rows), float(mean_query_length))
    if rc:
        raise OracleError(rc, o.oj_last_error().decode())
    return rows


def extra(cols, target_len, rows, max_query_length):
    """`--extra` metrics of the finalized oracle rows; returns (EXTRA_DTYPE array, reads dropped by htslib's pileup cap)."""
    from portcullis_b200.columnar import batch_struct
    o = load()
    b, keep = batch_struct(cols)
    tl = np.ascontiguousarray(target_len, dtype=np.int32)
    rows = np.ascontiguousarray(rows)
    out = np.zeros(len(rows), dtype=L.EXTRA_DTYPE)
    capped = C.c_int64()
    rc = o.oj_extra(C.byref(b), len(tl), tl.ctypes.data, rows.ctypes.data, len(rows), int(max_query_length),
                    out.ctypes.data, C.byref(capped))
    if rc:
        raise OracleError(rc, o.oj_last_error().decode())
    del keep
    return out, capped.value
