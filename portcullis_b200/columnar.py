"""Columnar alignment batches (the pj_batch layout of include/portcullis_junc.h) as numpy arrays."""
import ctypes as C

import numpy as np

from . import _lib as L

COLUMNS = [("tid", np.int32), ("pos", np.int32), ("flag", np.uint16), ("mapq", np.uint8), ("xs", np.uint8),
           ("l_qseq", np.int32), ("mtid", np.int32), ("mpos", np.int32), ("cigar_off", np.uint32), ("cigar", np.uint32),
           ("seq_off", np.uint64), ("seq4", np.uint8)]

CIGAR_OPS = "MIDNSHP=XB"
NT16 = "=ACMGRSVTWYHKDBN"


def batch_struct(cols):
    """Build a PjBatch pointing at the numpy columns. Returns (struct, keepalive list)."""
    keep = []
    b = L.PjBatch()
    n = len(cols["pos"])
    b.n_records = n
    for name, dt in COLUMNS:
        a = np.ascontiguousarray(cols[name], dtype=dt)
        keep.append(a)
        setattr(b, name, a.ctypes.data if a.size else None)
    if len(cols["cigar_off"]) != n + 1 or len(cols["seq_off"]) != n + 1:
        raise ValueError("cigar_off / seq_off must have n_records + 1 entries")
    return b, keep


def from_batch(b):
    """Copy a PjBatch (e.g. filled by pjh_prep_decode) into owned numpy columns."""
    n = b.n_records
    out = {}

    def arr(ptr, count, dt):
        if not ptr or count == 0:
            return np.zeros(count, dtype=dt)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(count,)).copy()

    for name, dt in COLUMNS:
        if name in ("cigar", "seq4"):
            continue
        cnt = n + 1 if name in ("cigar_off", "seq_off") else n
        out[name] = arr(getattr(b, name), cnt, dt)
    if n == 0:
        out["cigar_off"] = np.zeros(1, np.uint32)
        out["seq_off"] = np.zeros(1, np.uint64)
    out["cigar"] = arr(b.cigar, int(out["cigar_off"][-1]), np.uint32)
    out["seq4"] = arr(b.seq4, int(out["seq_off"][-1]), np.uint8)
    return out


def encode_cigar(cigar_str):
    """'50M100N50M' -> list of BAM CIGAR words."""
    words, num = [], ""
    for ch in cigar_str:
        if ch.isdigit():
            num += ch
        else:
            words.append((int(num) << 4) | CIGAR_OPS.index(ch))
            num = ""
    return words


def encode_seq(seq):
    """Text SEQ -> BAM 4-bit packed bytes (high nibble first)."""
    codes = [NT16.index(c) if c in NT16 else 15 for c in seq.upper()]
    if len(codes) & 1:
        codes.append(0)
    return bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(codes), 2))


def from_records(records):
    """records: iterable of dicts(tid,pos,flag,mapq,xs,cigar(str),seq(str or None),mtid,mpos) in BAM order."""
    cols = {k: [] for k, _ in COLUMNS}
    cols["cigar_off"].append(0)
    cols["seq_off"].append(0)
    seq_bytes = bytearray()
    for r in records:
        cols["tid"].append(r["tid"]); cols["pos"].append(r["pos"]); cols["flag"].append(r.get("flag", 0))
        cols["mapq"].append(r.get("mapq", 60))
        xs = r.get("xs", 0)
        cols["xs"].append(ord(xs) if isinstance(xs, str) else xs)
        seq = r.get("seq")
        cols["l_qseq"].append(r.get("l_qseq", len(seq) if seq else 0))
        cols["mtid"].append(r.get("mtid", -1)); cols["mpos"].append(r.get("mpos", -1))
        w = encode_cigar(r["cigar"]) if isinstance(r["cigar"], str) else list(r["cigar"])
        cols["cigar"].extend(w); cols["cigar_off"].append(len(cols["cigar"]))
        if seq and any((x & 0xF) == 3 for x in w):
            seq_bytes += encode_seq(seq)
        cols["seq_off"].append(len(seq_bytes))
    out = {k: np.array(cols[k], dtype=dt) for k, dt in COLUMNS if k != "seq4"}
    out["seq4"] = np.frombuffer(bytes(seq_bytes), dtype=np.uint8).copy()
    return out
