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
    if cols.get("name_code") is not None:        # optional column: only the `--extra` metrics read it
        a = np.ascontiguousarray(cols["name_code"], dtype=np.uint64)
        if len(a) != n:
            raise ValueError("name_code must have n_records entries")
        keep.append(a)
        b.name_code = a.ctypes.data if a.size else None
    if len(cols["cigar_off"]) != n + 1 or len(cols["seq_off"]) != n + 1:
        raise ValueError("cigar_off / seq_off must have n_records + 1 entries")
    return b, keep


def from_batch(b, copy=True):
    """Copy a PjBatch (e.g. filled by pjh_prep_decode) into owned numpy columns.  copy=False returns views of the library's
    own arrays instead (valid until the next decode / close of the prep handle): no second copy of a multi-GB shard."""
    n = b.n_records
    out = {}

    def arr(ptr, count, dt):
        if not ptr or count == 0:
            return np.zeros(count, dtype=dt)
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(count,))
        return a.copy() if copy else a

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
    if b.name_code:
        out["name_code"] = arr(b.name_code, n, np.uint64)
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


_M64 = (1 << 64) - 1


def name_code(qname, flag):
    """64-bit code of BamAlignment::deriveName() (bam_alignment.cc:233-242); same function as pjio::name_code."""
    s = qname.encode() if isinstance(qname, str) else bytes(qname)
    if flag & 0x1:
        s += b"_R1" if flag & 0x40 else b"_R2" if flag & 0x80 else b"_R?"
    h = 0xCBF29CE484222325
    for c in s:
        h = ((h ^ c) * 0x100000001B3) & _M64
    h ^= h >> 33; h = (h * 0xFF51AFD7ED558CCD) & _M64
    h ^= h >> 33; h = (h * 0xC4CEB9FE1A85EC53) & _M64
    h ^= h >> 33
    return h


def from_records(records):
    """records: iterable of dicts(tid,pos,flag,mapq,xs,cigar(str),seq(str or None),mtid,mpos) in BAM order."""
    cols = {k: [] for k, _ in COLUMNS}
    names = []
    cols["cigar_off"].append(0)
    cols["seq_off"].append(0)
    seq_bytes = bytearray()
    for r in records:
        cols["tid"].append(r["tid"]); cols["pos"].append(r["pos"]); cols["flag"].append(r.get("flag", 0))
        cols["mapq"].append(r.get("mapq", 60))
        if "name" in r:
            names.append(name_code(r["name"], r.get("flag", 0)))
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
    if names:
        if len(names) != len(out["pos"]):
            raise ValueError("either every record has a name or none has")
        out["name_code"] = np.array(names, dtype=np.uint64)
    return out


LEAN_COLUMNS = [("pos", np.int32), ("flag", np.uint16), ("mapq", np.uint8), ("xs", np.uint8), ("l_qseq", np.int32),
                ("n_cigar", np.uint16), ("cigar", np.uint32), ("seq2", np.uint8), ("seqx_pos", np.uint64), ("seqx_code", np.uint8)]


def lean_runs_from_batch(b, runs, n_runs, copy=False, with_whole=False):
    """Split a lean PjBatch describing a whole segment (pjh_plan_decode_lean) into one lean batch per target stretch.
    Returns a list of dicts: tid + the lean columns (views of the library's arrays unless copy=True; the exception
    positions are made relative to the stretch, so that small array is always a copy)."""
    n = b.n_records

    def arr(ptr, count, dt):
        if not ptr or count == 0:
            return np.zeros(count, dtype=dt)
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(count,))
        return a.copy() if copy else a

    whole = {"pos": arr(b.pos, n, np.int32), "flag": arr(b.flag, n, np.uint16), "mapq": arr(b.mapq, n, np.uint8), "xs": arr(b.xs, n, np.uint8),
             "l_qseq": arr(b.l_qseq, n, np.int32), "n_cigar": arr(b.n_cigar, n, np.uint16), "cigar": arr(b.cigar, b.n_cigar_total, np.uint32),
             "seq2": arr(b.seq2, b.n_seq2_bytes, np.uint8), "seqx_pos": arr(b.seqx_pos, b.n_seqx, np.uint64), "seqx_code": arr(b.seqx_code, b.n_seqx, np.uint8)}
    if b.mtid and b.mpos:
        whole["mtid"] = arr(b.mtid, n, np.int32)
        whole["mpos"] = arr(b.mpos, n, np.int32)
    if b.name_code:
        whole["name_code"] = arr(b.name_code, n, np.uint64)
    out = []
    ends = [(runs[k + 1].rec0, runs[k + 1].cig0, runs[k + 1].seq0, runs[k + 1].seqx0) for k in range(n_runs - 1)] + [(n, b.n_cigar_total, b.n_seq2_bytes, b.n_seqx)]
    for k in range(n_runs):
        r = runs[k]
        r1, c1, s1, x1 = ends[k]
        d = {"tid": int(r.tid)}
        for name in ("pos", "flag", "mapq", "xs", "l_qseq", "n_cigar", "mtid", "mpos", "name_code"):
            if name in whole:
                d[name] = whole[name][r.rec0:r1]
        d["cigar"] = whole["cigar"][r.cig0:c1]
        d["seq2"] = whole["seq2"][r.seq0:s1]
        d["seqx_pos"] = whole["seqx_pos"][r.seqx0:x1] - np.uint64(r.seq0 * 4)
        d["seqx_code"] = whole["seqx_code"][r.seqx0:x1]
        out.append(d)
    return (out, whole) if with_whole else out


def lean_batch_struct(d):
    """PjBatch (lean form) pointing at the arrays of one stretch returned by lean_runs_from_batch. Returns (struct, keepalive)."""
    keep = []
    b = L.PjBatch()
    n = len(d["pos"])
    b.n_records = n
    b.lean = 1
    b.const_tid = int(d["tid"])
    for name, dt in LEAN_COLUMNS:
        a = np.ascontiguousarray(d[name], dtype=dt)
        keep.append(a)
        setattr(b, name, a.ctypes.data if a.size else None)
    for name, dt in (("mtid", np.int32), ("mpos", np.int32), ("name_code", np.uint64)):
        if d.get(name) is not None:
            a = np.ascontiguousarray(d[name], dtype=dt)
            keep.append(a)
            setattr(b, name, a.ctypes.data if a.size else None)
    b.n_cigar_total = len(d["cigar"])
    b.n_seq2_bytes = len(d["seq2"])
    b.n_seqx = len(d["seqx_pos"])
    return b, keep


def lean_nbytes(d):
    return int(sum(np.asarray(v).nbytes for k, v in d.items() if k != "tid"))
