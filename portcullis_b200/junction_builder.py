"""Python mirror of the reference's junc-stage interface.

``JunctionBuilder`` follows portcullis::JunctionBuilder (/root/reference/src/junction_builder.cc:63-150):
constructor ``(prep_dir, output_prefix)``, the same setters, ``process()``.  ``JuncGpu`` is the lower
seam — the replacement of ``findJuncs`` — driving the C ABI of include/portcullis_junc.h directly with
columnar host buffers.  Everything computes in the CUDA library; there is no Python fallback.
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L
from .columnar import batch_struct, from_batch, lean_batch_struct, lean_runs_from_batch


def _check(rc, msg_fn):
    if rc != 0:
        m = msg_fn()
        raise L.PjError(rc, m.decode() if isinstance(m, bytes) else str(m))


class PrepDir:
    """A prepared data directory (PreparedFiles, /root/reference/src/prepare.hpp:74-145)."""

    def __init__(self, prep_dir, use_csi=False):
        self._lib = L.load()
        self._p = C.c_void_p()
        _check(self._lib.pjh_prep_open(os.fsencode(prep_dir), int(use_csi), C.byref(self._p)), self._lib.pjh_last_error)
        n = self._lib.pjh_prep_n_targets(self._p)
        self.names = [self._lib.pjh_prep_target_name(self._p, t).decode() for t in range(n)]
        self.lengths = np.array([self._lib.pjh_prep_target_len(self._p, t) for t in range(n)], dtype=np.int32)

    def target_records(self, tid):
        return self._lib.pjh_prep_target_records(self._p, tid)

    def plan_shards(self, n_gpus):
        """gpu_of_target[tid] for an n_gpus launch (LPT on index record counts); identical on every rank."""
        owner = np.zeros(len(self.names), dtype=np.int32)
        _check(self._lib.pjh_plan_shards(self._p, n_gpus, owner.ctypes.data), self._lib.pjh_last_error)
        return owner

    def plan(self, n_parts, seg_records=0, whole_targets=False):
        """(segments per part, gap cuts) of the range plan the junc driver executes for n_parts GPUs."""
        seg = np.zeros(n_parts, dtype=np.int32)
        cuts = C.c_int32()
        _check(self._lib.pjh_plan_describe(self._p, n_parts, int(whole_targets), int(seg_records), seg.ctypes.data, C.byref(cuts)), self._lib.pjh_last_error)
        return seg, cuts.value

    def decode_segment(self, n_parts, part, segment, seg_records=0, threads=1, whole_targets=False, copy=True):
        """Numpy columns of one segment of the plan (copy=False: views of the handle's arrays, valid until its next decode)."""
        b = L.PjBatch()
        _check(self._lib.pjh_plan_decode(self._p, n_parts, int(whole_targets), int(seg_records), part, segment, threads, C.byref(b)), self._lib.pjh_last_error)
        return from_batch(b, copy=copy)

    def decode_segment_lean(self, n_parts, part, segment, seg_records=0, threads=1, whole_targets=False, keep_mate=True, copy=False, with_whole=False):
        """One segment of the plan in the LEAN batch form, as a list of per-target stretches (dicts for JuncGpu.submit_lean);
        copy=False: views of the handle's arrays, valid until its next decode."""
        b = L.PjBatch()
        runs = C.POINTER(L.PjhLeanRun)()
        n_runs = C.c_int32()
        _check(self._lib.pjh_plan_decode_lean(self._p, n_parts, int(whole_targets), int(seg_records), part, segment, threads, int(keep_mate),
                                              C.byref(b), C.byref(runs), C.byref(n_runs)), self._lib.pjh_last_error)
        return lean_runs_from_batch(b, runs, n_runs.value, copy=copy, with_whole=with_whole)   # with_whole: also the unsplit arrays (e.g. to page-lock them once)

    def decode(self, tid=-1, threads=1, names=False):
        """Decode one target (or all with tid=-1) into owned numpy columns; names=True adds the name_code column."""
        b = L.PjBatch()
        self._lib.pjh_prep_want_names(self._p, 1 if names else 0)
        _check(self._lib.pjh_prep_decode(self._p, tid, threads, C.byref(b)), self._lib.pjh_last_error)
        return from_batch(b)

    def genome(self, tid):
        s = C.c_char_p()
        n = C.c_int64()
        _check(self._lib.pjh_prep_genome(self._p, tid, C.byref(s), C.byref(n)), self._lib.pjh_last_error)
        return C.string_at(s, n.value)

    def close(self):
        if self._p:
            self._lib.pjh_prep_close(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class JuncGpu:
    """One GPU context of the C ABI (pj_ctx)."""

    def __init__(self, device=0, orientation="UNKNOWN", match_group=0, legacy_sort=0, extra=False):
        self._lib = L.load()
        cfg = L.PjConfig()
        cfg.device = device
        cfg.extra_metrics = 1 if extra else 0    # keep what the `--extra` metrics need; batches must carry name_code
        cfg.reserved[0] = match_group        # lanes per (read, junction) pair in k_match; 0 = chosen from the data
        cfg.reserved[1] = legacy_sort        # 1 = multi-kernel radix sort instead of the one-sweep sort
        cfg.orientation = L.ORIENT[orientation] if isinstance(orientation, str) else int(orientation)
        self._ctx = C.c_void_p()
        _check(self._lib.pj_create(C.byref(cfg), C.byref(self._ctx)), self._lib.pj_global_last_error)
        self.n_targets = 0

    def _err(self):
        return self._lib.pj_last_error(self._ctx)

    def set_targets(self, lengths):
        tl = np.ascontiguousarray(lengths, dtype=np.int32)
        _check(self._lib.pj_targets_set(self._ctx, len(tl), tl.ctypes.data), self._err)
        self.n_targets = len(tl)

    def set_genome(self, tid, bases):
        buf = bytes(bases)
        _check(self._lib.pj_genome_set_target(self._ctx, tid, C.cast(C.c_char_p(buf), C.c_void_p), len(buf)), self._err)

    def shard_begin(self, n_records=0, n_cigar=0, n_seq=0):
        _check(self._lib.pj_shard_begin(self._ctx, n_records, n_cigar, n_seq), self._err)

    def submit(self, cols):
        b, keep = batch_struct(cols)
        _check(self._lib.pj_batch_submit(self._ctx, C.byref(b)), self._err)
        del keep

    def submit_lean(self, stretch):
        """One lean batch (a stretch of PrepDir.decode_segment_lean): the form the junc driver ships over PCIe."""
        b, keep = lean_batch_struct(stretch)
        _check(self._lib.pj_batch_submit(self._ctx, C.byref(b)), self._err)
        del keep

    def submit_pinned(self, cols):
        """Copy the columns into the context's pinned staging slot, then enqueue the async H2D."""
        n = len(cols["pos"])
        st = L.PjBatch()
        _check(self._lib.pj_staging_acquire(self._ctx, n, len(cols["cigar"]), len(cols["seq4"]), C.byref(st)), self._err)
        for name, dt in __import__("portcullis_b200.columnar", fromlist=["COLUMNS"]).COLUMNS:
            a = np.ascontiguousarray(cols[name], dtype=dt)
            if a.size:
                C.memmove(getattr(st, name), a.ctypes.data, a.nbytes)
        if st.name_code and cols.get("name_code") is not None and n:
            a = np.ascontiguousarray(cols["name_code"], dtype=np.uint64)
            C.memmove(st.name_code, a.ctypes.data, a.nbytes)
        st.n_records = n
        _check(self._lib.pj_batch_submit(self._ctx, C.byref(st)), self._err)

    def run(self):
        _check(self._lib.pj_shard_run(self._ctx), self._err)
        return self._lib.pj_shard_num_junctions(self._ctx)

    def fetch(self, out=None):
        """Rows + per-target stats of the last run.  out: optional preallocated JUNCTION_DTYPE array (e.g. a view of pinned
        host memory) with room for the rows; a slice of it is returned instead of a fresh pageable array."""
        n = self._lib.pj_shard_num_junctions(self._ctx)
        if out is not None:
            if out.dtype != L.JUNCTION_DTYPE or len(out) < n or not out.flags["C_CONTIGUOUS"]:
                raise ValueError("fetch(out=...): need a contiguous JUNCTION_DTYPE array with at least %d rows" % n)
            rows = out[:max(n, 0)]
        else:
            rows = np.zeros(max(n, 0), dtype=L.JUNCTION_DTYPE)
        stats = (L.PjTargetStats * self.n_targets)()
        _check(self._lib.pj_shard_fetch(self._ctx, rows.ctypes.data, len(rows), C.addressof(stats), self.n_targets), self._err)
        st = np.array([(s.spliced_count, s.unspliced_count, s.sum_query_lengths, s.min_query_length, s.max_query_length)
                       for s in stats], dtype=[("spliced", "u8"), ("unspliced", "u8"), ("sumq", "u8"), ("minq", "i4"), ("maxq", "i4")])
        return rows, st

    # ---- `--extra` metrics (contexts created with extra=True, after run()) ----
    def export_names(self):
        n = self._lib.pj_extra_num_spliced_names(self._ctx)
        codes = np.zeros(max(n, 0), dtype=np.uint64)
        _check(self._lib.pj_extra_export_names(self._ctx, codes.ctypes.data, len(codes)), self._err)
        return codes

    def import_names(self, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        _check(self._lib.pj_extra_import_names(self._ctx, codes.ctypes.data, len(codes)), self._err)

    def extra_run(self, max_query_length):
        """up_aln / down_aln / mm_n / mm_m of the shard's junctions (pj_shard_fetch order)."""
        n = self._lib.pj_shard_num_junctions(self._ctx)
        out = np.zeros(max(n, 0), dtype=L.EXTRA_DTYPE)
        _check(self._lib.pj_extra_run(self._ctx, int(max_query_length), out.ctypes.data, len(out)), self._err)
        return out

    def extra_timing(self):
        """(device ms, kernel launches, [(stage, ms)]) of the last extra_run()."""
        ms = C.c_float()
        nl = C.c_int32()
        _check(self._lib.pj_extra_timing(self._ctx, C.byref(ms), C.byref(nl)), self._err)
        k = C.c_int32()
        self._lib.pj_extra_kernel_times(self._ctx, 0, None, None, C.byref(k))
        tm = (C.c_float * k.value)()
        nm = (C.c_char_p * k.value)()
        self._lib.pj_extra_kernel_times(self._ctx, k.value, tm, nm, C.byref(k))
        return ms.value, nl.value, [(nm[i].decode(), tm[i]) for i in range(k.value)]

    def target_pileup(self, tid):
        cov = C.c_int32()
        mx = C.c_uint32()
        _check(self._lib.pj_extra_target_pileup(self._ctx, tid, C.byref(cov), C.byref(mx)), self._err)
        return bool(cov.value), mx.value

    def coverage(self, depth_tid, starts, ends):
        s = np.ascontiguousarray(starts, dtype=np.int32)
        e = np.ascontiguousarray(ends, dtype=np.int32)
        out = np.zeros((len(s), 4), dtype=np.uint32)
        _check(self._lib.pj_extra_coverage(self._ctx, int(depth_tid), len(s), s.ctypes.data, e.ctypes.data, out.ctypes.data), self._err)
        return out

    def coverage_batch(self, depth_tid, starts, ends):
        """cov_sum of junctions of many targets in one launch; depth_tid[j] = target whose depth vector junction j uses (-1: none)."""
        d = np.ascontiguousarray(depth_tid, dtype=np.int32)
        s = np.ascontiguousarray(starts, dtype=np.int32)
        e = np.ascontiguousarray(ends, dtype=np.int32)
        out = np.zeros((len(s), 4), dtype=np.uint32)
        _check(self._lib.pj_extra_coverage_batch(self._ctx, len(s), d.ctypes.data, s.ctypes.data, e.ctypes.data, out.ctypes.data), self._err)
        return out

    def extra(self, rows, max_query_length):
        """All four `--extra` columns for a single-context run: rows as returned by fetch().  Returns
        (EXTRA_DTYPE array in the order of rows, {tid: live-read maximum} of the targets where htslib's pileup read cap was replayed)."""
        x = self.extra_run(max_query_length)
        covered = np.zeros(self.n_targets, dtype=np.uint8)
        over = {}
        for t in range(self.n_targets):
            c, mx = self.target_pileup(t)
            covered[t] = c
            if mx >= 8000:
                over[t] = mx
        src = coverage_source(covered)
        if len(rows):
            x["cov_sum"] = self.coverage_batch(src[rows["tid"]], rows["start"], rows["end"])
        return extra_finalize(x), over

    def timing(self):
        ms = C.c_float()
        nl = C.c_int32()
        self._lib.pj_shard_timing(self._ctx, C.byref(ms), C.byref(nl))
        k = C.c_int32()
        self._lib.pj_shard_kernel_times(self._ctx, 0, None, None, C.byref(k))
        tm = (C.c_float * k.value)()
        nm = (C.c_char_p * k.value)()
        self._lib.pj_shard_kernel_times(self._ctx, k.value, tm, nm, C.byref(k))
        return ms.value, nl.value, [(nm[i].decode(), tm[i]) for i in range(k.value)]

    def close(self):
        if self._ctx:
            self._lib.pj_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# layout of pj_target_stats (include/portcullis_junc.h)
TARGET_STATS_DTYPE = np.dtype([("spliced", "<u8"), ("unspliced", "<u8"), ("sumq", "<u8"), ("minq", "<i4"), ("maxq", "<i4")])


def merge_target_stats(parts):
    """Per-target scalars of several parts -> one array (counts and sums added, min / max combined)."""
    out = parts[0].copy()
    for p in parts[1:]:
        out["spliced"] += p["spliced"]; out["unspliced"] += p["unspliced"]; out["sumq"] += p["sumq"]
        out["minq"] = np.minimum(out["minq"], p["minq"]); out["maxq"] = np.maximum(out["maxq"], p["maxq"])
    return out


def finalize(rows, mean_query_length):
    """A12/A13 on the host (pj_junctions_finalize)."""
    lib = L.load()
    rows = np.ascontiguousarray(rows)
    _check(lib.pj_junctions_finalize(rows.ctypes.data, len(rows), float(mean_query_length)), lambda: b"finalize failed")
    return rows


def coordinate_order(tid, pos, flag, device=0):
    """samtools' coordinate order of records, computed on the GPU (pj_coordinate_order): order[k] = index of the k-th record."""
    lib = L.load()
    t = np.ascontiguousarray(tid, dtype=np.int32); p = np.ascontiguousarray(pos, dtype=np.int32); f = np.ascontiguousarray(flag, dtype=np.uint16)
    order = np.zeros(len(t), dtype=np.uint32)
    _check(lib.pj_coordinate_order(int(device), len(t), t.ctypes.data, p.ctypes.data, f.ctypes.data, order.ctypes.data), lib.pj_global_last_error)
    return order


class Prepare:
    """Mirror of portcullis::Prepare (src/prepare.hpp:148-216): ``Prepare(output_dir).prepare(bam_files, genome_file)``."""

    def __init__(self, output_dir="portcullis_prep"):
        self.output_dir = output_dir
        self.force = False
        self.use_links = True
        self.use_csi = False
        self.threads = 1
        self.verbose = False
        self.device = 0
        self.report = None

    def setForce(self, b): self.force = bool(b)
    def setUseLinks(self, b): self.use_links = bool(b)
    def setUseCsi(self, b): self.use_csi = bool(b)
    def setThreads(self, n): self.threads = int(n)
    def setVerbose(self, b): self.verbose = bool(b)

    def prepare(self, bam_files, genome_file):
        lib = L.load()
        o = L.PjhPrepOptions()
        lib.pjh_prep_options_default(C.byref(o))
        bams = [os.fsencode(b) for b in bam_files]
        arr = (C.c_char_p * len(bams))(*bams)
        keep = [os.fsencode(genome_file), os.fsencode(self.output_dir), arr]
        o.genome_file, o.output_dir = keep[0], keep[1]
        o.bam_files, o.n_bam_files = arr, len(bams)
        o.force, o.copy, o.use_csi, o.threads = int(self.force), int(not self.use_links), int(self.use_csi), self.threads
        o.verbose, o.quiet, o.device = int(self.verbose), 1, self.device
        rep = L.PjhPrepReport()
        _check(lib.pjh_prep_run(C.byref(o), C.byref(rep)), lib.pjh_prep_last_error)
        self.report = {f: getattr(rep, f) for f, _ in L.PjhPrepReport._fields_}
        return self.report


class BamFilter:
    """Mirror of portcullis::BamFilter (src/bam_filter.hpp): ``BamFilter(junction_file, bam_file, output_bam).filter()``."""

    def __init__(self, junction_file, bam_file, output_bam="filtered.bam"):
        self.junction_file, self.bam_file, self.output_bam = junction_file, bam_file, output_bam
        self.clip_mode = "HARD"
        self.save_msrs = False
        self.use_csi = False
        self.threads = 1
        self.device = 0
        self.report = None

    def setClipMode(self, m): self.clip_mode = m.upper()
    def setSaveMSRs(self, b): self.save_msrs = bool(b)
    def setUseCsi(self, b): self.use_csi = bool(b)
    def setThreads(self, n): self.threads = int(n)

    def filter(self):
        lib = L.load()
        o = L.PjhBamfiltOptions()
        lib.pjh_bamfilt_options_default(C.byref(o))
        keep = [os.fsencode(self.junction_file), os.fsencode(self.bam_file), os.fsencode(self.output_bam)]
        o.junction_file, o.bam_file, o.output_bam = keep
        o.clip_mode, o.save_msrs, o.use_csi = L.CLIP_MODE[self.clip_mode], int(self.save_msrs), int(self.use_csi)
        o.threads, o.quiet, o.device = self.threads, 1, self.device
        rep = L.PjhBamfiltReport()
        _check(lib.pjh_bamfilt_run(C.byref(o), C.byref(rep)), lib.pjh_bamfilt_last_error)
        self.report = {f: getattr(rep, f) for f, _ in L.PjhBamfiltReport._fields_}
        return self.report


class JunctionSet:
    """A junction set on the GPU (pj_jset): ``keep, n_nops = JunctionSet(tid, start, end).filter(cols)``."""

    def __init__(self, tid, start, end, device=0):
        self._lib = L.load()
        order = np.lexsort((end, start, tid))
        t, s, e = (np.ascontiguousarray(np.asarray(a)[order], dtype=np.int32) for a in (tid, start, end))
        self._s = C.c_void_p()
        _check(self._lib.pj_jset_create(int(device), len(t), t.ctypes.data, s.ctypes.data, e.ctypes.data, C.byref(self._s)), self._lib.pj_global_last_error)

    def filter(self, cols):
        n = len(cols["pos"])
        t = np.ascontiguousarray(cols["tid"], dtype=np.int32); p = np.ascontiguousarray(cols["pos"], dtype=np.int32)
        co = np.ascontiguousarray(cols["cigar_off"], dtype=np.uint32); cg = np.ascontiguousarray(cols["cigar"], dtype=np.uint32)
        keep = np.zeros(n, dtype=np.uint8); nn = np.zeros(n, dtype=np.uint8)
        _check(self._lib.pj_jset_filter(self._s, n, t.ctypes.data, p.ctypes.data, co.ctypes.data, cg.ctypes.data if cg.size else None,
                                        keep.ctypes.data, nn.ctypes.data), self._lib.pj_global_last_error)
        return keep, nn

    def close(self):
        if self._s:
            self._lib.pj_jset_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def separate_bams(prep_dir, output_prefix, use_csi=False, threads=1):
    """`junc --separate` on its own (host only): writes <prefix>.spliced/.unspliced/.unmapped.bam and the two indices.
    Returns (n_spliced, n_unspliced, n_unmapped)."""
    lib = L.load()
    counts = np.zeros(3, dtype=np.uint64)
    _check(lib.pjh_separate_bams(os.fsencode(prep_dir), os.fsencode(output_prefix), int(use_csi), int(threads), counts.ctypes.data),
           lib.pjh_last_error)
    return tuple(int(x) for x in counts)


def coverage_source(covered):
    """Q14: which target's depth vector the reference applies to the junctions of each target (-1: none)."""
    lib = L.load()
    cov = np.ascontiguousarray(covered, dtype=np.uint8)
    src = np.zeros(len(cov), dtype=np.int32)
    lib.pj_extra_coverage_source(len(cov), cov.ctypes.data, src.ctypes.data)
    return src


def extra_finalize(x):
    lib = L.load()
    x = np.ascontiguousarray(x)
    lib.pj_extra_finalize(x.ctypes.data, len(x))
    return x


def write_outputs(prefix, rows, names, lengths, source="portcullis", version="1.2.4", exon_gff=False, intron_gff=False,
                  extra=None):
    """tab / bed / gff writers; extra: EXTRA_DTYPE array aligned with rows (the `--extra` columns) or None."""
    lib = L.load()
    rows = np.ascontiguousarray(rows)
    if extra is not None:
        extra = np.ascontiguousarray(extra, dtype=L.EXTRA_DTYPE)
        if len(extra) != len(rows):
            raise ValueError("extra must have one entry per row")
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        tl = np.ascontiguousarray(lengths, dtype=np.int32)
        _check(lib.pjh_write_outputs_extra(os.fsencode(prefix), rows.ctypes.data, extra.ctypes.data, len(rows), len(names), arr,
                                           tl.ctypes.data, source.encode(), version.encode(), int(exon_gff), int(intron_gff)),
               lib.pjh_last_error)
        return
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    tl = np.ascontiguousarray(lengths, dtype=np.int32)
    _check(lib.pjh_write_outputs(os.fsencode(prefix), rows.ctypes.data, len(rows), len(names), arr, tl.ctypes.data,
                                 source.encode(), version.encode(), int(exon_gff), int(intron_gff)), lib.pjh_last_error)


class JunctionBuilder:
    """Mirror of portcullis::JunctionBuilder: ``JunctionBuilder(prep_dir, output).process()``."""

    def __init__(self, prep_dir, output="portcullis_junc/portcullis"):
        self.prep_dir = prep_dir
        self.output = output
        self.threads = 1
        self.gpus = 1
        self.gpu_ids = None          # optional explicit device ordinals (len == gpus); default 0..gpus-1
        self.extra = False
        self.separate = False
        self.use_csi = False
        self.strand_specific = "UNKNOWN"
        self.orientation = "UNKNOWN"
        self.source = "portcullis"
        self.output_exon_gff = False
        self.output_intron_gff = False
        self.verbose = False
        self.quiet = True
        self.report = None

    # setters named after the reference's (junction_builder.hpp)
    def setThreads(self, n): self.threads = int(n)
    def setGpus(self, n): self.gpus = int(n)
    def setExtra(self, b): self.extra = bool(b)
    def setSeparate(self, b): self.separate = bool(b)
    def setSource(self, s): self.source = s
    def setStrandSpecific(self, s): self.strand_specific = s.upper()
    def setOrientation(self, s): self.orientation = s.upper()
    def setUseCsi(self, b): self.use_csi = bool(b)
    def setOutputExonGFF(self, b): self.output_exon_gff = bool(b)
    def setOutputIntronGFF(self, b): self.output_intron_gff = bool(b)
    def setVerbose(self, b): self.verbose = bool(b)

    def _options(self):
        lib = L.load()
        o = L.PjhOptions()
        lib.pjh_options_default(C.byref(o))
        keep = [os.fsencode(self.prep_dir), os.fsencode(self.output), self.source.encode()]
        o.prep_dir, o.output_prefix, o.source = keep
        o.threads, o.n_gpus = self.threads, self.gpus
        if self.gpu_ids is not None:
            ids = (C.c_int32 * len(self.gpu_ids))(*self.gpu_ids)
            keep.append(ids)
            o.gpu_ids = ids
        o.orientation = L.ORIENT[self.orientation]
        o.strandedness = L.STRANDEDNESS[self.strand_specific]
        o.use_csi, o.exon_gff, o.intron_gff = int(self.use_csi), int(self.output_exon_gff), int(self.output_intron_gff)
        o.verbose, o.separate, o.extra, o.quiet = int(self.verbose), int(self.separate), int(self.extra), int(self.quiet)
        return lib, o, keep

    def process_part(self, part, n_parts, device=0):
        """One process per GPU (torchrun): decode and run part `part` of the n_parts-way work plan on CUDA device `device`.
        Returns (rows, stats, report): rows in (tid, start, end) order, not yet finalized; stats one entry per target."""
        saved = self.gpu_ids
        self.gpu_ids = [int(device)]
        try:
            lib, o, keep = self._options()
        finally:
            self.gpu_ids = saved
        h = C.c_void_p()
        rep = L.PjhReport()
        _check(lib.pjh_junc_run_part(C.byref(o), int(part), int(n_parts), C.byref(h), C.byref(rep)), lib.pjh_last_error)
        try:
            pr, ps = C.c_void_p(), C.c_void_p()
            n = lib.pjh_partial_rows(h, C.byref(pr))
            t = lib.pjh_partial_stats(h, C.byref(ps))
            rows = np.frombuffer(C.string_at(pr, n * L.JUNCTION_DTYPE.itemsize), dtype=L.JUNCTION_DTYPE).copy() if n else np.zeros(0, dtype=L.JUNCTION_DTYPE)
            stats = np.frombuffer(C.string_at(ps, t * C.sizeof(L.PjTargetStats)), dtype=TARGET_STATS_DTYPE).copy()
        finally:
            lib.pjh_partial_free(h)
        return rows, stats, {f: getattr(rep, f) for f, _ in L.PjhReport._fields_}

    def finish(self, rows, stats):
        """Rows of all parts concatenated in part order + merged per-target stats -> A12/A13, output files, report."""
        lib, o, keep = self._options()
        rows = np.ascontiguousarray(rows, dtype=L.JUNCTION_DTYPE)
        stats = np.ascontiguousarray(stats, dtype=TARGET_STATS_DTYPE)
        rep = L.PjhReport()
        _check(lib.pjh_junc_finish(C.byref(o), rows.ctypes.data, len(rows), stats.ctypes.data, len(stats), C.byref(rep)), lib.pjh_last_error)
        self.report = {f: getattr(rep, f) for f, _ in L.PjhReport._fields_}
        return rows, self.report

    def process(self):
        lib = L.load()
        o = L.PjhOptions()
        lib.pjh_options_default(C.byref(o))
        keep = [os.fsencode(self.prep_dir), os.fsencode(self.output), self.source.encode()]
        o.prep_dir, o.output_prefix, o.source = keep
        o.threads, o.n_gpus = self.threads, self.gpus
        if self.gpu_ids is not None:
            if len(self.gpu_ids) != self.gpus:
                raise ValueError("gpu_ids must name one device per GPU")
            ids = (C.c_int32 * self.gpus)(*self.gpu_ids)
            keep.append(ids)
            o.gpu_ids = ids
        o.orientation = L.ORIENT[self.orientation]
        o.strandedness = L.STRANDEDNESS[self.strand_specific]
        o.use_csi, o.exon_gff, o.intron_gff = int(self.use_csi), int(self.output_exon_gff), int(self.output_intron_gff)
        o.verbose, o.separate, o.extra, o.quiet = int(self.verbose), int(self.separate), int(self.extra), int(self.quiet)
        rep = L.PjhReport()
        _check(lib.pjh_junc_run(C.byref(o), C.byref(rep)), lib.pjh_last_error)
        self.report = {f: getattr(rep, f) for f, _ in L.PjhReport._fields_}
        return self.report
