"""ctypes binding of the B200 junc library (include/portcullis_junc.h, include/portcullis_junc_host.h).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``make -C portcullis_b200/csrc``.
There is no Python or CPU fallback: if the library is missing this module raises, and every compute
entry point fails with PJ_ECUDA when no B200 is visible.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libportcullis_junc.so")

PJ_OK, PJ_EINVAL, PJ_ECUDA, PJ_ENOMEM, PJ_ESTATE, PJ_EDATA, PJ_EIO = 0, -1, -2, -3, -4, -5, -6
ORIENT = {"SE": 0, "FR": 1, "RF": 2, "FF": 3, "UNKNOWN": 4}
STRANDEDNESS = {"UNSTRANDED": 0, "FIRSTSTRAND": 1, "SECONDSTRAND": 2, "UNKNOWN": 3}
NB_JAD = 20


class PjError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("[pj %d] %s" % (code, msg))
        self.code = code


class PjConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("orientation", C.c_int32), ("reserved", C.c_int32 * 6),
                ("extra_metrics", C.c_int32), ("pad", C.c_int32)]


class PjBatch(C.Structure):
    _fields_ = [("n_records", C.c_int64), ("tid", C.c_void_p), ("pos", C.c_void_p), ("flag", C.c_void_p),
                ("mapq", C.c_void_p), ("xs", C.c_void_p), ("l_qseq", C.c_void_p), ("mtid", C.c_void_p),
                ("mpos", C.c_void_p), ("cigar_off", C.c_void_p), ("cigar", C.c_void_p), ("seq_off", C.c_void_p),
                ("seq4", C.c_void_p), ("name_code", C.c_void_p),
                # lean form (ABI 2)
                ("lean", C.c_int32), ("const_tid", C.c_int32), ("n_cigar", C.c_void_p), ("n_cigar_total", C.c_int64),
                ("seq2", C.c_void_p), ("n_seq2_bytes", C.c_int64), ("seqx_pos", C.c_void_p), ("seqx_code", C.c_void_p),
                ("n_seqx", C.c_int64)]


class PjhLeanRun(C.Structure):
    _fields_ = [("tid", C.c_int32), ("pad", C.c_int32), ("rec0", C.c_int64), ("cig0", C.c_int64), ("seq0", C.c_int64), ("seqx0", C.c_int64)]


class PjTargetStats(C.Structure):
    _fields_ = [("spliced_count", C.c_uint64), ("unspliced_count", C.c_uint64), ("sum_query_lengths", C.c_uint64),
                ("min_query_length", C.c_int32), ("max_query_length", C.c_int32)]


# numpy mirror of struct pj_junction (offsets follow the C layout; checked against sizeof in tests)
JUNCTION_DTYPE = np.dtype([
    ("tid", "<i4"), ("start", "<i4"), ("end", "<i4"), ("left", "<i4"), ("right", "<i4"),
    ("nb_raw_aln", "<u4"), ("nb_dist_aln", "<u4"), ("nb_ms_aln", "<u4"), ("nb_um_aln", "<u4"), ("nb_bpp_aln", "<u4"),
    ("nb_ppp_aln", "<u4"), ("nb_rel_aln", "<u4"), ("nb_r1_pos", "<u4"), ("nb_r1_neg", "<u4"), ("nb_r2_pos", "<u4"),
    ("nb_r2_neg", "<u4"), ("nb_xs_pos", "<u4"), ("nb_xs_neg", "<u4"), ("max_min_anc", "<u4"), ("maxmmes", "<u4"),
    ("nb_mismatches", "<u4"), ("hamming5p", "<u4"), ("hamming3p", "<u4"), ("nb_up_juncs", "<u4"), ("nb_down_juncs", "<u4"),
    ("jad", "<u4", (NB_JAD,)), ("pad_a", "<u4"), ("entropy", "<f8"),
    ("read_strand", "u1"), ("ss_strand", "u1"), ("consensus_strand", "u1"), ("canonical_ss", "u1"), ("suspicious", "u1"),
    ("ss1", "S2"), ("ss2", "S2"), ("pad0", "u1", (7,)),
    ("index", "<u4"), ("dist_2_up_junc", "<u4"), ("dist_2_down_junc", "<u4"), ("dist_nearest_junc", "<u4"),
    ("uniq_junc", "u1"), ("primary_junc", "u1"), ("pfp", "u1"), ("pad1", "u1", (5,)),
    ("mean_readlen", "<f8"), ("rel2raw", "<f8"), ("mean_mismatches", "<f8"),
], align=False)

# numpy mirror of struct pj_junction_extra (48 bytes)
EXTRA_DTYPE = np.dtype([("up_aln", "<u4"), ("down_aln", "<u4"), ("mm_n", "<u4"), ("mm_m", "<u4"), ("cov_sum", "<u4", (4,)),
                        ("mm_score", "<f8"), ("coverage", "<f8")], align=False)

# fields produced on the GPU (integer, string and enum columns: bit-exact parity; entropy within 1e-6 relative)
DEVICE_INT_FIELDS = ["tid", "start", "end", "left", "right", "nb_raw_aln", "nb_dist_aln", "nb_ms_aln", "nb_um_aln",
                     "nb_bpp_aln", "nb_ppp_aln", "nb_rel_aln", "nb_r1_pos", "nb_r1_neg", "nb_r2_pos", "nb_r2_neg",
                     "nb_xs_pos", "nb_xs_neg", "max_min_anc", "maxmmes", "nb_mismatches", "hamming5p", "hamming3p",
                     "nb_up_juncs", "nb_down_juncs", "jad", "read_strand", "ss_strand", "consensus_strand",
                     "canonical_ss", "suspicious", "ss1", "ss2"]
FINALIZE_FIELDS = ["index", "dist_2_up_junc", "dist_2_down_junc", "dist_nearest_junc", "uniq_junc", "primary_junc",
                   "pfp", "mean_readlen", "rel2raw", "mean_mismatches"]


class PjhOptions(C.Structure):
    _fields_ = [("prep_dir", C.c_char_p), ("output_prefix", C.c_char_p), ("threads", C.c_int32), ("n_gpus", C.c_int32),
                ("gpu_ids", C.POINTER(C.c_int32)), ("orientation", C.c_int32), ("strandedness", C.c_int32),
                ("use_csi", C.c_int32), ("exon_gff", C.c_int32), ("intron_gff", C.c_int32), ("source", C.c_char_p),
                ("verbose", C.c_int32), ("separate", C.c_int32), ("extra", C.c_int32), ("quiet", C.c_int32),
                ("version", C.c_char_p)]


class PjhReport(C.Structure):
    _fields_ = [("n_junctions", C.c_int64), ("n_spliced", C.c_uint64), ("n_unspliced", C.c_uint64),
                ("mean_query_length", C.c_double), ("min_query_length", C.c_int32), ("max_query_length", C.c_int32),
                ("t_open_s", C.c_double), ("t_genome_s", C.c_double), ("t_decode_s", C.c_double), ("t_gpu_ms", C.c_double),
                ("t_finalize_s", C.c_double), ("t_write_s", C.c_double), ("t_total_s", C.c_double),
                ("n_gpus_used", C.c_int32), ("n_kernel_launches", C.c_int32),
                ("t_init_s", C.c_double), ("t_run_s", C.c_double), ("t_teardown_s", C.c_double),
                ("t_extra_s", C.c_double), ("t_separate_s", C.c_double), ("n_segments", C.c_int32), ("n_gap_cuts", C.c_int32)]


class PjhPrepOptions(C.Structure):
    _fields_ = [("genome_file", C.c_char_p), ("bam_files", C.POINTER(C.c_char_p)), ("n_bam_files", C.c_int32), ("output_dir", C.c_char_p),
                ("force", C.c_int32), ("copy", C.c_int32), ("use_csi", C.c_int32), ("threads", C.c_int32), ("verbose", C.c_int32),
                ("quiet", C.c_int32), ("device", C.c_int32)]


class PjhPrepReport(C.Structure):
    _fields_ = [("n_records", C.c_int64), ("sorted_in_process", C.c_int32), ("t_sort_s", C.c_double), ("t_sort_gpu_s", C.c_double),
                ("t_total_s", C.c_double)]


class PjhBamfiltOptions(C.Structure):
    _fields_ = [("junction_file", C.c_char_p), ("bam_file", C.c_char_p), ("output_bam", C.c_char_p), ("clip_mode", C.c_int32),
                ("save_msrs", C.c_int32), ("use_csi", C.c_int32), ("threads", C.c_int32), ("verbose", C.c_int32), ("quiet", C.c_int32),
                ("device", C.c_int32)]


class PjhBamfiltReport(C.Structure):
    _fields_ = [("n_junctions", C.c_int64), ("n_in", C.c_uint64), ("n_out", C.c_uint64), ("n_modified", C.c_uint64),
                ("t_device_s", C.c_double), ("t_total_s", C.c_double)]


CLIP_MODE = {"HARD": 0, "SOFT": 1, "COMPLETE": 2}

# every symbol declared in include/*.h, with (restype, argtypes); used by load() and by the export test
_P = C.c_void_p
SYMBOLS = {
    "pj_abi_version": (C.c_int, []),
    "pj_junction_size": (C.c_int, []),
    "pj_global_last_error": (C.c_char_p, []),
    "pj_create": (C.c_int, [C.POINTER(PjConfig), C.POINTER(_P)]),
    "pj_destroy": (None, [_P]),
    "pj_last_error": (C.c_char_p, [_P]),
    "pj_targets_set": (C.c_int, [_P, C.c_int32, _P]),
    "pj_genome_set_target": (C.c_int, [_P, C.c_int32, _P, C.c_int64]),
    "pj_genome_load_fasta": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(C.c_char_p)]),
    "pj_shard_begin": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64]),
    "pj_staging_acquire": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.POINTER(PjBatch)]),
    "pj_staging_acquire_lean": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(PjBatch)]),
    "pj_batch_submit": (C.c_int, [_P, C.POINTER(PjBatch)]),
    "pj_shard_run": (C.c_int, [_P]),
    "pj_shard_num_junctions": (C.c_int64, [_P]),
    "pj_shard_fetch": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int32]),
    "pj_shard_timing": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "pj_shard_kernel_times": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_char_p), C.POINTER(C.c_int32)]),
    "pj_junctions_finalize": (C.c_int, [_P, C.c_int64, C.c_double]),
    "pj_jset_create": (C.c_int, [C.c_int32, C.c_int64, _P, _P, _P, C.POINTER(_P)]),
    "pj_jset_destroy": (None, [_P]),
    "pj_jset_filter": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P]),
    "pjh_bamfilt_options_default": (None, [C.POINTER(PjhBamfiltOptions)]),
    "pjh_bamfilt_run": (C.c_int, [C.POINTER(PjhBamfiltOptions), C.POINTER(PjhBamfiltReport)]),
    "pjh_bamfilt_last_error": (C.c_char_p, []),
    "pjh_bamfilt_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "pj_coordinate_order": (C.c_int, [C.c_int32, C.c_int64, _P, _P, _P, _P]),
    "pjh_prep_options_default": (None, [C.POINTER(PjhPrepOptions)]),
    "pjh_prep_run": (C.c_int, [C.POINTER(PjhPrepOptions), C.POINTER(PjhPrepReport)]),
    "pjh_prep_last_error": (C.c_char_p, []),
    "pjh_prep_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "pj_extra_num_spliced_names": (C.c_int64, [_P]),
    "pj_extra_export_names": (C.c_int, [_P, _P, C.c_int64]),
    "pj_extra_import_names": (C.c_int, [_P, _P, C.c_int64]),
    "pj_extra_run": (C.c_int, [_P, C.c_int32, _P, C.c_int64]),
    "pj_extra_timing": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "pj_extra_kernel_times": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_char_p), C.POINTER(C.c_int32)]),
    "pj_extra_target_pileup": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]),
    "pj_extra_coverage": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P, _P]),
    "pj_extra_coverage_batch": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P]),
    "pj_extra_coverage_source": (None, [C.c_int32, _P, _P]),
    "pj_extra_finalize": (None, [_P, C.c_int64]),
    "pj_features_create": (C.c_int, [_P, C.POINTER(_P)]),
    "pj_features_destroy": (None, [_P]),
    "pj_features_train_coding": (C.c_int, [_P, _P, C.c_int64, _P]),
    "pj_features_train_splicing": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "pj_features_intron_threshold": (C.c_uint32, [_P, C.c_int64, _P]),
    "pj_features_run": (C.c_int, [_P, _P, C.c_int64, C.c_uint32, _P, C.POINTER(C.c_float)]),
    "pjh_options_default": (None, [C.POINTER(PjhOptions)]),
    "pjh_junc_run": (C.c_int, [C.POINTER(PjhOptions), C.POINTER(PjhReport)]),
    "pjh_junc_run_part": (C.c_int, [C.POINTER(PjhOptions), C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(PjhReport)]),
    "pjh_partial_rows": (C.c_int64, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "pjh_partial_stats": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "pjh_partial_free": (None, [C.c_void_p]),
    "pjh_junc_finish": (C.c_int, [C.POINTER(PjhOptions), C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.POINTER(PjhReport)]),
    "pjh_last_error": (C.c_char_p, []),
    "pjh_junc_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "pjh_prep_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "pjh_prep_close": (None, [_P]),
    "pjh_prep_n_targets": (C.c_int32, [_P]),
    "pjh_prep_target_name": (C.c_char_p, [_P, C.c_int32]),
    "pjh_prep_target_len": (C.c_int32, [_P, C.c_int32]),
    "pjh_prep_target_records": (C.c_int64, [_P, C.c_int32]),
    "pjh_prep_decode": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(PjBatch)]),
    "pjh_prep_want_names": (None, [_P, C.c_int32]),
    "pjh_prep_genome": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int64)]),
    "pjh_inflate_selftest": (C.c_int, [C.c_int32]),
    "pjh_format_selftest": (C.c_int, [C.c_int32]),
    "pjh_plan_shards": (C.c_int, [_P, C.c_int32, _P]),
    "pjh_plan_describe": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int64, _P, C.POINTER(C.c_int32)]),
    "pjh_plan_decode_lean": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(PjBatch),
                                       C.POINTER(C.POINTER(PjhLeanRun)), C.POINTER(C.c_int32)]),
    "pjh_plan_decode": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(PjBatch)]),
    "pjh_separate_bams": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, _P]),
    "pjh_write_outputs_extra": (C.c_int, [C.c_char_p, _P, _P, C.c_int64, C.c_int32, C.POINTER(C.c_char_p), _P, C.c_char_p, C.c_char_p,
                                          C.c_int32, C.c_int32]),
    "pjh_write_outputs": (C.c_int, [C.c_char_p, _P, C.c_int64, C.c_int32, C.POINTER(C.c_char_p), _P, C.c_char_p, C.c_char_p,
                                    C.c_int32, C.c_int32]),
}

_lib = None


def load():
    """Load the in-tree shared library (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` or "
                              "`make -C portcullis_b200/csrc` (this package has no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.pj_junction_size() != JUNCTION_DTYPE.itemsize:
            raise ImportError("pj_junction layout mismatch between the library and portcullis_b200/_lib.py")
        _lib = lib
    return _lib
