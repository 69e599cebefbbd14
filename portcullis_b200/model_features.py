"""Python mirror of the reference's filt-stage feature extraction.

``ModelFeatures`` follows portcullis::ml::ModelFeatures (/root/reference/lib/include/portcullis/ml/model_features.hpp:67-126,
lib/src/model_features.cc): ``calcIntronThreshold``, ``trainCodingPotentialModel``, ``trainSplicingModels`` and
``juncs2FeatureVectors`` with the reference's argument meaning, computed by the CUDA library (pj_features_*, csrc/pj_features.cu)
on a context whose genome is resident.  There is no Python fallback."""
import ctypes as C

import numpy as np

from . import _lib as L

# column names of the feature matrix (model_features.hpp:45-60 + Junction::JAD_NAMES)
VAR_NAMES = ["Genuine", "rna_usrs", "rna_dist", "rna_rel", "rna_entropy", "rna_rel2raw", "rna_maxminanc", "rna_maxmmes", "rna_missmatch",
             "rna_intron", "dna_minhamm", "dna_coding", "dna_pws", "dna_ss"] + ["JAD%02d" % k for k in range(1, 21)]
NB_FEATURES = 34


def _mask(m, n):
    if m is None:
        return None
    a = np.ascontiguousarray(np.asarray(m) != 0, dtype=np.uint8)
    if len(a) != n:
        raise ValueError("subset mask must have one entry per junction")
    return a


class ModelFeatures:
    def __init__(self, gpu):
        """gpu: a JuncGpu whose targets are set and whose genome is loaded (set_genome for every target that has junctions)."""
        self._lib = L.load()
        self._gpu = gpu
        self._m = C.c_void_p()
        rc = self._lib.pj_features_create(gpu._ctx, C.byref(self._m))
        if rc:
            raise L.PjError(rc, self._lib.pj_last_error(gpu._ctx).decode())
        self.L95 = 0
        self.device_ms = 0.0

    def _check(self, rc):
        if rc:
            raise L.PjError(rc, self._lib.pj_last_error(self._gpu._ctx).decode())

    def calcIntronThreshold(self, rows, subset=None):
        rows = np.ascontiguousarray(rows, dtype=L.JUNCTION_DTYPE)
        s = _mask(subset, len(rows))
        self.L95 = int(self._lib.pj_features_intron_threshold(rows.ctypes.data, len(rows), s.ctypes.data if s is not None else None))
        return self.L95

    def trainCodingPotentialModel(self, rows, subset=None):
        rows = np.ascontiguousarray(rows, dtype=L.JUNCTION_DTYPE)
        s = _mask(subset, len(rows))
        self._check(self._lib.pj_features_train_coding(self._m, rows.ctypes.data, len(rows), s.ctypes.data if s is not None else None))

    def trainSplicingModels(self, rows, pass_mask, fail_mask):
        rows = np.ascontiguousarray(rows, dtype=L.JUNCTION_DTYPE)
        p, f = _mask(pass_mask, len(rows)), _mask(fail_mask, len(rows))
        self._check(self._lib.pj_features_train_splicing(self._m, rows.ctypes.data, len(rows), p.ctypes.data, f.ctypes.data))

    def juncs2FeatureVectors(self, rows):
        rows = np.ascontiguousarray(rows, dtype=L.JUNCTION_DTYPE)
        out = np.zeros((len(rows), NB_FEATURES), dtype=np.float64)
        ms = C.c_float()
        self._check(self._lib.pj_features_run(self._m, rows.ctypes.data, len(rows), self.L95, out.ctypes.data, C.byref(ms)))
        self.device_ms = ms.value
        return out

    def close(self):
        if self._m:
            self._lib.pj_features_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass




def load_junction_tab(path):
    """Rows (JUNCTION_DTYPE) from a junctions.tab as the reference's filter reads it (JunctionSystem::load /
    Junction::parse, lib/src/junction.cc:1229-1326): the text values, so doubles carry the file's 6 significant digits."""
    rows = []
    with open(path) as f:
        header = f.readline().rstrip("\n").split("\t")
        col = {n: i for i, n in enumerate(header)}
        for line in f:
            c = line.rstrip("\n").split("\t")
            if len(c) < len(header):
                continue
            rows.append(c)
    out = np.zeros(len(rows), dtype=L.JUNCTION_DTYPE)
    strand = {"+": 0, "-": 1, "?": 2}
    for k, c in enumerate(rows):
        r = out[k]
        r["index"] = int(c[col["index"]]); r["tid"] = int(c[col["refid"]]); r["start"] = int(c[col["start"]]); r["end"] = int(c[col["end"]])
        r["left"] = int(c[col["left"]]); r["right"] = int(c[col["right"]])
        r["read_strand"] = strand[c[col["read-strand"]]]; r["ss_strand"] = strand[c[col["ss-strand"]]]; r["consensus_strand"] = strand[c[col["consensus-strand"]]]
        for name in ("nb_raw_aln", "nb_dist_aln", "nb_ms_aln", "nb_um_aln", "nb_bpp_aln", "nb_ppp_aln", "nb_rel_aln", "nb_r1_pos", "nb_r1_neg", "nb_r2_pos",
                     "nb_r2_neg", "max_min_anc", "maxmmes", "hamming5p", "hamming3p", "nb_up_juncs", "nb_down_juncs"):
            r[name] = int(c[col[name]])
        for name in ("entropy", "rel2raw", "mean_mismatches", "mean_readlen"):
            r[name] = float(c[col[name]])
        r["jad"] = [int(c[col["JAD%02d" % q]]) for q in range(1, 21)]
    return out
