"""ctypes binding of oracle/movi_oracle.c (the CPU restatement of the reference algorithm).

TEST INFRASTRUCTURE ONLY -- see oracle/movi_oracle.h.  Builds the shared object on first use
(`make -C oracle port`; gcc only, no GPU).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libmovi_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "movi_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.movi_oracle_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.movi_oracle_load.restype = C.c_int
        L.movi_oracle_free.argtypes = [C.c_void_p]
        L.movi_oracle_last_error.restype = C.c_char_p
        for name in ("mode", "sigma", "has_separators"):
            f = getattr(L, "movi_oracle_" + name); f.argtypes = [C.c_void_p]; f.restype = C.c_int
        for name in ("r", "n", "end_bwt_idx"):
            f = getattr(L, "movi_oracle_" + name); f.argtypes = [C.c_void_p]; f.restype = C.c_uint64
        for name in ("get_n", "get_offset", "get_id"):
            f = getattr(L, "movi_oracle_" + name); f.argtypes = [C.c_void_p, C.c_uint64]; f.restype = C.c_uint64
        L.movi_oracle_get_c.argtypes = [C.c_void_p, C.c_uint64]; L.movi_oracle_get_c.restype = C.c_int
        L.movi_oracle_get_threshold_bit.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.movi_oracle_get_threshold_bit.restype = C.c_int
        L.movi_oracle_lf_move.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.movi_oracle_lf_move.restype = C.c_int64
        L.movi_oracle_query_ml_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        L.movi_oracle_query_ml_batch.restype = C.c_int
        L.movi_oracle_query_count_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        L.movi_oracle_query_count_batch.restype = C.c_int
        L.movi_oracle_classify.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64] + [C.POINTER(C.c_uint64)] * 4
        L.movi_oracle_classify.restype = C.c_int
        L.movi_oracle_colors_load.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p)]; L.movi_oracle_colors_load.restype = C.c_int
        L.movi_oracle_colors_free.argtypes = [C.c_void_p]
        L.movi_oracle_colors_num_species.argtypes = [C.c_void_p]; L.movi_oracle_colors_num_species.restype = C.c_uint32
        L.movi_oracle_colors_taxon.argtypes = [C.c_void_p, C.c_uint32]; L.movi_oracle_colors_taxon.restype = C.c_uint32
        L.movi_oracle_query_multi_classify.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                                       C.c_void_p, C.c_void_p, C.c_void_p]
        L.movi_oracle_query_multi_classify.restype = C.c_int
        L.movi_oracle_ssa_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]; L.movi_oracle_ssa_load.restype = C.c_int
        L.movi_oracle_ssa_free.argtypes = [C.c_void_p]
        L.movi_oracle_ssa_build.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]; L.movi_oracle_ssa_build.restype = C.c_int
        for name in ("rate", "samples", "rows"):
            f = getattr(L, "movi_oracle_ssa_" + name); f.argtypes = [C.c_void_p]; f.restype = C.c_uint64
        for name in ("sampled", "all_p"):
            f = getattr(L, "movi_oracle_ssa_" + name); f.argtypes = [C.c_void_p]; f.restype = C.c_void_p
        L.movi_oracle_query_pml_sa.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.movi_oracle_query_pml_sa.restype = C.c_int
        _lib = L
    return _lib


def pack_reads(reads):
    """list of bytes/str -> (bases uint8[total], offsets uint64[n+1])."""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return bases, offsets


class OracleIndex:
    """The reference algorithm on one index.movi, on the CPU."""

    def __init__(self, index_file):
        if os.path.isdir(index_file):
            index_file = os.path.join(index_file, "index.movi")
        self._h = C.c_void_p()
        if lib().movi_oracle_load(index_file.encode(), C.byref(self._h)) != 0:
            raise RuntimeError("oracle: " + lib().movi_oracle_last_error().decode())
        L = lib()
        self.mode = L.movi_oracle_mode(self._h)
        self.r = L.movi_oracle_r(self._h)
        self.n = L.movi_oracle_n(self._h)
        self.end_bwt_idx = L.movi_oracle_end_bwt_idx(self._h)
        self.sigma = L.movi_oracle_sigma(self._h)
        self.separators = bool(L.movi_oracle_has_separators(self._h))
        self.index_dir = os.path.dirname(index_file)
        self._colors = C.c_void_p()
        self._ssa = C.c_void_p()

    def __del__(self):
        if getattr(self, "_colors", None):
            lib().movi_oracle_colors_free(self._colors)
            self._colors = None
        if getattr(self, "_ssa", None):
            lib().movi_oracle_ssa_free(self._ssa)
            self._ssa = None
        if getattr(self, "_h", None):
            lib().movi_oracle_free(self._h)
            self._h = None

    def row(self, idx):
        L = lib()
        return dict(n=L.movi_oracle_get_n(self._h, idx), offset=L.movi_oracle_get_offset(self._h, idx),
                    id=L.movi_oracle_get_id(self._h, idx), c=L.movi_oracle_get_c(self._h, idx),
                    thr=tuple(L.movi_oracle_get_threshold_bit(self._h, idx, k) for k in range(3)))

    def lf_move(self, offset, idx):
        o, i = C.c_uint64(offset), C.c_uint64(idx)
        ff = lib().movi_oracle_lf_move(self._h, C.byref(o), C.byref(i))
        if ff < 0:
            raise RuntimeError("oracle: " + lib().movi_oracle_last_error().decode())
        return o.value, i.value, ff

    def _ml(self, kind, bases, offsets, threads):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = np.zeros(int(offsets[-1]), dtype=np.uint16)
        rc = lib().movi_oracle_query_ml_batch(self._h, kind, bases.ctypes.data, offsets.ctypes.data,
                                              len(offsets) - 1, out.ctypes.data, threads)
        if rc != 0:
            raise RuntimeError("oracle query failed: " + lib().movi_oracle_last_error().decode())
        return out

    def query_pml(self, bases, offsets, threads=1):
        """PMLs in the reference's storage order (index 0 of each read = its LAST base)."""
        return self._ml(0, bases, offsets, threads)

    def query_zml(self, bases, offsets, threads=1):
        return self._ml(1, bases, offsets, threads)

    def query_rpml(self, bases, offsets, threads=1):
        """--rpml: PML with the offset-based repositioning instead of the thresholds."""
        return self._ml(2, bases, offsets, threads)

    def query_count(self, bases, offsets, threads=1):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        counts = np.zeros(n, dtype=np.uint64)
        pos = np.zeros(n, dtype=np.int32)
        rc = lib().movi_oracle_query_count_batch(self._h, bases.ctypes.data, offsets.ctypes.data, n,
                                                 counts.ctypes.data, pos.ctypes.data, threads)
        if rc != 0:
            raise RuntimeError("oracle count failed: " + lib().movi_oracle_last_error().decode())
        return counts, pos

    # ---- Movi Color (reference `--multi-classify` / `--sa-entries`, --no-prefetch path)
    def load_colors(self):
        """doc_sets_flat.bin + ref.fa.doc_offsets [+ ref.fa.doc_ids] of the index directory."""
        if not self._colors:
            if lib().movi_oracle_colors_load(self.index_dir.encode(), self.r, C.byref(self._colors)) != 0:
                raise RuntimeError("oracle: " + lib().movi_oracle_last_error().decode())
        return lib().movi_oracle_colors_num_species(self._colors)

    def taxon(self, species):
        self.load_colors()
        return lib().movi_oracle_colors_taxon(self._colors, species)

    def query_multi_classify(self, bases, offsets, min_len=1, pvalue=False):
        """-> (taxa uint32[n, 2], raw uint32[n, 3] = best doc / second doc / sum of lengths, score float64[n, 2], ml uint16[total])."""
        self.load_colors()
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        taxa = np.zeros((n, 2), np.uint32); raw = np.zeros((n, 3), np.uint32); score = np.zeros((n, 2), np.float64)
        ml = np.zeros(int(offsets[-1]) + 1, dtype=np.uint16)
        for i in range(n):
            s, e = int(offsets[i]), int(offsets[i + 1])
            rc = lib().movi_oracle_query_multi_classify(self._h, self._colors, bases.ctypes.data + s, e - s, int(min_len), int(bool(pvalue)),
                                                        ml.ctypes.data + 2 * s, raw[i].ctypes.data, score[i].ctypes.data, taxa[i].ctypes.data)
            if rc != 0:
                raise RuntimeError("oracle multi-classify failed: " + lib().movi_oracle_last_error().decode())
        return taxa, raw, score, ml[:-1]

    def build_ssa(self, rate=100):
        """The reference's `build-SA` action on this index -> (rate, sampled uint64[], all_p uint64[]); later query_pml_sa calls use it."""
        L = lib()
        if self._ssa:
            L.movi_oracle_ssa_free(self._ssa); self._ssa = C.c_void_p()
        if L.movi_oracle_ssa_build(self._h, rate, C.byref(self._ssa)) != 0:
            raise RuntimeError("oracle: " + L.movi_oracle_last_error().decode())
        k, r = L.movi_oracle_ssa_samples(self._ssa), L.movi_oracle_ssa_rows(self._ssa)
        sampled = np.ctypeslib.as_array(C.cast(L.movi_oracle_ssa_sampled(self._ssa), C.POINTER(C.c_uint64)), shape=(k,)).copy()
        all_p = np.ctypeslib.as_array(C.cast(L.movi_oracle_ssa_all_p(self._ssa), C.POINTER(C.c_uint64)), shape=(r,)).copy()
        return L.movi_oracle_ssa_rate(self._ssa), sampled, all_p

    def query_pml_sa(self, bases, offsets):
        """-> (ml uint16[total], sa uint64[total]): `--pml --sa-entries`, same orientation as the lengths."""
        if not self._ssa:
            if lib().movi_oracle_ssa_load(self.index_dir.encode(), C.byref(self._ssa)) != 0:
                raise RuntimeError("oracle: " + lib().movi_oracle_last_error().decode())
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        ml = np.zeros(int(offsets[-1]) + 1, dtype=np.uint16); sa = np.zeros(int(offsets[-1]) + 1, dtype=np.uint64)
        for i in range(n):
            s, e = int(offsets[i]), int(offsets[i + 1])
            if lib().movi_oracle_query_pml_sa(self._h, self._ssa, bases.ctypes.data + s, e - s, ml.ctypes.data + 2 * s, sa.ctypes.data + 8 * s) != 0:
                raise RuntimeError("oracle sa-entries failed: " + lib().movi_oracle_last_error().decode())
        return ml[:-1], sa[:-1]


def classify(ml, bin_width, max_value_thr):
    """Classifier::classify on one read's PMLs -> (found, bins_above, bins_below, sum_max, n_bins)."""
    ml = np.ascontiguousarray(ml, dtype=np.uint16)
    a, b, s, n = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    found = lib().movi_oracle_classify(ml.ctypes.data, len(ml), bin_width, max_value_thr,
                                       C.byref(a), C.byref(b), C.byref(s), C.byref(n))
    return bool(found), a.value, b.value, s.value, n.value
