"""Shared pieces of the Movi Color tests: the CPU emulation of the device kernel's lanes (tests/native/ext_emulate.cu) and the
host-side loaders / writers of include/movi_host.h."""
import ctypes as C
import os
import subprocess
import threading

import numpy as np

from movi_b200 import _native as N
from movi_b200.api import COLOR_RESULT_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_SRC = os.path.join(ROOT, "tests", "native", "ext_emulate.cu")
EMU_LIB = os.path.join(ROOT, "tests", "_build", "libext_emulate.so")
EXT_COLOR, EXT_PVALUE, EXT_SA = 1, 2, 4


def build_emulator(force=False):
    deps = [EMU_SRC] + [os.path.join(ROOT, "movi_b200", "csrc", f) for f in ("ext_kernel.cuh", "kernel_common.cuh", "dev_index.cuh")]
    if force or not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
        from movi_b200.build import nvcc_path
        subprocess.check_call([nvcc_path(), "-Wno-deprecated-gpu-targets", "-O2", "-std=c++17", "--compiler-options", "-fPIC", "-shared",
                               EMU_SRC, "-o", EMU_LIB])
    return EMU_LIB


_emu = None


def emulator():
    global _emu
    if _emu is None:
        L = C.CDLL(build_emulator())
        L.movi_ext_emulate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                       C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.movi_ext_emulate.restype = C.c_int
        _emu = L
    return _emu


class HostColors:
    def __init__(self, index_dir, r, kind=0):
        H = N.host_lib()
        self._h = C.c_void_p()
        if H.movi_host_colors_open_kind(os.fsencode(index_dir), r, kind, C.byref(self._h)) != 0:
            raise RuntimeError(H.movi_host_last_error().decode())
        self.desc = N.MoviGpuColorDesc()
        H.movi_host_colors_desc(self._h, C.byref(self.desc))
        self.num_species = H.movi_host_colors_num_species(self._h)

    def taxon(self, s):
        return N.host_lib().movi_host_colors_taxon(self._h, s)

    def text(self, reads, res, min_len=1, pvalue=False):
        """movi_host_write_multi_classify into a pipe -> str."""
        H = N.host_lib()
        rfd, wfd = os.pipe()
        buf = []
        t = threading.Thread(target=lambda: buf.append(os.fdopen(rfd, "rb").read()))
        t.start()
        try:
            rc = H.movi_host_write_multi_classify(wfd, reads._h, self._h, res.ctypes.data, int(min_len), int(bool(pvalue)))
        finally:
            os.close(wfd)
            t.join()
        if rc != 0:
            raise RuntimeError(H.movi_host_last_error().decode())
        return buf[0].decode()

    def __del__(self):
        if getattr(self, "_h", None):
            N.host_lib().movi_host_colors_close(self._h)
            self._h = None


class HostSsa:
    def __init__(self, index_dir):
        H = N.host_lib()
        self._h = C.c_void_p()
        if H.movi_host_ssa_open(os.fsencode(index_dir), C.byref(self._h)) != 0:
            raise RuntimeError(H.movi_host_last_error().decode())
        self.rate = H.movi_host_ssa_rate(self._h)
        self.sampled = H.movi_host_ssa_sampled(self._h)
        self.all_p = H.movi_host_ssa_all_p(self._h)
        self.rows = H.movi_host_ssa_rows(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            N.host_lib().movi_host_ssa_close(self._h)
            self._h = None


def emulate_color(host_index, colors, bases, offsets, min_len=1, pvalue=False, lanes=7, wide=False):
    bases = np.ascontiguousarray(bases, np.uint8); offsets = np.ascontiguousarray(offsets, np.uint64)
    n = len(offsets) - 1
    res = np.zeros(n, COLOR_RESULT_DTYPE)
    ml = np.zeros(int(offsets[-1]) + 1, np.uint16)
    rc = emulator().movi_ext_emulate(C.addressof(host_index.desc), EXT_COLOR | (EXT_PVALUE if pvalue else 0), C.addressof(colors.desc), 0, None, None,
                                     bases.ctypes.data, offsets.ctypes.data, n, int(min_len), lanes, int(wide), res.ctypes.data, ml.ctypes.data, None)
    assert rc == 0, rc
    return res, ml[:-1]


def emulate_sa(host_index, ssa, bases, offsets, lanes=5, wide=False):
    bases = np.ascontiguousarray(bases, np.uint8); offsets = np.ascontiguousarray(offsets, np.uint64)
    n = len(offsets) - 1
    ml = np.zeros(int(offsets[-1]) + 1, np.uint16); sa = np.zeros(int(offsets[-1]) + 1, np.uint64)
    rc = emulator().movi_ext_emulate(C.addressof(host_index.desc), EXT_SA, None, ssa.rate, ssa.sampled, ssa.all_p, bases.ctypes.data, offsets.ctypes.data,
                                     n, 0, lanes, int(wide), None, ml.ctypes.data, sa.ctypes.data)
    assert rc == 0, rc
    return ml[:-1], sa[:-1]


def result_scores(res, pvalue):
    """(best, second) counters of a result array as float64, the way the oracle reports them."""
    if pvalue:
        return np.stack([res["best_raw"].view(np.float64), res["second_raw"].view(np.float64)], axis=1)
    return np.stack([res["best_raw"].astype(np.float64), res["second_raw"].astype(np.float64)], axis=1)


def random_reads(ref_index_dir, n, seed, max_len=600):
    """Ragged reads for the colour index: pieces of its documents with errors, random ones, some with N, empty and 1-base reads."""
    import lzma
    rng = np.random.default_rng(seed)
    src = os.path.join(ROOT, "tests", "golden", "color", "reads.fa")
    seqs = [l.strip() for l in open(src) if not l.startswith(">")]
    pool = "".join(s for s in seqs if len(s) > 50)
    reads = []
    for i in range(n):
        L = int(rng.integers(0, max_len)) if i % 11 else int(rng.integers(0, 3))
        p = int(rng.integers(0, max(1, len(pool) - L)))
        s = bytearray(pool[p:p + L].encode())
        for j in range(len(s)):
            u = rng.random()
            if u < 0.01:
                s[j] = b"ACGT"[int(rng.integers(0, 4))]
            elif u < 0.013:
                s[j] = ord("N")
        if i % 13 == 5:
            s = bytearray(bytes(rng.choice(list(b"ACGT"), size=L).astype(np.uint8)))
        reads.append(bytes(s))
    return reads
