"""Python side of datagen/synth_gpu.cu: human-scale synthetic move tables built directly in HBM (bench data, not product)."""
import ctypes as C

import numpy as np

from .. import _native as N
from .. import build as _build
from ..api import MoveStructure, MoviGpuError

_lib = None


def lib():
    global _lib
    if _lib is None:
        _build.build_synth()
        N.gpu_lib()
        L = C.CDLL(_build.SYNTH_LIB)
        L.movi_synth_index.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(N.MoviGpuIndexDesc)]
        L.movi_synth_index.restype = C.c_int
        L.movi_synth_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.movi_synth_reads.restype = C.c_int
        L.movi_synth_reads_resync.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                              C.c_uint32, C.c_uint64]
        L.movi_synth_reads_resync.restype = C.c_int
        _lib = L
    return _lib


class SyntheticStructure(MoveStructure):
    """A MoveStructure whose row table was synthesised on the GPU (r rows, seeded)."""

    def __init__(self, r, seed=44, device=0, profile=0):
        self._L = N.gpu_lib()
        self._ctx = C.c_void_p()
        rc = self._L.movi_gpu_create(device, C.byref(self._ctx))
        if rc != 0:
            raise MoviGpuError(rc, self._L.movi_gpu_last_error(None).decode())
        self.device, self.seed = device, seed
        n = C.c_uint64()
        self.desc = N.MoviGpuIndexDesc()
        rc = lib().movi_synth_index(self._ctx, r, seed, profile, C.byref(n), C.byref(self.desc))
        if rc != 0:
            raise MoviGpuError(rc, self._L.movi_gpu_last_error(self._ctx).decode() or "synthetic index generation failed")
        self.r, self.n, self.mode, self.host_index = r, n.value, 6, None

    @property
    def device_rows(self):
        return self._L.movi_gpu_index_device_rows(self._ctx)

    def make_reads(self, d_bases_ptr, n_reads, read_len, sub_per_1024=51, rand_every=10):
        """Fill device memory with n_reads x read_len matching reads (LF walks, ~5 % substitutions, 1 in 10 random)."""
        rc = lib().movi_synth_reads(self.device_rows, self.r, self.seed, d_bases_ptr, n_reads, read_len, sub_per_1024, rand_every)
        if rc != 0:
            raise RuntimeError("movi_synth_reads failed")

    def make_reads_resync(self, d_bases_ptr, n_reads, read_len, sub_per_1024=51, rand_every=10, lead=24, read_seed=0):
        """Reads written down by the query's own walk: after each substitution the generator repositions the way the query will,
        so the read keeps matching along the new path (PMLs recover after an error, as on a suffix-sorted index)."""
        rc = lib().movi_synth_reads_resync(self.device_rows, self.r, self.seed, d_bases_ptr, n_reads, read_len, sub_per_1024, rand_every,
                                           lead, read_seed)
        if rc != 0:
            raise RuntimeError("movi_synth_reads_resync failed")

    def write_index_file(self, directory, chunk_rows=1 << 27):
        """index.movi (regular-thresholds) of this table, for the CPU oracle / the unmodified reference.  The rows come back from
        HBM in the reference's row format 1 GB at a time, so a table of any size costs no more host memory than that."""
        from .index_file import write_regular_thresholds
        d = self.desc

        def chunks():
            for first in range(0, self.r, chunk_rows):
                m = min(chunk_rows, self.r - first)
                rows = np.empty(m * 2, dtype=np.uint32)
                self._check(self._L.movi_gpu_index_export_rows(self._ctx, first, m, rows.ctypes.data))
                yield rows

        return write_regular_thresholds(directory, self.n, self.r, d.end_bwt_idx, list(d.end_bwt_idx_thresholds), chunks(),
                                        d.last_runs, d.last_offsets, d.first_runs, d.first_offsets)
