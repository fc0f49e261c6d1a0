"""Python side of datagen/synth_gpu.cu: human-scale synthetic move tables built directly in HBM (bench data, not product)."""
import ctypes as C
import os
import struct

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

    def write_index_file(self, directory):
        """index.movi (regular-thresholds) of this table, for the CPU oracle / the reference (small tables only)."""
        rows = np.empty(self.r * 2, dtype=np.uint32)
        self._check(self._L.movi_gpu_index_export_rows(self._ctx, 0, self.r, rows.ctypes.data))   # back in the reference's row format
        d = self.desc
        os.makedirs(directory, exist_ok=True)
        alphamap = np.full(256, 256, dtype=np.uint64)
        for k, ch in enumerate(b"ACGT"):
            alphamap[ch] = k
        with open(os.path.join(directory, "index.movi"), "wb") as f:
            h = bytearray(48)
            struct.pack_into("<IBBBb", h, 0, 0x4D4F5649, 2, 0, 0, 6)
            struct.pack_into("<QQQQ", h, 16, self.n, self.r, self.r, d.end_bwt_idx)
            f.write(h)
            f.write(struct.pack("<4Q", *d.end_bwt_idx_thresholds)); f.write(b"\0" * 64)
            f.write(struct.pack("<Q", 256)); f.write(alphamap.tobytes())
            f.write(struct.pack("<Q", 4)); f.write(b"ACGT"); f.write(struct.pack("<HB", 0, 0))
            f.write(rows.tobytes())
            f.write(struct.pack("<QQQ", 0, 0, 0))
            f.write(struct.pack("<Q4Q", 4, 0, 0, 0, 0))          # counts: not used by the query path
            f.write(struct.pack("<Q", 5))
            for arr in (d.last_runs, d.last_offsets, d.first_runs, d.first_offsets):
                f.write(struct.pack("<5Q", *arr[:5]))
        return directory
