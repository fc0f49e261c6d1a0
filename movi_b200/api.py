"""Thin object layer over the C ABIs; every compute call goes through libmovi_gpu.so."""
import ctypes as C
import os

import numpy as np

from . import _native as N

LAYOUT_AUTO, LAYOUT_EXPANDED8, LAYOUT_COMPACT = 0, 1, 2


class MoviGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("movi_gpu error %d: %s" % (code, msg))
        self.code = code


def device_count():
    return N.gpu_lib().movi_gpu_device_count()


FLAG_RPML = 1          # MOVI_GPU_FLAG_RPML (include/movi_gpu.h)


class HostIndex:
    """index.movi opened on the host (MoveStructure::deserialize, reference src/move_structure_io.cpp:471-511)."""

    def __init__(self, path):
        self._h = C.c_void_p()
        L = N.host_lib()
        if L.movi_host_index_open(os.fsencode(path), C.byref(self._h)) != 0:
            raise RuntimeError(L.movi_host_last_error().decode())
        self.desc = N.MoviGpuIndexDesc()
        L.movi_host_index_desc(self._h, C.byref(self.desc))
        self.type_name = L.movi_host_index_type_name(self._h).decode()
        self.original_r = L.movi_host_index_original_r(self._h)
        self.path = path

    mode = property(lambda s: s.desc.mode)
    r = property(lambda s: s.desc.r)
    n = property(lambda s: s.desc.length)
    end_bwt_idx = property(lambda s: s.desc.end_bwt_idx)
    alphabet = property(lambda s: bytes(s.desc.alphabet[: s.desc.sigma]))

    def close(self):
        if self._h:
            N.host_lib().movi_host_index_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Reads:
    """A FASTA / FASTQ file parsed like the reference's BatchLoader (src/batch_loader.cpp:26-143)."""

    def __init__(self, path, reverse=False):
        self._h = C.c_void_p()
        L = N.host_lib()
        if L.movi_host_reads_load(os.fsencode(path), int(reverse), C.byref(self._h)) != 0:
            raise RuntimeError(L.movi_host_last_error().decode())
        self.count = L.movi_host_reads_count(self._h)
        off = L.movi_host_reads_offsets(self._h)
        self.offsets = np.ctypeslib.as_array(C.cast(off, N.u64p), shape=(self.count + 1,)).copy()
        total = int(self.offsets[-1])
        b = L.movi_host_reads_bases(self._h)
        self.bases = np.ctypeslib.as_array(C.cast(b, N.u8p), shape=(total,)).copy() if total else np.zeros(0, np.uint8)

    @property
    def id_blob(self):
        """(ids uint8[total], id_offsets uint64[count + 1]): the read ids concatenated, as movi_gpu_query_bpf takes them."""
        L = N.host_lib()
        off = np.ctypeslib.as_array(C.cast(L.movi_host_reads_id_offsets(self._h), N.u64p), shape=(self.count + 1,)).copy()
        total = int(off[-1])
        blob = np.ctypeslib.as_array(C.cast(L.movi_host_reads_ids(self._h), N.u8p), shape=(total,)).copy() if total else np.zeros(0, np.uint8)
        return blob, off

    def id(self, i):
        n = C.c_uint32()
        p = N.host_lib().movi_host_reads_id(self._h, i, C.byref(n))
        return C.string_at(p, n.value).decode()

    def ids(self):
        return [self.id(i) for i in range(self.count)]

    def read(self, i):
        return self.bases[int(self.offsets[i]):int(self.offsets[i + 1])].tobytes().decode()

    def close(self):
        if self._h:
            N.host_lib().movi_host_reads_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


COLOR_RESULT_DTYPE = np.dtype([("best_doc", "<u4"), ("second_best_doc", "<u4"), ("sum_matching_lengths", "<u4"), ("reserved", "<u4"),
                               ("best_raw", "<u8"), ("second_raw", "<u8")])      # MoviGpuColorResult, include/movi_gpu.h


def _ptr(a):
    return a.ctypes.data if isinstance(a, np.ndarray) else int(a)


class MoveStructure:
    """The move structure resident in one GPU's HBM + the three queries of the hot path.

    Mirrors the reference's MoveStructure for this path: deserialize() -> upload, query_pml / query_zml /
    query_backward_search on whole batches (numpy host arrays), *_device variants on raw device pointers
    (e.g. torch tensors' data_ptr()).
    """

    def __init__(self, index, device=0, layout=LAYOUT_AUTO):
        self._L = N.gpu_lib()
        self._ctx = C.c_void_p()
        rc = self._L.movi_gpu_create(device, C.byref(self._ctx))
        if rc != 0:
            raise MoviGpuError(rc, self._L.movi_gpu_last_error(None).decode())
        self.device = device
        self.host_index = index if isinstance(index, HostIndex) else HostIndex(index)
        rc = self._L.movi_gpu_upload_index(self._ctx, C.byref(self.host_index.desc), layout)
        self._check(rc)
        self.r = self.host_index.r
        self.mode = self.host_index.mode

    @classmethod
    def sharing(cls, owner):
        """A second context on the owner's GPU that queries the owner's uploaded index (no copy): one context serves one host
        thread at a time, two let two batches overlap.  The owner must outlive it."""
        self = cls.__new__(cls)
        self._L = owner._L
        self._ctx = C.c_void_p()
        rc = self._L.movi_gpu_create(owner.device, C.byref(self._ctx))
        if rc != 0:
            raise MoviGpuError(rc, self._L.movi_gpu_last_error(None).decode())
        self.device, self.host_index, self.r, self.mode = owner.device, owner.host_index, owner.r, owner.mode
        self._check(self._L.movi_gpu_share_index(self._ctx, owner._ctx))
        return self

    def _check(self, rc):
        if rc != 0:
            raise MoviGpuError(rc, self._L.movi_gpu_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_colors", None):
            N.host_lib().movi_host_colors_close(self._colors)
            self._colors = None
        if getattr(self, "_ctx", None):
            self._L.movi_gpu_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- options / introspection
    def set_option(self, key, value):
        self._check(self._L.movi_gpu_set_option(self._ctx, key.encode(), int(value)))

    def get_option(self, key):
        return self._L.movi_gpu_get_option(self._ctx, key.encode())

    @property
    def launch_count(self):
        return self._L.movi_gpu_launch_count(self._ctx)

    @property
    def device_layout(self):
        return self._L.movi_gpu_index_device_layout(self._ctx)

    @property
    def index_device_bytes(self):
        return self._L.movi_gpu_index_device_bytes(self._ctx)

    # ---- host-buffer queries (copies inside)
    @staticmethod
    def _prep(bases, offsets):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        return bases, offsets, len(offsets) - 1

    def query_pml(self, bases, offsets, out=None, rpml=False):
        """rpml=True: `movi query --rpml` (reposition_randomly instead of the thresholds)."""
        bases, offsets, n = self._prep(bases, offsets)
        if out is None:
            out = np.empty(int(offsets[-1]) if n >= 0 and len(offsets) else 0, dtype=np.uint16)
        self._check(self._L.movi_gpu_query_pml(self._ctx, _ptr(bases), _ptr(offsets), max(n, 0), FLAG_RPML if rpml else 0, _ptr(out)))
        return out

    def query_classify(self, bases, offsets, bin_width, max_value_thr, zml=False, want_ml=False, rpml=False):
        """PML (or ZML) + binary classification on the device -> (bins_above, bins_below, sum_max[, ml])."""
        bases, offsets, n = self._prep(bases, offsets)
        n = max(n, 0)
        above, below = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        total = np.zeros(n, np.uint64)
        ml = np.empty(int(offsets[-1]) if len(offsets) else 0, dtype=np.uint16) if want_ml else None
        self._check(self._L.movi_gpu_query_classify(self._ctx, int(zml), _ptr(bases), _ptr(offsets), n, FLAG_RPML if rpml else 0, bin_width, max_value_thr,
                                                    _ptr(above), _ptr(below), _ptr(total), _ptr(ml) if want_ml else None))
        return (above, below, total, ml) if want_ml else (above, below, total)

    # ---- Movi Color (reference `movi query --multi-classify` / `--sa-entries`, --no-prefetch path)
    def load_colors(self, index_dir=None):
        """load_color_table (src/movi.cpp:121-152): doc_sets_flat.bin + document info of the index directory -> HBM."""
        H = N.host_lib()
        d = index_dir or self.host_index.path
        if os.path.isfile(d):
            d = os.path.dirname(d)
        h = C.c_void_p()
        if H.movi_host_colors_open(os.fsencode(d), self.r, C.byref(h)) != 0:
            raise RuntimeError(H.movi_host_last_error().decode())
        if getattr(self, "_colors", None):
            H.movi_host_colors_close(self._colors)
        self._colors = h
        desc = N.MoviGpuColorDesc()
        H.movi_host_colors_desc(h, C.byref(desc))
        self._check(self._L.movi_gpu_upload_colors(self._ctx, C.byref(desc)))
        self.num_species = H.movi_host_colors_num_species(h)
        return self.num_species

    def query_multi_classify(self, bases, offsets, min_len=1, pvalue=False, want_ml=False):
        """-> numpy record array (best_doc, second_best_doc, sum_matching_lengths, best_raw, second_raw) per read [, ml]."""
        bases, offsets, n = self._prep(bases, offsets)
        n = max(n, 0)
        res = np.zeros(n, dtype=COLOR_RESULT_DTYPE)
        ml = np.empty(int(offsets[-1]) if len(offsets) else 0, dtype=np.uint16) if want_ml else None
        self._check(self._L.movi_gpu_query_multi_classify(self._ctx, _ptr(bases), _ptr(offsets), n, 0, int(min_len), int(bool(pvalue)),
                                                          _ptr(res), _ptr(ml) if want_ml else None))
        return (res, ml) if want_ml else res

    def multi_classify_text(self, reads, res, min_len=1, pvalue=False):
        """The reference's -o file for these reads: "<id>,<taxon>,<taxon>" per line (src/move_structure_query.cpp:439-471)."""
        H = N.host_lib()
        rfd, wfd = os.pipe()
        try:
            import threading
            buf = []
            t = threading.Thread(target=lambda: buf.append(os.fdopen(rfd, "rb").read()))
            t.start()
            rc = H.movi_host_write_multi_classify(wfd, reads._h, self._colors, _ptr(res), int(min_len), int(bool(pvalue)))
        finally:
            os.close(wfd)
            t.join()
        if rc != 0:
            raise RuntimeError(H.movi_host_last_error().decode())
        return buf[0].decode()

    def load_sampled_sa(self, index_dir=None):
        """deserialize_sampled_SA (src/move_structure_io.cpp:716-735): ssa.movi -> HBM."""
        H = N.host_lib()
        d = index_dir or self.host_index.path
        if os.path.isfile(d):
            d = os.path.dirname(d)
        h = C.c_void_p()
        if H.movi_host_ssa_open(os.fsencode(d), C.byref(h)) != 0:
            raise RuntimeError(H.movi_host_last_error().decode())
        try:
            if H.movi_host_ssa_rows(h) != self.r:
                raise RuntimeError("ssa.movi belongs to another index (%d rows, the index has %d)" % (H.movi_host_ssa_rows(h), self.r))
            self._check(self._L.movi_gpu_upload_sampled_sa(self._ctx, H.movi_host_ssa_rate(h), H.movi_host_ssa_sampled(h),
                                                           H.movi_host_ssa_samples(h), H.movi_host_ssa_all_p(h)))
            return H.movi_host_ssa_rate(h)
        finally:
            H.movi_host_ssa_close(h)

    def query_pml_sa(self, bases, offsets):
        """`--pml --sa-entries` -> (ml uint16[total], sa uint64[total]), same orientation."""
        bases, offsets, n = self._prep(bases, offsets)
        total = int(offsets[-1]) if len(offsets) else 0
        ml, sa = np.empty(total, np.uint16), np.empty(total, np.uint64)
        self._check(self._L.movi_gpu_query_pml_sa(self._ctx, _ptr(bases), _ptr(offsets), max(n, 0), 0, _ptr(ml), _ptr(sa)))
        return ml, sa

    def export_rows(self, first=0, count=None):
        """Rows of the resident 8-byte table back in the reference's regular-thresholds MoveRow format -> uint32[count, 2]."""
        if count is None:
            count = self.r - first
        out = np.empty((count, 2), dtype=np.uint32)
        self._check(self._L.movi_gpu_index_export_rows(self._ctx, first, count, _ptr(out)))
        return out

    def query_ml8(self, bases, offsets, zml=False, out=None, rpml=False):
        """PML (or ZML) of a batch of reads of at most 255 bases, returned as uint8 (half the copy-back of query_pml)."""
        bases, offsets, n = self._prep(bases, offsets)
        if out is None:
            out = np.empty(int(offsets[-1]) if len(offsets) else 0, dtype=np.uint8)
        self._check(self._L.movi_gpu_query_ml8(self._ctx, int(zml), _ptr(bases), _ptr(offsets), max(n, 0), FLAG_RPML if rpml else 0, _ptr(out)))
        return out

    def query_bpf(self, bases, offsets, ids, id_offsets, zml=False, rpml=False):
        """PML (or ZML) with the .bpf records assembled on the device -> uint8 image = the bytes the reference appends to
        <prefix>.pml.bpf after the 12-byte header (src/utils.cpp:202-246)."""
        bases, offsets, n = self._prep(bases, offsets)
        ids = np.ascontiguousarray(ids, dtype=np.uint8)
        id_offsets = np.ascontiguousarray(id_offsets, dtype=np.uint64)
        n = max(n, 0)
        size = self._L.movi_gpu_bpf_image_bytes(_ptr(offsets), _ptr(id_offsets), n) if n else 0
        image = np.empty(size, dtype=np.uint8)
        self._check(self._L.movi_gpu_query_bpf(self._ctx, int(zml), _ptr(bases), _ptr(offsets), n, FLAG_RPML if rpml else 0,
                                               _ptr(ids), _ptr(id_offsets), _ptr(image), size))
        return image

    def parse_chunk(self, raw, fastq, final=True):
        """A chunk of FASTA / FASTQ bytes parsed on the device -> (bases, offsets, ids, id_offsets, consumed_bytes, parse_flags)."""
        raw = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else raw
        n = len(raw)
        bases, ids = np.zeros(n + 1, np.uint8), np.zeros(n + 1, np.uint8)
        offsets, id_offsets = np.zeros(n // 2 + 2, np.uint64), np.zeros(n // 2 + 2, np.uint64)
        res = N.MoviGpuChunkResult()
        self._check(self._L.movi_gpu_parse_chunk(self._ctx, _ptr(raw) if n else None, n, int(fastq), int(final), _ptr(bases), n + 1, _ptr(offsets),
                                                 _ptr(ids), n + 1, _ptr(id_offsets), n // 2 + 1, C.byref(res)))
        k = int(res.n_reads)
        return (bases[: int(res.n_bases)], offsets[: k + 1], ids[: int(res.id_bytes)], id_offsets[: k + 1], int(res.consumed_bytes),
                int(res.parse_flags))

    def query_file_chunk(self, raw, fastq, final=True, zml=False, rpml=False, want_image=True):
        """File bytes in, .bpf records out: parse + walk + record assembly on the device -> (image uint8[], result struct)."""
        raw = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else raw
        n = len(raw)
        cap = 3 * n + 64
        image = np.zeros(cap if want_image else 1, np.uint8)
        res = N.MoviGpuChunkResult()
        self._check(self._L.movi_gpu_query_file_chunk(self._ctx, int(zml), _ptr(raw) if n else None, n, int(fastq), int(final),
                                                      FLAG_RPML if rpml else 0, _ptr(image) if want_image else None, cap, C.byref(res)))
        return image[: int(res.image_bytes)] if want_image else None, res

    def classify_file_chunk(self, raw, fastq, bin_width, max_value_thr, final=True, zml=False, want_reads=False):
        """File bytes in, the three report columns per read out (movi_gpu_file_chunk_parse + _classify [+ _fetch]) ->
        (above, below, sum_max, result struct[, (bases, offsets, ids, id_offsets)])."""
        raw = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else raw
        res = N.MoviGpuChunkResult()
        self._check(self._L.movi_gpu_file_chunk_parse(self._ctx, _ptr(raw) if len(raw) else None, len(raw), int(fastq), int(final), C.byref(res)))
        n = int(res.n_reads)
        above, below, total = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint64)
        if res.parse_flags == 0 and n:
            self._check(self._L.movi_gpu_file_chunk_classify(self._ctx, int(zml), 0, bin_width, max_value_thr, _ptr(above), _ptr(below), _ptr(total),
                                                             None, 0, -1, 0))
        if not want_reads:
            return above, below, total, res
        bases, ids = np.zeros(int(res.n_bases), np.uint8), np.zeros(int(res.id_bytes), np.uint8)
        off, id_off = np.zeros(n + 1, np.uint64), np.zeros(n + 1, np.uint64)
        if res.parse_flags == 0 and n:
            self._check(self._L.movi_gpu_file_chunk_fetch(self._ctx, _ptr(bases), _ptr(off), _ptr(ids), _ptr(id_off)))
        return above, below, total, res, (bases, off, ids, id_off)

    def filter_file_chunk(self, raw, fastq, bin_width, max_value_thr, invert=False, final=True, zml=False):
        """File bytes in, the echo of the reads `--filter [-v]` keeps out (assembled on the device) -> (text bytes, result struct)."""
        raw = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else raw
        res = N.MoviGpuChunkResult()
        self._check(self._L.movi_gpu_file_chunk_parse(self._ctx, _ptr(raw) if len(raw) else None, len(raw), int(fastq), int(final), C.byref(res)))
        if res.parse_flags or res.n_reads == 0:
            return b"", res
        nb = C.c_uint64()
        self._check(self._L.movi_gpu_file_chunk_filter(self._ctx, int(zml), 0, bin_width, max_value_thr, int(invert), C.byref(nb)))
        text = np.zeros(nb.value + 1, np.uint8)
        self._check(self._L.movi_gpu_file_chunk_filter_deliver(self._ctx, _ptr(text), nb.value, -1))
        return text[: nb.value].tobytes(), res

    def query_zml(self, bases, offsets, out=None):
        bases, offsets, n = self._prep(bases, offsets)
        if out is None:
            out = np.empty(int(offsets[-1]) if len(offsets) else 0, dtype=np.uint16)
        self._check(self._L.movi_gpu_query_zml(self._ctx, _ptr(bases), _ptr(offsets), max(n, 0), 0, _ptr(out)))
        return out

    def query_backward_search(self, bases, offsets):
        """count query -> (counts uint64[n], pos_on_r int32[n])"""
        bases, offsets, n = self._prep(bases, offsets)
        counts = np.zeros(max(n, 0), dtype=np.uint64)
        pos = np.zeros(max(n, 0), dtype=np.int32)
        self._check(self._L.movi_gpu_query_count(self._ctx, _ptr(bases), _ptr(offsets), max(n, 0), 0, _ptr(counts), _ptr(pos)))
        return counts, pos

    query_count = query_backward_search

    # ---- raw pointer forms (host pinned or device pointers as ints)
    def query_pml_ptr(self, bases_ptr, offsets_ptr, n_reads, out_ptr, rpml=False):
        self._check(self._L.movi_gpu_query_pml(self._ctx, bases_ptr, offsets_ptr, n_reads, FLAG_RPML if rpml else 0, out_ptr))

    def query_pml_device(self, d_bases, d_offsets, n_reads, d_out, stream=0, rpml=False):
        self._check(self._L.movi_gpu_query_pml_device(self._ctx, d_bases, d_offsets, n_reads, FLAG_RPML if rpml else 0, d_out, stream))

    def query_zml_device(self, d_bases, d_offsets, n_reads, d_out, stream=0):
        self._check(self._L.movi_gpu_query_zml_device(self._ctx, d_bases, d_offsets, n_reads, 0, d_out, stream))

    def query_count_device(self, d_bases, d_offsets, n_reads, d_counts, d_pos, stream=0):
        self._check(self._L.movi_gpu_query_count_device(self._ctx, d_bases, d_offsets, n_reads, 0, d_counts, d_pos, stream))

    def check_device_status(self):
        self._check(self._L.movi_gpu_check_device_status(self._ctx))

    def classify_device(self, d_ml, d_offsets, n_reads, bin_width, max_value_thr, d_above, d_below, d_sum, stream=0):
        self._check(self._L.movi_gpu_classify_device(self._ctx, d_ml, d_offsets, n_reads, bin_width, max_value_thr,
                                                     d_above, d_below, d_sum, stream))

    def gather_probe(self, table_bytes, load_bytes=8, dependent=True, steps=256, blocks_per_sm=8, threads_per_block=256):
        sec, loads = C.c_double(), C.c_uint64()
        self._check(self._L.movi_gpu_gather_probe(self._ctx, table_bytes, load_bytes, int(dependent), steps, blocks_per_sm,
                                                  threads_per_block, C.byref(sec), C.byref(loads)))
        return sec.value, loads.value
