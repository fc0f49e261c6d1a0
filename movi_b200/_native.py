"""ctypes bindings of the two C ABIs (include/movi_gpu.h, include/movi_host.h).

The GPU library is the product: if it is missing or no CUDA device is present, calls FAIL
(there is no CPU fallback anywhere in this package).
"""
import ctypes as C
import os

from . import build as _build

u8p, u16p, u32p, u64p, i32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64, C.c_int32))


class MoviGpuChunkResult(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_bases", C.c_uint64), ("id_bytes", C.c_uint64), ("consumed_bytes", C.c_uint64),
                ("image_bytes", C.c_uint64), ("parse_flags", C.c_uint32), ("reserved", C.c_uint32)]


class MoviGpuIndexDesc(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("sigma", C.c_uint32),
        ("length", C.c_uint64), ("r", C.c_uint64), ("end_bwt_idx", C.c_uint64),
        ("rows", C.c_void_p),
        ("alphabet", C.c_uint8 * 8),
        ("alphamap", C.c_void_p),
        ("end_bwt_idx_thresholds", C.c_uint64 * 4),
        ("first_runs", C.c_uint64 * 8), ("first_offsets", C.c_uint64 * 8),
        ("last_runs", C.c_uint64 * 8), ("last_offsets", C.c_uint64 * 8),
        ("id_blocks", C.c_void_p), ("id_blocks_size", C.c_uint64), ("block_size", C.c_uint64),
        ("tally_ids", C.c_void_p), ("tally_len", C.c_uint64), ("tally_checkpoints", C.c_uint32), ("reserved0", C.c_uint32),
        ("separators_thresholds", C.c_void_p), ("sep_count", C.c_uint64),
        ("separators_map", C.c_void_p), ("sep_map_size", C.c_uint64),
    ]


class MoviGpuColorDesc(C.Structure):
    _fields_ = [("flat_colors", C.c_void_p), ("flat_size", C.c_uint64), ("doc_set_flat_inds", C.c_void_p),
                ("num_species", C.c_uint32), ("reserved", C.c_uint32), ("log_lens", C.c_void_p)]


class MoviGpuColorResult(C.Structure):
    _fields_ = [("best_doc", C.c_uint32), ("second_best_doc", C.c_uint32), ("sum_matching_lengths", C.c_uint32), ("reserved", C.c_uint32),
                ("best_raw", C.c_uint64), ("second_raw", C.c_uint64)]


GPU_SYMBOLS = {
    "movi_gpu_abi_version": (C.c_int, []),
    "movi_gpu_device_count": (C.c_int, []),
    "movi_gpu_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "movi_gpu_destroy": (C.c_int, [C.c_void_p]),
    "movi_gpu_last_error": (C.c_char_p, [C.c_void_p]),
    "movi_gpu_upload_index": (C.c_int, [C.c_void_p, C.POINTER(MoviGpuIndexDesc), C.c_int]),
    "movi_gpu_adopt_device_rows": (C.c_int, [C.c_void_p, C.POINTER(MoviGpuIndexDesc), C.c_void_p]),
    "movi_gpu_share_index": (C.c_int, [C.c_void_p, C.c_void_p]),
    "movi_gpu_index_device_rows": (C.c_void_p, [C.c_void_p]),
    "movi_gpu_index_device_layout": (C.c_int, [C.c_void_p]),
    "movi_gpu_index_export_rows": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "movi_gpu_index_device_bytes": (C.c_uint64, [C.c_void_p]),
    "movi_gpu_query_pml": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]),
    "movi_gpu_query_ml8": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]),
    "movi_gpu_query_classify": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "movi_gpu_bpf_image_bytes": (C.c_uint64, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "movi_gpu_query_bpf": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "movi_gpu_bpf_assemble_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "movi_gpu_query_file_chunk": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_uint64,
                                            C.POINTER(MoviGpuChunkResult)]),
    "movi_gpu_file_chunk_parse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(MoviGpuChunkResult)]),
    "movi_gpu_file_chunk_query": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64]),
    "movi_gpu_file_chunk_classify": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_uint64, C.c_int, C.c_uint64]),
    "movi_gpu_file_chunk_filter": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]),
    "movi_gpu_file_chunk_filter_deliver": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]),
    "movi_gpu_file_chunk_fetch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "movi_gpu_parse_chunk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                       C.c_void_p, C.c_uint64, C.POINTER(MoviGpuChunkResult)]),
    "movi_gpu_query_zml": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]),
    "movi_gpu_query_count": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]),
    "movi_gpu_query_pml_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]),
    "movi_gpu_query_zml_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]),
    "movi_gpu_query_count_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "movi_gpu_check_device_status": (C.c_int, [C.c_void_p]),
    "movi_gpu_classify_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "movi_gpu_upload_colors": (C.c_int, [C.c_void_p, C.POINTER(MoviGpuColorDesc)]),
    "movi_gpu_query_multi_classify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]),
    "movi_gpu_upload_sampled_sa": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]),
    "movi_gpu_query_pml_sa": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]),
    "movi_gpu_launch_count": (C.c_uint64, [C.c_void_p]),
    "movi_gpu_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "movi_gpu_get_option": (C.c_int64, [C.c_void_p, C.c_char_p]),
    "movi_gpu_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    "movi_gpu_host_free": (C.c_int, [C.c_void_p]),
    "movi_gpu_gather_probe": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
}

HOST_SYMBOLS = {
    "movi_host_last_error": (C.c_char_p, []),
    "movi_host_index_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "movi_host_index_desc": (C.c_int, [C.c_void_p, C.POINTER(MoviGpuIndexDesc)]),
    "movi_host_index_original_r": (C.c_uint64, [C.c_void_p]),
    "movi_host_index_type_name": (C.c_char_p, [C.c_void_p]),
    "movi_host_index_close": (None, [C.c_void_p]),
    "movi_host_reads_load": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "movi_host_reader_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "movi_host_reader_next": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "movi_host_reader_close": (None, [C.c_void_p]),
    "movi_host_reads_count": (C.c_uint64, [C.c_void_p]),
    "movi_host_reads_bases": (C.c_void_p, [C.c_void_p]),
    "movi_host_reads_offsets": (C.c_void_p, [C.c_void_p]),
    "movi_host_reads_id": (C.c_void_p, [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]),
    "movi_host_reads_free": (None, [C.c_void_p]),
    "movi_host_nulldb_threshold": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_uint64)]),
    "movi_host_write_bpf": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p]),
    "movi_host_reads_ids": (C.c_void_p, [C.c_void_p]),
    "movi_host_reads_id_offsets": (C.c_void_p, [C.c_void_p]),
    "movi_host_bpf_append_image": (C.c_int, [C.c_int, C.c_void_p, C.c_uint64]),
    "movi_host_reader_next_reads": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p)]),
    "movi_host_reader_raw": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    "movi_host_record_boundary": (C.c_uint64, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int]),
    "movi_host_bpf_begin": (C.c_int, [C.c_char_p]),
    "movi_host_bpf_append": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p]),
    "movi_host_bpf_end": (C.c_int, [C.c_int]),
    "movi_host_write_ml_text": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p]),
    "movi_host_write_counts": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "movi_host_classify_bins": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "movi_host_classify": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "movi_host_reader_seek": (C.c_int, [C.c_void_p, C.c_uint64]),
    "movi_host_reads_adopt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "movi_host_colors_open": (C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "movi_host_colors_open_kind": (C.c_int, [C.c_char_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]),
    "movi_host_colors_desc": (C.c_int, [C.c_void_p, C.POINTER(MoviGpuColorDesc)]),
    "movi_host_colors_num_species": (C.c_uint32, [C.c_void_p]),
    "movi_host_colors_taxon": (C.c_uint32, [C.c_void_p, C.c_uint32]),
    "movi_host_colors_close": (None, [C.c_void_p]),
    "movi_host_write_multi_classify": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]),
    "movi_host_ssa_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "movi_host_ssa_rate": (C.c_uint64, [C.c_void_p]),
    "movi_host_ssa_rows": (C.c_uint64, [C.c_void_p]),
    "movi_host_ssa_samples": (C.c_uint64, [C.c_void_p]),
    "movi_host_ssa_sampled": (C.c_void_p, [C.c_void_p]),
    "movi_host_ssa_all_p": (C.c_void_p, [C.c_void_p]),
    "movi_host_ssa_close": (None, [C.c_void_p]),
    "movi_host_sa_append": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p]),
}

_gpu = None
_host = None


def _bind(lib, table):
    for name, (res, args) in table.items():
        f = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        f.restype = res
        f.argtypes = args
    return lib


def gpu_lib(build=True):
    """libmovi_gpu.so (built in-tree on first use).  Raises if it cannot be built or loaded."""
    global _gpu
    if _gpu is None:
        if build:
            _build.build_gpu()
        if not os.path.exists(_build.GPU_LIB):
            raise RuntimeError("libmovi_gpu.so is missing: run `python -m movi_b200.build` (no CPU fallback exists)")
        _gpu = _bind(C.CDLL(_build.GPU_LIB), GPU_SYMBOLS)
    return _gpu


def host_lib(build=True):
    global _host
    if _host is None:
        if build:
            _build.build_host()
        _host = _bind(C.CDLL(_build.HOST_LIB), HOST_SYMBOLS)
    return _host
