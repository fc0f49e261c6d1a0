/* movi_gpu.h -- C ABI of the B200 (sm_100a) implementation of Movi's query hot path.
 *
 * Movi (the reference) has no plugin / FFI interface; its in-process seam for this path is
 *     MoveStructure::query_pml(MoveQuery&)                      include/move_structure.hpp:150
 *     MoveStructure::query_zml(MoveQuery&)                      include/move_structure.hpp:151
 *     MoveStructure::query_backward_search(MoveQuery&, int32&)  include/move_structure.hpp:89
 * called per read from query() (src/movi.cpp:221-400) through handle_pml_zml / handle_count
 * (src/movi.cpp:23-76), or per batch from ReadProcessor::process_latency_hiding
 * (include/read_processor.hpp:62, src/read_processor.cpp:641-730).  The entry points below are what
 * a binding at that seam calls instead: whole batches of reads in, the same per-base / per-read
 * results out, bit for bit.  INTEGRATION.md shows the reference-side call sites.
 *
 * Conventions
 *   - plain C, no exceptions: every call returns MOVI_GPU_OK or an error code; the text of the last
 *     error of a context is available from movi_gpu_last_error().
 *   - one context per GPU; a context is not thread safe (use one host thread per context, like the
 *     reference uses one ReadProcessor strand set per OpenMP thread).
 *   - the host owns every host buffer; device buffers belong to the context.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Read batches are passed concatenated: read i is bases[read_offsets[i] .. read_offsets[i+1]).
 * Matching lengths come back at the same offsets in the reference's storage order, i.e. for every
 * read, element 0 belongs to the LAST base (MoveQuery::add_ml appends while the read is walked
 * right to left: include/move_query.hpp:26-38; BPF files store them in this order, src/utils.cpp:202-246).
 */
#ifndef MOVI_GPU_H
#define MOVI_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOVI_GPU_ABI_VERSION 1

enum {
    MOVI_GPU_OK = 0,
    MOVI_GPU_ERR_ARG = 1,        /* bad argument */
    MOVI_GPU_ERR_CUDA = 2,       /* CUDA runtime failure, or no device */
    MOVI_GPU_ERR_INDEX = 3,      /* unsupported / inconsistent index description */
    MOVI_GPU_ERR_QUERY = 4,      /* a read drove the walk into a state where the reference throws
                                    ("This should not happen": src/move_structure_query.cpp:348,582-598,
                                    src/move_structure.cpp:63-65) */
    MOVI_GPU_ERR_NOMEM = 5
};

/* Index modes = the reference's compile-time MODE (include/utils.hpp:133-147). */
#define MOVI_MODE_BLOCKED            2   /* 6-byte rows + id_blocks, no thresholds; PML repositions with the --rpml rule */
#define MOVI_MODE_REGULAR            3   /* 8-byte rows, no thresholds (12-bit lengths); the same */
#define MOVI_MODE_SAMPLED            5   /* 3-byte rows + id checkpoints, no thresholds; the same (PML only) */
#define MOVI_MODE_REGULAR_THRESHOLDS 6   /* 8-byte rows  */
#define MOVI_MODE_SAMPLED_THRESHOLDS 7   /* 3-byte rows + id checkpoints ("tally") */
#define MOVI_MODE_BLOCKED_THRESHOLDS 8   /* 6-byte rows + id_blocks */

/* What MoveStructure holds after deserialize() (include/move_structure.hpp:328-380), by pointer.
 * All pointers are HOST pointers; they only need to stay valid during movi_gpu_upload_index(). */
typedef struct MoviGpuIndexDesc {
    int32_t  mode;                       /* MOVI_MODE_* (MoviHeader.type, include/utils.hpp:32-61) */
    uint32_t sigma;                      /* alphabet.size(): 4, or 5 with separators ("%ACGT") */
    uint64_t length;                     /* n, BWT length */
    uint64_t r;                          /* number of move rows */
    uint64_t end_bwt_idx;                /* row of '$' */
    const void* rows;                    /* rlbwt.data(): r packed MoveRow (8 / 6 / 3 bytes each, include/move_row.hpp:104-154) */
    uint8_t  alphabet[8];                /* alphabet[i] */
    const uint64_t* alphamap;            /* 256 entries, 256 == not in the alphabet (src/move_structure_build.cpp:428-447) */
    uint64_t end_bwt_idx_thresholds[4];  /* src/move_structure_io.cpp:147 */
    uint64_t first_runs[8], first_offsets[8], last_runs[8], last_offsets[8];   /* sigma+1 used; io.cpp:269-287 */
    /* MODE 8 and 2 (src/move_structure_io.cpp:305-324) */
    const uint32_t* id_blocks;           /* sigma * id_blocks_size, character major */
    uint64_t id_blocks_size;
    uint64_t block_size;
    /* MODE 7 and 5 (src/move_structure_io.cpp:338-349) */
    const void* tally_ids;               /* sigma * tally_len MoveTally {u32 right; u8 left}, 5 bytes each */
    uint64_t tally_len;
    uint32_t tally_checkpoints;
    uint32_t reserved0;
    /* separators (src/move_structure_io.cpp:415-433); zero / NULL when the index has none (the non-threshold modes never do) */
    const uint16_t* separators_thresholds;   /* sep_count * 4 (ThresholdsRow) */
    uint64_t sep_count;
    const uint64_t* separators_map;          /* sep_map_size pairs (row, slot) in any order */
    uint64_t sep_map_size;
} MoviGpuIndexDesc;

typedef struct movi_gpu_ctx movi_gpu_ctx;

/* Query flags */
#define MOVI_GPU_FLAG_NONE 0u
/* PML entry points (query_pml, _pml_device, _classify, _ml8): `movi query --rpml`, i.e. reposition with reposition_randomly
 * (src/move_structure_query.cpp:604-688, selected at :304-311) instead of the thresholds.  Ignored by ZML / count. */
#define MOVI_GPU_FLAG_RPML 1u

/* Device-resident row layouts (chosen at upload). */
#define MOVI_GPU_LAYOUT_AUTO     0   /* EXPANDED8 when it fits in free HBM, else COMPACT where the mode has one */
#define MOVI_GPU_LAYOUT_EXPANDED8 1  /* every mode re-laid out as 8-byte rows with the full id resolved:
                                        one 32-byte sector holds 4 rows, one sector fetch per LF step */
#define MOVI_GPU_LAYOUT_COMPACT  2   /* the packed rows stay as they are on disk.  blocked-thresholds: 6 bytes per row, the
                                        destination id is completed per step from id_blocks (shared memory / L2).
                                        sampled-thresholds: 3 bytes per row + id checkpoints (4 bytes per row in all), the
                                        destination id costs a checkpoint scan + walk-back per LF step, as in the reference */

int  movi_gpu_abi_version(void);
int  movi_gpu_device_count(void);

/* replaces: MoveStructure construction per process (src/movi.cpp:661-667); one ctx per GPU */
int  movi_gpu_create(int device, movi_gpu_ctx** ctx);
int  movi_gpu_destroy(movi_gpu_ctx* ctx);
const char* movi_gpu_last_error(const movi_gpu_ctx* ctx);   /* ctx may be NULL: last create() error */

/* replaces: MoveStructure::deserialize() filling rlbwt & side tables (src/move_structure_io.cpp:471-511).
 * Copies the row table and side tables into HBM once (re-laid out, see MOVI_GPU_LAYOUT_*). */
int  movi_gpu_upload_index(movi_gpu_ctx* ctx, const MoviGpuIndexDesc* desc, int layout);
/* Same, for a row table that is ALREADY resident in this GPU's HBM in the 8-byte regular-thresholds format
 * (r + 4 rows allocated with cudaMalloc; e.g. produced by another component or a generator): desc->rows is ignored,
 * the side tables are taken from desc, ownership of d_rows moves to the context when the call succeeds (on failure it stays
 * with the caller).  The rows are rewritten IN PLACE into the device row format (same fields, other bit positions). */
int  movi_gpu_adopt_device_rows(movi_gpu_ctx* ctx, const MoviGpuIndexDesc* desc, void* d_rows);
/* A second context on the SAME GPU that queries the index another context has uploaded (no copy; the owner must outlive it):
 * a context serves one host thread at a time, two contexts let two batches / file chunks overlap their copies and walks. */
int  movi_gpu_share_index(movi_gpu_ctx* ctx, const movi_gpu_ctx* owner);
/* device pointer of the resident 8-byte row table in the DEVICE row format (movi_b200/csrc/dev_index.cuh; read-only;
 * NULL without an index or with a compact layout) */
const void* movi_gpu_index_device_rows(const movi_gpu_ctx* ctx);
/* rows [first, first + count) of the resident 8-byte table, converted back to the reference's regular-thresholds MoveRow
 * (include/move_row.hpp:131-142) in host memory -- the inverse of what upload / adopt do.  Tests and generators use it to
 * hand a device-built table to the CPU side. */
int  movi_gpu_index_export_rows(movi_gpu_ctx* ctx, uint64_t first, uint64_t count, void* host_rows);
/* MOVI_GPU_LAYOUT_EXPANDED8 or MOVI_GPU_LAYOUT_COMPACT of the resident index (-1 without one) */
int  movi_gpu_index_device_layout(const movi_gpu_ctx* ctx);
/* bytes of HBM held by the uploaded index */
uint64_t movi_gpu_index_device_bytes(const movi_gpu_ctx* ctx);

/* replaces: MoveStructure::query_pml, one call per batch instead of per read (src/move_structure_query.cpp:234-474).
 * HOST buffers.  out_ml has read_offsets[n_reads] entries.  Copies in, runs, copies out (pipelined). */
int  movi_gpu_query_pml(movi_gpu_ctx* ctx, const char* bases, const uint64_t* read_offsets, uint64_t n_reads,
                        uint32_t flags, uint16_t* out_ml);
/* replaces: query_pml / query_zml for batches of SHORT reads (every read <= 255 bases, e.g. 150 bp Illumina).  A matching
 * length never exceeds the number of bases consumed (match_len starts at 0 and grows by one per base,
 * src/move_structure_query.cpp:237-275, :696-766), so it fits one byte: results come back as uint8 -- 1 byte per base over
 * PCIe instead of 2, the copy that bounds the host-buffer path -- and widen in the caller's MoveQuery::add_ml loop
 * (INTEGRATION.md).  Same values, same order as movi_gpu_query_pml / _zml.  MOVI_GPU_ERR_ARG if a read is longer. */
int  movi_gpu_query_ml8(movi_gpu_ctx* ctx, int zml, const char* bases, const uint64_t* read_offsets, uint64_t n_reads,
                        uint32_t flags, uint8_t* out_ml8);
/* replaces: query_pml / query_zml FOLLOWED BY Classifier::classify (handle_pml_zml, src/movi.cpp:23-60; src/classifier.cpp:99-143)
 * for the --classify / --filter workflows: the matching lengths are binned where they sit in HBM and only the three report
 * columns per read come back (bins_above, bins_below, sum of bin maxima; found = above / (above + below) > 0.5), i.e. 16 bytes
 * per read instead of 2 bytes per base over PCIe.  zml != 0 classifies ZMLs.  out_ml may be NULL (e.g. --filter, which writes
 * no .bpf); if given it also receives the matching lengths exactly as movi_gpu_query_pml / _zml would return them. */
int  movi_gpu_query_classify(movi_gpu_ctx* ctx, int zml, const char* bases, const uint64_t* read_offsets, uint64_t n_reads,
                             uint32_t flags, uint32_t bin_width, uint32_t max_value_thr,
                             uint32_t* out_bins_above, uint32_t* out_bins_below, uint64_t* out_sum_max, uint16_t* out_ml);
/* replaces: MoveStructure::query_zml (src/move_structure_query.cpp:690-786) */
int  movi_gpu_query_zml(movi_gpu_ctx* ctx, const char* bases, const uint64_t* read_offsets, uint64_t n_reads,
                        uint32_t flags, uint16_t* out_ml);
/* replaces: MoveStructure::query_backward_search + MoveInterval::count (src/move_structure_search.cpp:340-352,
 * include/move_intervals.hpp:47-58).  out_pos_on_r[i] is pos_on_r as the reference leaves it (the text
 * output prints len - pos_on_r, src/utils.cpp:248-256). */
int  movi_gpu_query_count(movi_gpu_ctx* ctx, const char* bases, const uint64_t* read_offsets, uint64_t n_reads,
                          uint32_t flags, uint64_t* out_count, int32_t* out_pos_on_r);

/* PML (zml = 0) or ZML (zml = 1) with DEVICE-SIDE .bpf EMISSION: out_image receives the records
 *     { u16 id_len; char id[id_len]; u64 L; u16 ml[L] }
 * of the batch back to back -- exactly the bytes output_base_stats / output_binary (src/utils.cpp:202-246) append to
 * <prefix>.pml.bpf after the 12-byte BPFHeader (include/utils.hpp:64-86) -- assembled on the GPU per pipeline slice, so the host
 * does one write() per batch and no per-read work.  ids = the read ids concatenated (the reference's id rule: through and
 * including the first blank, src/batch_loader.cpp:117-119), id_offsets = n_reads + 1 prefix offsets into it.
 * movi_gpu_bpf_image_bytes = 10 * n + id bytes + 2 * bases: the size out_image must have. */
uint64_t movi_gpu_bpf_image_bytes(const uint64_t* read_offsets, const uint64_t* id_offsets, uint64_t n_reads);
int  movi_gpu_query_bpf(movi_gpu_ctx* ctx, int zml, const char* bases, const uint64_t* read_offsets, uint64_t n_reads, uint32_t flags,
                        const char* ids, const uint64_t* id_offsets, uint8_t* out_image, uint64_t image_capacity);
/* The assembly alone on device buffers (d_ml as written by a *_device query: d_read_offsets relative to it). */
int  movi_gpu_bpf_assemble_device(movi_gpu_ctx* ctx, const uint16_t* d_ml, const uint64_t* d_read_offsets, uint64_t n_reads,
                                  const char* d_ids, const uint64_t* d_id_offsets, uint8_t* d_image, void* cuda_stream);

/* DEVICE-SIDE INGEST: a chunk of a FASTA / FASTQ read file (raw bytes, starting at a record boundary; < 4 GiB) is copied to the GPU
 * and parsed there -- BatchLoader::loadBatch + grabNextRead, src/batch_loader.cpp:26-143: multi-line FASTA, four-line FASTQ,
 * trailing white space dropped, ids through and including their first blank -- walked (PML or ZML) and, when out_image is not
 * NULL, returned as the .bpf records of its reads (as movi_gpu_query_bpf).  final_chunk = 0: the chunk ends somewhere inside the
 * file; its last, possibly incomplete record is left to the next chunk: res->consumed_bytes says where that one starts.
 * res->parse_flags != 0 (MOVI_GPU_PARSE_*): the chunk holds something the reference rejects or stops at; nothing was queried and
 * the caller should hand the input to the host parser (movi_host_reader_*), which reproduces the reference's behaviour. */
typedef struct MoviGpuChunkResult {
    uint64_t n_reads, n_bases, id_bytes, consumed_bytes, image_bytes;
    uint32_t parse_flags, reserved;
} MoviGpuChunkResult;
#define MOVI_GPU_PARSE_BAD_HEADER   1u   /* a record does not start with '>' / '@' */
#define MOVI_GPU_PARSE_NO_ID        2u   /* "header line is missing an id" */
#define MOVI_GPU_PARSE_EMPTY_HEADER 4u   /* an empty line where a header is due: the reference stops reading there */
#define MOVI_GPU_PARSE_OVERFLOW     8u   /* an id longer than 65535 bytes */
int  movi_gpu_query_file_chunk(movi_gpu_ctx* ctx, int zml, const char* raw, uint64_t n_bytes, int fastq, int final_chunk, uint32_t flags,
                               uint8_t* out_image, uint64_t image_capacity, MoviGpuChunkResult* res);
/* The same in two steps, for callers that place the records of several chunks (and GPUs) in one file: _parse copies the chunk up
 * and parses it (res->image_bytes = the size of its records); _query walks the reads parsed last and delivers the records to
 * out_image (may be NULL) and / or straight into the file out_fd at file_offset (-1: no file), through a ring of pinned buffers
 * -- no host buffer of the size of the image is needed.  A file that already extends to file_offset + image_bytes (ftruncate) is
 * filled through a shared mapping by several threads; otherwise with positional writes. */
int  movi_gpu_file_chunk_parse(movi_gpu_ctx* ctx, const char* raw, uint64_t n_bytes, int fastq, int final_chunk, MoviGpuChunkResult* res);
int  movi_gpu_file_chunk_query(movi_gpu_ctx* ctx, int zml, uint32_t flags, uint8_t* out_image, uint64_t image_capacity, int out_fd,
                               uint64_t file_offset);
/* The chunk parsed last, for the --classify / --filter workflows (handle_pml_zml, src/movi.cpp:23-60): walked, binned on the device
 * (Classifier::classify, src/classifier.cpp:99-143; the three report columns per read come back, as from movi_gpu_query_classify) and,
 * when a sink is given, also delivered as .bpf records like movi_gpu_file_chunk_query (the report of --classify comes with its .bpf). */
int  movi_gpu_file_chunk_classify(movi_gpu_ctx* ctx, int zml, uint32_t flags, uint32_t bin_width, uint32_t max_value_thr,
                                  uint32_t* out_bins_above, uint32_t* out_bins_below, uint64_t* out_sum_max,
                                  uint8_t* out_image, uint64_t image_capacity, int out_fd, uint64_t file_offset);
/* The chunk parsed last, for --filter [-v] (src/movi.cpp:36-47): walked, binned, and the echo of the reads Classifier::classify keeps --
 * ">" id "\n" sequence "\n" each (output_read, src/utils.cpp:291-294) -- assembled on the device; *text_bytes = its size.  _deliver
 * fetches it into out_text (may be NULL) and / or writes it to out_fd with write() calls in order (stdout may be a pipe). */
int  movi_gpu_file_chunk_filter(movi_gpu_ctx* ctx, int zml, uint32_t flags, uint32_t bin_width, uint32_t max_value_thr, int invert,
                                uint64_t* text_bytes);
int  movi_gpu_file_chunk_filter_deliver(movi_gpu_ctx* ctx, uint8_t* out_text, uint64_t capacity, int out_fd);
/* The reads of the chunk parsed last copied back to the host: n_bases bytes, n_reads + 1 offsets, id_bytes bytes, n_reads + 1 id
 * offsets (sizes in the MoviGpuChunkResult of the parse; any pointer may be NULL) -- the ids a report prints, the sequences --filter echoes. */
int  movi_gpu_file_chunk_fetch(movi_gpu_ctx* ctx, char* out_bases, uint64_t* out_offsets, char* out_ids, uint64_t* out_id_offsets);
/* The parse alone, results copied back to the host (tests, tools). */
int  movi_gpu_parse_chunk(movi_gpu_ctx* ctx, const char* raw, uint64_t n_bytes, int fastq, int final_chunk,
                          char* out_bases, uint64_t bases_capacity, uint64_t* out_offsets, char* out_ids, uint64_t ids_capacity,
                          uint64_t* out_id_offsets, uint64_t reads_capacity, MoviGpuChunkResult* res);

/* Same queries on DEVICE buffers (already resident in HBM), enqueued on `cuda_stream` (a cudaStream_t,
 * NULL = default stream) without synchronising: for callers that pipeline their own copies.
 * d_bases and d_out_ml must be 32-byte aligned and d_bases readable up to the next multiple of 32 bytes (the
 * kernels move whole aligned 32-byte sectors); read_offsets are relative to d_bases; total bases per call < 2^32 (a
 * larger batch is refused: the kernels flag it and movi_gpu_check_device_status returns MOVI_GPU_ERR_ARG; the host-buffer
 * entry points slice such batches themselves).  d_out_ml has read_offsets[n_reads] entries.  At most 64 *_device calls
 * of one context may be in flight at a time (each takes one of a ring of 64 work counters). */
int  movi_gpu_query_pml_device(movi_gpu_ctx* ctx, const char* d_bases, const uint64_t* d_read_offsets, uint64_t n_reads,
                               uint32_t flags, uint16_t* d_out_ml, void* cuda_stream);
int  movi_gpu_query_zml_device(movi_gpu_ctx* ctx, const char* d_bases, const uint64_t* d_read_offsets, uint64_t n_reads,
                               uint32_t flags, uint16_t* d_out_ml, void* cuda_stream);
int  movi_gpu_query_count_device(movi_gpu_ctx* ctx, const char* d_bases, const uint64_t* d_read_offsets, uint64_t n_reads,
                                 uint32_t flags, uint64_t* d_out_count, int32_t* d_out_pos_on_r, void* cuda_stream);
/* After a *_device call has completed (caller synchronised the stream): MOVI_GPU_OK, MOVI_GPU_ERR_ARG for a batch of
 * 2^32 bases or more, or MOVI_GPU_ERR_QUERY if any read hit a "should not happen" state.  Clears the flag. */
int  movi_gpu_check_device_status(movi_gpu_ctx* ctx);

/* Binary classification on the device from PMLs already in HBM (Classifier::classify, src/classifier.cpp:99-143):
 * per read bins_above / bins_below / sum of bin maxima; found = above/(above+below) > 0.5. */
int  movi_gpu_classify_device(movi_gpu_ctx* ctx, const uint16_t* d_ml, const uint64_t* d_read_offsets, uint64_t n_reads,
                              uint32_t bin_width, uint32_t max_value_thr,
                              uint32_t* d_bins_above, uint32_t* d_bins_below, uint64_t* d_sum_max, void* cuda_stream);

/* ---- Movi Color: multi-class classification with PMLs and suffix-array entries.  Threshold index types, expanded
 * device layout.  Results follow the reference's --no-prefetch path (the deterministic one). */

/* What load_color_table leaves in MoveStructure for the default flat colour table (src/movi.cpp:121-152,
 * src/move_structure_io.cpp:587-607, 643-700): HOST pointers, needed during the call only. */
typedef struct MoviGpuColorDesc {
    const uint16_t* flat_colors;         /* flat_colors: {size, documents...} per colour, back to back */
    uint64_t flat_size;                  /* flat_colors.size() */
    const void* doc_set_flat_inds;       /* r MoveTally {u32 right; u8 left} (5 bytes each): the row's offset into flat_colors */
    uint32_t num_species;                /* documents after the taxon-id compression of load_document_info */
    uint32_t reserved;
    const double* log_lens;              /* num_species: log(total length of the species' documents); NULL: no --pvalue-scoring */
} MoviGpuColorDesc;
/* Per read, what the end of query_pml reads to print "<id>,<taxon>,<taxon>" (src/move_structure_query.cpp:439-471):
 * best_doc / second_best_doc = species indices (0xFFFF = none), their counters (uint32 counts, or the bits of the double scores
 * under --pvalue-scoring) and the uint32 sum of the read's matching lengths (PML_mean = sum / length). */
typedef struct MoviGpuColorResult {
    uint32_t best_doc, second_best_doc;
    uint32_t sum_matching_lengths, reserved;
    uint64_t best_raw, second_raw;
} MoviGpuColorResult;
/* replaces: load_color_table (src/movi.cpp:121-152): the colour table goes to HBM next to the uploaded index */
int  movi_gpu_upload_colors(movi_gpu_ctx* ctx, const MoviGpuColorDesc* desc);
/* replaces: MoveStructure::query_pml with is_multi_classify() (src/move_structure_query.cpp:251-260, 373-437), per batch.
 * min_len = --min-len (default 1), pvalue != 0 = --pvalue-scoring.  out_ml (may be NULL) also receives the matching lengths. */
int  movi_gpu_query_multi_classify(movi_gpu_ctx* ctx, const char* bases, const uint64_t* read_offsets, uint64_t n_reads,
                                   uint32_t flags, uint32_t min_len, int pvalue, MoviGpuColorResult* out, uint16_t* out_ml);
/* replaces: deserialize_sampled_SA (src/move_structure_io.cpp:716-735): ssa.movi = sample rate, sampled_SA_entries, all_p (r) */
int  movi_gpu_upload_sampled_sa(movi_gpu_ctx* ctx, uint64_t sample_rate, const uint64_t* sampled_sa, uint64_t n_sampled,
                                const uint64_t* all_p);
/* replaces: query_pml with is_get_sa_entries() (src/move_structure_query.cpp:353-357, get_SA_entries src/move_structure.cpp:35-48):
 * one suffix-array entry per base at the same offsets and in the same order as the matching lengths (out_ml may be NULL). */
int  movi_gpu_query_pml_sa(movi_gpu_ctx* ctx, const char* bases, const uint64_t* read_offsets, uint64_t n_reads,
                           uint32_t flags, uint16_t* out_ml, uint64_t* out_sa);

/* Introspection for benchmarks: kernels launched by this context so far, name and time of the last one. */
uint64_t movi_gpu_launch_count(const movi_gpu_ctx* ctx);
/* Tuning knobs (benchmarks / profiling / tests): "pml_sector" (-1 by table size / 0 / 1: sector-cached row fetch),
 * "blocks_per_sm", "threads_per_block", "force_wide" (64-bit row indices on small tables), "smem_carveout";
 * host-buffer pipeline: "batch_bases" (bases per slice), "pipeline_slots" (2..8 slices in flight), "kernel_streams"
 * (1..4 slices walking concurrently), "size_grid_to_batch" (k >= 1: a small batch gets a grid of k reads per lane; 1 measured best). */
int  movi_gpu_set_option(movi_gpu_ctx* ctx, const char* key, int64_t value);
int64_t movi_gpu_get_option(const movi_gpu_ctx* ctx, const char* key);

/* Pinned host memory for callers that want DMA-speed copies through the host-buffer entry points. */
int  movi_gpu_host_alloc(void** p, uint64_t bytes);
int  movi_gpu_host_free(void* p);

/* Random-access roofline probe: `n_threads` independent chains of `steps` dependent (or independent)
 * 8- or 32-byte loads at pseudo-random row addresses of the uploaded table (or a scratch table of
 * table_bytes when no index is uploaded).  Returns seconds of device time in *seconds. */
int  movi_gpu_gather_probe(movi_gpu_ctx* ctx, uint64_t table_bytes, uint32_t load_bytes, uint32_t dependent,
                           uint32_t steps, uint32_t blocks_per_sm, uint32_t threads_per_block, double* seconds,
                           uint64_t* loads_done);

#ifdef __cplusplus
}
#endif
#endif /* MOVI_GPU_H */
