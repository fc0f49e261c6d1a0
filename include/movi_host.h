/* movi_host.h -- C ABI of the host-side helpers around the GPU query path (pure C++ inside, no CUDA).
 *
 * These restate the reference's host code either side of the hot path so that a caller can go from
 * Movi's on-disk files to movi_gpu_* calls and back to Movi's output files:
 *   index.movi reader        MoveStructure::deserialize            src/move_structure_io.cpp:471-511
 *   FASTA / FASTQ batches    BatchLoader::loadBatch / grabNextRead  src/batch_loader.cpp:26-143
 *   .bpf / count / stdout    output_base_stats / output_counts      src/utils.cpp:202-256, 319-384
 *   binary classification    Classifier                              src/classifier.cpp:24-143
 */
#ifndef MOVI_HOST_H
#define MOVI_HOST_H

#include <stdint.h>
#include "movi_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct movi_host_index movi_host_index;
typedef struct movi_host_reads movi_host_reads;
typedef struct movi_host_reader movi_host_reader;

const char* movi_host_last_error(void);

/* index.movi (path may be the index directory or the file).  The file is memory-mapped; the
 * descriptor's pointers stay valid until movi_host_index_close(). */
int  movi_host_index_open(const char* path, movi_host_index** out);
int  movi_host_index_desc(const movi_host_index* ix, MoviGpuIndexDesc* out);
uint64_t movi_host_index_original_r(const movi_host_index* ix);
const char* movi_host_index_type_name(const movi_host_index* ix);   /* "regular-thresholds", ... (src/utils.cpp:10-45) */
void movi_host_index_close(movi_host_index* ix);

/* Whole read file (FASTA, possibly multi-line, or 4-line FASTQ; "-" = stdin) into one concatenated batch.
 * Read ids follow the reference: header from its 2nd character through AND INCLUDING the first
 * blank / tab / CR (src/batch_loader.cpp:117-119).  reverse != 0 reverses every read (--reverse). */
int  movi_host_reads_load(const char* path, int reverse, movi_host_reads** out);
/* The same parser batch by batch (the reference's BatchLoader hands out batches too, src/batch_loader.cpp:26-75): every call
 * returns the next whole records holding about max_bases bases (at least one record), *out = NULL at the end of the input.
 * The file stays memory-mapped, so inputs larger than RAM stream through. */
int  movi_host_reader_open(const char* path, int reverse, movi_host_reader** out);
int  movi_host_reader_next(movi_host_reader* r, uint64_t max_bases, movi_host_reads** out);
/* The reference's loadBatch(input, num_bases, min_number_of_reads) (src/batch_loader.cpp:26-75): the batch grows until it
 * holds max_bases bases AND min_reads reads -- a GPU wants one read per lane in flight however long the reads are -- but never
 * beyond cap_bases bases (0 = no cap; the bound on host memory per batch). */
int  movi_host_reader_next_reads(movi_host_reader* r, uint64_t max_bases, uint64_t min_reads, uint64_t cap_bases, movi_host_reads** out);
/* The whole input as the reader holds it (memory-mapped file / slurped stdin / inflated gzip) and its format: what the device-side
 * ingest (movi_gpu_query_file_chunk) is fed from.  The pointer stays valid until movi_host_reader_close. */
int  movi_host_reader_raw(const movi_host_reader* r, const char** data, uint64_t* size, int* fastq);
/* First record boundary at or after `from` (the offset of a '>' / '@' that starts a record; `size` if there is none): FASTA = a
 * line starting with '>'; FASTQ = a line starting with '@' whose next-but-one line starts with '+' (a quality line may start
 * with '@' too, but then the line two further down is a sequence, never a '+' line). */
uint64_t movi_host_record_boundary(const char* data, uint64_t size, uint64_t from, int fastq);
/* Continue at byte `at` of the input (a record boundary found by movi_host_record_boundary / consumed by the device-side parser). */
int  movi_host_reader_seek(movi_host_reader* r, uint64_t at);
void movi_host_reader_close(movi_host_reader* r);
/* A batch from parts (copies): what the device-side parser returns through movi_gpu_file_chunk_fetch.  bases may be NULL when only the
 * ids and lengths are needed (a classification report); offsets and id_offsets start at 0. */
int  movi_host_reads_adopt(const char* bases, const uint64_t* offsets, uint64_t n_reads, const char* ids, const uint64_t* id_offsets,
                           movi_host_reads** out);
uint64_t movi_host_reads_count(const movi_host_reads* rd);
const char* movi_host_reads_bases(const movi_host_reads* rd);          /* concatenated */
const uint64_t* movi_host_reads_offsets(const movi_host_reads* rd);    /* count + 1 */
const char* movi_host_reads_id(const movi_host_reads* rd, uint64_t i, uint32_t* len);
const char* movi_host_reads_ids(const movi_host_reads* rd);             /* the ids concatenated ... */
const uint64_t* movi_host_reads_id_offsets(const movi_host_reads* rd);  /* ... and their count + 1 offsets (movi_gpu_query_bpf takes both) */
void movi_host_reads_free(movi_host_reads* rd);

/* movi.<q>.nulldb -> Classifier::initialize_report_file's threshold: max(percentile, 3) + 1
 * (src/classifier.cpp:24-32, src/emperical_null_database.cpp:106-126) */
int  movi_host_nulldb_threshold(const char* index_dir, const char* query_type, uint64_t* max_value_thr);

/* Writers.  `ml` holds the matching lengths of read i at offsets[i]..offsets[i+1] in storage order
 * (element 0 = last base), exactly what movi_gpu_query_pml / _zml return. */
int  movi_host_write_bpf(const char* file, const movi_host_reads* rd, const uint16_t* ml);          /* src/utils.cpp:202-246 + BPFHeader */
/* the same file written batch by batch: begin returns the descriptor (-1 on error) after writing the header */
int  movi_host_bpf_begin(const char* file);
int  movi_host_bpf_append(int fd, const movi_host_reads* rd, const uint16_t* ml);
/* Append records already assembled on the device (movi_gpu_query_bpf): one image per batch, written by several threads at once. */
int  movi_host_bpf_append_image(int fd, const uint8_t* image, uint64_t bytes);
int  movi_host_bpf_end(int fd);
int  movi_host_write_ml_text(int fd, const movi_host_reads* rd, const uint16_t* ml);                /* --stdout form */
int  movi_host_write_counts(int fd, const movi_host_reads* rd, const uint64_t* counts, const int32_t* pos_on_r);   /* src/utils.cpp:248-256 */
/* Classification report / filter (src/classifier.cpp:99-143, src/movi.cpp:36-47).
 * mode: 0 = report (with header) to fd, 1 = filter: print FOUND reads, 2 = filter --invert; | MOVI_HOST_CLASSIFY_CONTINUE for the
 * second and later batches of one report (no header).  found_out may be NULL. */
#define MOVI_HOST_CLASSIFY_CONTINUE 4
int  movi_host_classify(int fd, int mode, const movi_host_reads* rd, const uint16_t* ml, uint64_t bin_width,
                        uint64_t max_value_thr, uint8_t* found_out);

/* Same report / filter output from the per-read columns computed on the device (movi_gpu_query_classify). */
int  movi_host_classify_bins(int fd, int mode, const movi_host_reads* rd, uint64_t max_value_thr, const uint32_t* bins_above,
                             const uint32_t* bins_below, const uint64_t* sum_max, uint8_t* found_out);

/* ---- Movi Color */
typedef struct movi_host_colors movi_host_colors;
typedef struct movi_host_ssa movi_host_ssa;
/* load_color_table (reference src/movi.cpp:121-152) for the default flat colour table: <index_dir>/doc_sets_flat.bin (r = rows of
 * the index), ref.fa.doc_offsets and, if present, ref.fa.doc_ids (taxon id per document; src/move_structure_io.cpp:587-607, 643-700) */
int  movi_host_colors_open(const char* index_dir, uint64_t r, movi_host_colors** out);
/* kind 0: the same; kind 1: <index_dir>/doc_sets.bin, the vectors-of-vectors table `--color-vectors` names (src/move_structure_io.cpp:571-641),
 * flattened on load (the device takes one form) */
int  movi_host_colors_open_kind(const char* index_dir, uint64_t r, int kind, movi_host_colors** out);
int  movi_host_colors_desc(const movi_host_colors* c, MoviGpuColorDesc* out);     /* pointers valid until _close */
uint32_t movi_host_colors_num_species(const movi_host_colors* c);
uint32_t movi_host_colors_taxon(const movi_host_colors* c, uint32_t species);     /* to_taxon_id */
void movi_host_colors_close(movi_host_colors* c);
/* the lines "<id>,<taxon>,<taxon>" query_pml writes to the -o file under --multi-classify (src/move_structure_query.cpp:439-471),
 * from the per-read results of movi_gpu_query_multi_classify */
int  movi_host_write_multi_classify(int fd, const movi_host_reads* rd, const movi_host_colors* c, const MoviGpuColorResult* res,
                                    uint32_t min_len, int pvalue);
/* <index_dir>/ssa.movi (deserialize_sampled_SA, src/move_structure_io.cpp:716-735): what movi_gpu_upload_sampled_sa takes */
int  movi_host_ssa_open(const char* index_dir, movi_host_ssa** out);
uint64_t movi_host_ssa_rate(const movi_host_ssa* s);
uint64_t movi_host_ssa_rows(const movi_host_ssa* s);
uint64_t movi_host_ssa_samples(const movi_host_ssa* s);
const uint64_t* movi_host_ssa_sampled(const movi_host_ssa* s);
const uint64_t* movi_host_ssa_all_p(const movi_host_ssa* s);
void movi_host_ssa_close(movi_host_ssa* s);
/* <prefix>.sa_entries.bpf records (output_base_stats, DataType::sa_entry; the file has no header: src/utils.cpp:371-375) */
int  movi_host_sa_append(int fd, const movi_host_reads* rd, const uint64_t* sa);

#ifdef __cplusplus
}
#endif
#endif /* MOVI_HOST_H */
