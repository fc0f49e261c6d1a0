/* movi_oracle.h -- CPU restatement of Movi's PML / ZML / count query path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: a plain-C restatement of the reference
 * algorithm (no GPU, no torch).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product path (movi_b200/csrc,
 * movi_b200/host) never links or calls it and fails loudly without the CUDA library.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks it against the reference's own golden
 * vectors (tests_data/sample.fastq.pmls.sorted, the reads.fasta classification report lines,
 * sample.fasta.pmls.filtered_notfound.sorted) and against outputs of the unmodified reference
 * binaries compiled into oracle/_ref/ (committed under tests/golden/ with the generating script).
 */
#ifndef MOVI_ORACLE_H
#define MOVI_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct movi_oracle_index movi_oracle_index;

/* Parse an index.movi file (modern 48-byte header; modes 6 / 8 / 7).  Returns 0 on success. */
int movi_oracle_load(const char* index_file, movi_oracle_index** out);
void movi_oracle_free(movi_oracle_index* ix);
const char* movi_oracle_last_error(void);

int      movi_oracle_mode(const movi_oracle_index* ix);           /* 6, 8 or 7 */
uint64_t movi_oracle_r(const movi_oracle_index* ix);
uint64_t movi_oracle_n(const movi_oracle_index* ix);
uint64_t movi_oracle_end_bwt_idx(const movi_oracle_index* ix);
int      movi_oracle_sigma(const movi_oracle_index* ix);
int      movi_oracle_has_separators(const movi_oracle_index* ix);

/* Row accessors (reference MoveStructure::get_n/get_offset/get_id/get_char, thresholds bit i). */
uint64_t movi_oracle_get_n(const movi_oracle_index* ix, uint64_t idx);
uint64_t movi_oracle_get_offset(const movi_oracle_index* ix, uint64_t idx);
uint64_t movi_oracle_get_id(const movi_oracle_index* ix, uint64_t idx);
int      movi_oracle_get_c(const movi_oracle_index* ix, uint64_t idx);   /* raw 3-bit field */
int      movi_oracle_get_threshold_bit(const movi_oracle_index* ix, uint64_t idx, int i);

/* One LF step, in place (reference LF_move).  Returns the fast-forward count or -1 on error. */
int64_t movi_oracle_lf_move(const movi_oracle_index* ix, uint64_t* offset, uint64_t* idx);

/* PML of one read (reference query_pml).  out[k] is the k-th value produced, i.e. out[0]
 * belongs to the LAST base.  Returns 0, or <0 when the reference would have thrown. */
int movi_oracle_query_pml(const movi_oracle_index* ix, const char* read, int64_t len, uint16_t* out);
/* --rpml: PML with reposition_randomly (move_structure_query.cpp:604-688) instead of the thresholds */
int movi_oracle_query_rpml(const movi_oracle_index* ix, const char* read, int64_t len, uint16_t* out);
/* ZML of one read (reference query_zml); same orientation. Returns number of values written
 * (== len for every read the reference handles without UB) or <0. */
int64_t movi_oracle_query_zml(const movi_oracle_index* ix, const char* read, int64_t len, uint16_t* out);
/* count query (reference query_backward_search + MoveInterval::count).  pos_on_r as left by the
 * reference (the text prints len - pos_on_r). */
int movi_oracle_query_count(const movi_oracle_index* ix, const char* read, int64_t len,
                            uint64_t* count, int32_t* pos_on_r);

/* Batch forms over concatenated reads (read i = bases[offsets[i] .. offsets[i+1]) ), OpenMP over
 * reads with `threads` threads (<=0: all).  kind: 0 = PML, 1 = ZML, 2 = PML with --rpml.  out uses the same offsets. */
int movi_oracle_query_ml_batch(const movi_oracle_index* ix, int kind, const char* bases,
                               const uint64_t* offsets, uint64_t n_reads, uint16_t* out, int threads);
int movi_oracle_query_count_batch(const movi_oracle_index* ix, const char* bases,
                                  const uint64_t* offsets, uint64_t n_reads,
                                  uint64_t* counts, int32_t* pos_on_r, int threads);

/* Binary classification of one read from its PMLs (reference Classifier::classify).
 * Returns 1 = FOUND, 0 = NOT_PRESENT; fills the report columns. */
int movi_oracle_classify(const uint16_t* ml, uint64_t len, uint64_t bin_width, uint64_t max_value_thr,
                         uint64_t* bins_above, uint64_t* bins_below, uint64_t* sum_max, uint64_t* n_bins);

/* ---- Movi Color: `--multi-classify` on the flat color table and `--sa-entries` (reference --no-prefetch path) */
typedef struct movi_oracle_colors movi_oracle_colors;
typedef struct movi_oracle_ssa movi_oracle_ssa;
int  movi_oracle_colors_load(const char* index_dir, uint64_t r, movi_oracle_colors** out);   /* doc_sets_flat.bin, ref.fa.doc_offsets[, ref.fa.doc_ids] */
void movi_oracle_colors_free(movi_oracle_colors* c);
uint32_t movi_oracle_colors_num_species(const movi_oracle_colors* c);
uint32_t movi_oracle_colors_taxon(const movi_oracle_colors* c, uint32_t species);
/* raw = {best_doc, second_best_doc (65535: none), sum of matching lengths}; score = the two counters (or p-value scores);
 * taxa = the two numbers of the reference's output line "<id>,<taxa[0]>,<taxa[1]>". */
int movi_oracle_query_multi_classify(const movi_oracle_index* ix, const movi_oracle_colors* col, const char* read, int64_t len,
                                     int min_len, int pvalue, uint16_t* out_ml, uint32_t raw[3], double score[2], uint32_t taxa[2]);
int  movi_oracle_ssa_load(const char* index_dir, movi_oracle_ssa** out);                      /* ssa.movi */
void movi_oracle_ssa_free(movi_oracle_ssa* s);
/* the `build-SA` action itself (find_sampled_SA_entries, move_structure_build.cpp:1174-1212) and what it yields */
int  movi_oracle_ssa_build(const movi_oracle_index* ix, uint64_t rate, movi_oracle_ssa** out);
uint64_t movi_oracle_ssa_rate(const movi_oracle_ssa* s);
uint64_t movi_oracle_ssa_samples(const movi_oracle_ssa* s);
uint64_t movi_oracle_ssa_rows(const movi_oracle_ssa* s);
const uint64_t* movi_oracle_ssa_sampled(const movi_oracle_ssa* s);
const uint64_t* movi_oracle_ssa_all_p(const movi_oracle_ssa* s);
int movi_oracle_query_pml_sa(const movi_oracle_index* ix, const movi_oracle_ssa* s, const char* read, int64_t len,
                             uint16_t* out_ml, uint64_t* out_sa);

#ifdef __cplusplus
}
#endif
#endif
