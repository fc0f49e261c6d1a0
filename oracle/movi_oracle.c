/* movi_oracle.c -- CPU restatement of Movi's PML / ZML / count query path (see movi_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: the parity oracle.  Never linked into or called from the product path.
 * Parity status: PINNED (tests/test_oracle.py: reference golden vectors + oracle/_ref outputs).
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.  The
 * rows are read in their ON-DISK packed form, one decoder per layout, exactly like the reference's
 * per-MODE getters; nothing here shares code with the CUDA implementation.
 */
#define _GNU_SOURCE
#include "movi_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MODE_BLOCKED 2              /* "blocked": 6-byte rows + id_blocks, no thresholds, two more id bits (move_row_configs.hpp:54-73) */
#define MODE_REGULAR 3              /* "regular": 8-byte rows without thresholds (utils.hpp:136), 12-bit lengths */
#define MODE_SAMPLED 5              /* "sampled": 3-byte rows + id checkpoints, no thresholds, 10-bit lengths (:107-118) */
#define MODE_REGULAR_THRESHOLDS 6
#define MODE_SAMPLED_THRESHOLDS 7
#define MODE_BLOCKED_THRESHOLDS 8
#define MOVI_MAGIC 0x4D4F5649u      /* include/utils.hpp:29 */
#define SEPARATOR '%'               /* include/commons.hpp:63 */

struct movi_oracle_index {
    int mode;
    int row_bytes;
    uint64_t length, r, original_r, end_bwt_idx;
    uint64_t end_thr[4];            /* end_bwt_idx_thresholds */
    uint64_t alphamap[256];         /* 256 == illegal */
    uint64_t alphamap_size;
    int sigma;
    unsigned char alphabet[8];
    int separators;                 /* use_separator(): move_structure.cpp:547-552 */
    int no_thresholds;              /* NO_THRESHOLDS modes (utils.hpp:146): PML repositions with reposition_randomly */
    uint8_t* rows;                  /* r * row_bytes, packed exactly as on disk */
    void* map; size_t map_len;      /* row tables beyond 1 GiB are mapped from the file instead of copied (human-scale bench tables) */
    /* MODE 7 */
    uint32_t tally_checkpoints;
    uint64_t tally_len;
    uint8_t* tally;                 /* sigma * tally_len * 5 bytes (MoveTally) */
    /* backward-search seeds */
    uint64_t counts[8];
    uint64_t last_runs[9], last_offsets[9], first_runs[9], first_offsets[9];
    /* MODE 8 */
    uint64_t id_blocks_size, block_size;
    uint32_t* id_blocks;            /* sigma * id_blocks_size */
    /* separators */
    uint64_t sep_cnt;  uint16_t* sep_thr;            /* sep_cnt * 4 */
    uint64_t sep_map_size; uint64_t* sep_map;        /* pairs (row, slot), sorted by row */
};

static __thread char g_err[512];
const char* movi_oracle_last_error(void) { return g_err; }
static int fail(const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return -1; }

/* utils.cpp:5-8 */
static const uint32_t alphamap_3[4][4] = {{3, 0, 1, 2}, {0, 3, 1, 2}, {0, 1, 3, 2}, {0, 1, 2, 3}};

/* ---------------------------------------------------------------- index.movi reader
 * move_structure_io.cpp:471-511 (deserialize) and the sub-readers it calls. */
static int rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n ? 0 : -1; }

static int cmp_pair(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : x > y;
}

int movi_oracle_load(const char* path, movi_oracle_index** out) {
    FILE* f = fopen(path, "rb");
    if (!f) return fail("cannot open index file");
    movi_oracle_index* ix = (movi_oracle_index*)calloc(1, sizeof *ix);
    /* MoviHeader, utils.hpp:32-61 (48 bytes with padding); io.cpp:66-109 */
    uint8_t h[48];
    if (rd(f, h, 48)) goto bad;
    uint32_t magic; memcpy(&magic, h, 4);
    if (magic != MOVI_MAGIC) { fail("bad magic"); goto bad2; }
    ix->mode = (int)(signed char)h[7];
    memcpy(&ix->length, h + 16, 8); memcpy(&ix->r, h + 24, 8);
    memcpy(&ix->original_r, h + 32, 8); memcpy(&ix->end_bwt_idx, h + 40, 8);
    switch (ix->mode) {
        case MODE_REGULAR:
        case MODE_REGULAR_THRESHOLDS: ix->row_bytes = 8; break;
        case MODE_BLOCKED:
        case MODE_BLOCKED_THRESHOLDS: ix->row_bytes = 6; break;
        case MODE_SAMPLED:
        case MODE_SAMPLED_THRESHOLDS: ix->row_bytes = 3; break;
        default: fail("unsupported index mode"); goto bad2;
    }
    /* read_basic_index_data, io.cpp:145-184 */
    uint64_t skip[8];
    if (rd(f, ix->end_thr, 32) || rd(f, skip, 32) || rd(f, skip, 32)) goto bad;
    if (rd(f, &ix->alphamap_size, 8) || ix->alphamap_size != 256) goto bad;
    if (rd(f, ix->alphamap, 256 * 8)) goto bad;
    uint64_t sigma;
    if (rd(f, &sigma, 8) || sigma < 1 || sigma > 7) goto bad;
    ix->sigma = (int)sigma;
    if (rd(f, ix->alphabet, sigma)) goto bad;
    uint16_t nt_splitting; uint8_t constant;
    if (rd(f, &nt_splitting, 2) || rd(f, &constant, 1)) goto bad;
    ix->separators = (ix->alphabet[0] == SEPARATOR && ix->sigma == 5);
    ix->no_thresholds = (ix->mode == MODE_REGULAR || ix->mode == MODE_BLOCKED || ix->mode == MODE_SAMPLED);   /* NO_THRESHOLDS, utils.hpp:146 */
    /* read_main_table, io.cpp:361-397 */
    {
        const size_t table = ix->r * (size_t)ix->row_bytes;
        if (table > ((size_t)1 << 30)) {
            const long at = ftell(f);
            if (at < 0) goto bad;
            ix->map_len = (size_t)at + table;
            ix->map = mmap(NULL, ix->map_len, PROT_READ, MAP_PRIVATE, fileno(f), 0);
            if (ix->map == MAP_FAILED) { ix->map = NULL; goto bad; }
            ix->rows = (uint8_t*)ix->map + at;
            if (fseek(f, (long)table, SEEK_CUR)) goto bad;
        } else {
            ix->rows = (uint8_t*)malloc(table + 16);
            if (!ix->rows || rd(f, ix->rows, table)) goto bad;
        }
    }
    /* read_tally_table, io.cpp:338-349 */
    if (ix->mode == MODE_SAMPLED_THRESHOLDS || ix->mode == MODE_SAMPLED) {
        if (rd(f, &ix->tally_checkpoints, 4) || rd(f, &ix->tally_len, 8)) goto bad;
        ix->tally = (uint8_t*)malloc((size_t)ix->sigma * ix->tally_len * 5 + 8);
        if (!ix->tally || rd(f, ix->tally, (size_t)ix->sigma * ix->tally_len * 5)) goto bad;
    }
    /* read_overflow_tables, io.cpp:219-237 (always empty for these modes, but parse generally) */
    uint64_t sz;
    if (rd(f, &sz, 8) || fseek(f, (long)(sz * 8), SEEK_CUR)) goto bad;
    if (rd(f, &sz, 8) || fseek(f, (long)(sz * 8), SEEK_CUR)) goto bad;
    if (rd(f, &sz, 8) || fseek(f, (long)(sz * 8 * (sigma - 1)), SEEK_CUR)) goto bad;
    /* read_counts_data, io.cpp:269-287 */
    if (rd(f, &sz, 8) || sz != sigma || rd(f, ix->counts, sigma * 8)) goto bad;
    if (rd(f, &sz, 8) || sz != sigma + 1) goto bad;
    if (rd(f, ix->last_runs, sz * 8) || rd(f, ix->last_offsets, sz * 8) ||
        rd(f, ix->first_runs, sz * 8) || rd(f, ix->first_offsets, sz * 8)) goto bad;
    /* read_id_blocks, io.cpp:305-324 */
    if (ix->mode == MODE_BLOCKED_THRESHOLDS || ix->mode == MODE_BLOCKED) {
        if (rd(f, &ix->id_blocks_size, 8)) goto bad;
        if (ix->id_blocks_size) {
            ix->id_blocks = (uint32_t*)malloc((size_t)ix->sigma * ix->id_blocks_size * 4);
            if (!ix->id_blocks || rd(f, ix->id_blocks, (size_t)ix->sigma * ix->id_blocks_size * 4)) goto bad;
        }
        if (rd(f, &ix->block_size, 8)) goto bad;
    }
    /* read_separators_thresholds, io.cpp:415-433 (#if USE_THRESHOLDS, io.cpp:504-508) */
    if (ix->separators && !ix->no_thresholds) {
        if (rd(f, &ix->sep_cnt, 8)) goto bad;
        ix->sep_thr = (uint16_t*)malloc(ix->sep_cnt * 8 + 8);
        if (rd(f, ix->sep_thr, ix->sep_cnt * 8)) goto bad;
        if (rd(f, &ix->sep_map_size, 8)) goto bad;
        ix->sep_map = (uint64_t*)malloc(ix->sep_map_size * 16 + 16);
        if (rd(f, ix->sep_map, ix->sep_map_size * 16)) goto bad;
        qsort(ix->sep_map, ix->sep_map_size, 16, cmp_pair);
    }
    fclose(f);
    *out = ix;
    return 0;
bad:
    fail("truncated or malformed index file");
bad2:
    fclose(f);
    movi_oracle_free(ix);
    return -1;
}

void movi_oracle_free(movi_oracle_index* ix) {
    if (!ix) return;
    if (ix->map) munmap(ix->map, ix->map_len); else free(ix->rows);
    free(ix->tally); free(ix->id_blocks); free(ix->sep_thr); free(ix->sep_map);
    free(ix);
}

int      movi_oracle_mode(const movi_oracle_index* ix) { return ix->mode; }
uint64_t movi_oracle_r(const movi_oracle_index* ix) { return ix->r; }
uint64_t movi_oracle_n(const movi_oracle_index* ix) { return ix->length; }
uint64_t movi_oracle_end_bwt_idx(const movi_oracle_index* ix) { return ix->end_bwt_idx; }
int      movi_oracle_sigma(const movi_oracle_index* ix) { return ix->sigma; }
int      movi_oracle_has_separators(const movi_oracle_index* ix) { return ix->separators; }

/* ---------------------------------------------------------------- packed row decode
 * MODE 6: move_row.hpp:131-142 (fields), 231-258 (getters), 304-317 (thresholds); move_row_configs.hpp:34-51
 * MODE 8: move_row.hpp:128-130,139-142, 260-302, 319-332; move_row_configs.hpp:76-104
 * MODE 7: move_row.hpp:122-126, 208-222, 334-347; move_row_configs.hpp:120-136 */
static inline uint16_t ld16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline uint32_t ld32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

static inline uint64_t row_n(const movi_oracle_index* ix, uint64_t i) {
    const uint8_t* p = ix->rows + i * (size_t)ix->row_bytes;
    switch (ix->mode) {
        case 3: return ld16(p + 4) & 0xFFFu;                              /* n[0:11], move_row_configs.hpp:20-31 */
        case 6: return ld16(p + 4) & 0x7FFu;                              /* n[0:10] */
        case 2:
        case 8: return ld16(p + 2) & 0x3FFu;                              /* n[0:9] */
        case 5: return (uint64_t)p[0] | ((uint64_t)((p[2] >> 2) & 3) << 8);  /* n | c[2:4]<<8, move_row.hpp:208-211 with LENGTH_MASK_BITS 2 */
        default: return (uint64_t)p[0] | ((uint64_t)((p[2] >> 1) & 1) << 8); /* n | c[1]<<8 */
    }
}
static inline uint64_t row_offset(const movi_oracle_index* ix, uint64_t i) {
    const uint8_t* p = ix->rows + i * (size_t)ix->row_bytes;
    switch (ix->mode) {
        case 3: return ld16(p + 6) & 0xFFFu;
        case 6: return ld16(p + 6) & 0x7FFu;
        case 2:
        case 8: return ld16(p + 4) & 0x3FFu;
        case 5: return (uint64_t)p[1] | ((uint64_t)(p[2] & 3) << 8);
        default: return (uint64_t)p[1] | ((uint64_t)(p[2] & 1) << 8);
    }
}
static inline int row_c(const movi_oracle_index* ix, uint64_t i) {
    const uint8_t* p = ix->rows + i * (size_t)ix->row_bytes;
    switch (ix->mode) {
        case 3:                                                           /* REGULAR_MODES share the getters, move_row.hpp:231-258 */
        case 6: return (ld16(p + 4) >> 13) & 7;
        case 2:
        case 8: return (ld16(p + 4) >> 10) & 7;
        case 5: return (p[2] >> 4) & 15;                                  /* C_BITS 4, SHIFT_C 4 */
        default: return (p[2] >> 2) & 7;
    }
}
static inline int row_thr_bit(const movi_oracle_index* ix, uint64_t i, int k) {
    const uint8_t* p = ix->rows + i * (size_t)ix->row_bytes;
    switch (ix->mode) {
        case 6: return k == 0 ? (ld16(p + 6) >> 11) & 1 : k == 1 ? (ld16(p + 4) >> 11) & 1 : (ld16(p + 4) >> 12) & 1;
        case 8: return (ld16(p + 4) >> (13 + k)) & 1;
        default: return (p[2] >> (5 + k)) & 1;
    }
}
/* the id FIELD of the row (MoveRow::get_id), not MoveStructure::get_id */
static inline uint64_t row_id_field(const movi_oracle_index* ix, uint64_t i) {
    const uint8_t* p = ix->rows + i * (size_t)ix->row_bytes;
    if (ix->mode == 6 || ix->mode == 3) return (uint64_t)ld32(p) | ((uint64_t)((ld16(p + 6) >> 12) & 0xF) << 32);
    /* MODE 8: id16 | n[10:15] << 16 (ID_SIG_BITS2 == 0); MODE 2: two more bits from offset[14:15] at bit 22 (move_row.hpp:267-286) */
    uint64_t id = (uint64_t)ld16(p) | ((uint64_t)((ld16(p + 2) >> 10) & 0x3F) << 16);
    if (ix->mode == 2) id |= (uint64_t)((ld16(p + 4) >> 14) & 3) << 22;
    return id;
}
/* move_structure.cpp:288-293 get_char: '$' for the end row.  -1 stands for '$'. */
static inline int get_char(const movi_oracle_index* ix, uint64_t i) {
    return i == ix->end_bwt_idx ? -1 : (int)ix->alphabet[row_c(ix, i)];
}
/* alphabet[rlbwt[idx].get_c()] as used by query_pml / reposition_* (the '$' row decodes as alphabet[0]) */
static inline int raw_char(const movi_oracle_index* ix, uint64_t i) { return (int)ix->alphabet[row_c(ix, i)]; }

static inline uint64_t tally_get(const movi_oracle_index* ix, int c, uint64_t k) {   /* MoveTally::get, move_row.hpp:28-38 */
    const uint8_t* p = ix->tally + ((size_t)c * ix->tally_len + k) * 5;
    return (uint64_t)ld32(p) | ((uint64_t)p[4] << 32);
}

/* MoveStructure::get_id, move_structure.cpp:91-286 */
static int get_id(const movi_oracle_index* ix, uint64_t idx, uint64_t* out) {
    if (ix->mode == 6 || ix->mode == 3) { *out = row_id_field(ix, idx); return 0; }   /* :92-94 */
    if (ix->mode == 8 || ix->mode == 2) {                                         /* :95-102 */
        if (idx != ix->end_bwt_idx) {
            uint64_t block = idx / ix->block_size;
            int c = row_c(ix, idx);
            *out = row_id_field(ix, idx) + (uint64_t)ix->id_blocks[(size_t)c * ix->id_blocks_size + block] + ix->first_runs[c + 1];
        } else {
            *out = row_id_field(ix, idx);
        }
        return 0;
    }
    /* MODE 7, :103-219 (forward direction is always taken, :147-148) */
    const uint64_t cp = ix->tally_checkpoints, r = ix->r;
    if (idx == ix->end_bwt_idx) { *out = 0; return 0; }
    int ci = row_c(ix, idx);
    uint64_t tally_a = idx / cp;
    if (idx == r - 1) { *out = tally_get(ix, ci, ix->tally_len - 1); return 0; }
    if (idx % cp == 0) { *out = tally_get(ix, ci, tally_a); return 0; }
    uint64_t tally_b = tally_a + 1;
    uint64_t next_cp = tally_b * cp;
    if (tally_b * cp >= r) next_cp = r - 1;
    uint64_t rows_until = 0;
    uint64_t id = tally_get(ix, ci, tally_b);
    uint64_t last_id = r;
    for (uint64_t i = idx; i < next_cp; i++) {
        if (get_char(ix, i) == get_char(ix, idx)) { rows_until += row_n(ix, i); last_id = i; }
    }
    if (last_id == idx && get_char(ix, idx) != get_char(ix, next_cp)) { *out = id; return 0; }
    if (last_id == r) return fail("[get id - tally mode] last_id should never be equal to r.");
    uint64_t off = row_offset(ix, next_cp);
    if (get_char(ix, idx) != get_char(ix, next_cp)) {
        rows_until -= row_n(ix, last_id);
        off = row_offset(ix, last_id);
    }
    if (id >= r || off >= row_n(ix, id)) return fail("[get id - tally mode] offset >= n(id)");
    if (off >= rows_until) { *out = id; return 0; }
    rows_until -= (off + 1);
    id -= 1;
    while (rows_until != 0) {
        if (rows_until >= row_n(ix, id)) { rows_until -= row_n(ix, id); id -= 1; }
        else rows_until = 0;
    }
    *out = id;
    return 0;
}

uint64_t movi_oracle_get_n(const movi_oracle_index* ix, uint64_t idx) { return row_n(ix, idx); }
uint64_t movi_oracle_get_offset(const movi_oracle_index* ix, uint64_t idx) { return row_offset(ix, idx); }
int      movi_oracle_get_c(const movi_oracle_index* ix, uint64_t idx) { return row_c(ix, idx); }
int      movi_oracle_get_threshold_bit(const movi_oracle_index* ix, uint64_t idx, int i) { return row_thr_bit(ix, idx, i); }
uint64_t movi_oracle_get_id(const movi_oracle_index* ix, uint64_t idx) {
    uint64_t id = ~0ULL;
    if (get_id(ix, idx, &id)) return ~0ULL;
    return id;
}

/* LF_move (move_structure.cpp:59-87) with fast_forward (:524-545) inlined. */
static int64_t lf_move(const movi_oracle_index* ix, uint64_t* offset, uint64_t* i) {
    uint64_t idx;
    if (get_id(ix, *i, &idx)) return -1;
    if (idx >= ix->r) return fail("[LF_move] This should not happen.");
    *offset = row_offset(ix, *i) + *offset;
    uint64_t ff = 0;
    if (idx < ix->r - 1 && *offset >= row_n(ix, idx)) {
        while (idx < ix->r - 1 && *offset >= row_n(ix, idx)) { *offset -= row_n(ix, idx); idx += 1; ff++; }
        if (ff >= 65535) return fail("[LF_move] Number of fast forwards for a query was greater than 2^16");
    }
    *i = idx;
    return (int64_t)ff;
}
int64_t movi_oracle_lf_move(const movi_oracle_index* ix, uint64_t* offset, uint64_t* idx) { return lf_move(ix, offset, idx); }

/* check_alphabet, move_structure.cpp:383-397 (ignore_illegal_chars == 0) */
static inline int check_alphabet(const movi_oracle_index* ix, unsigned char c) {
    if (ix->separators && c == SEPARATOR) return 0;
    return ix->alphamap[c] != ix->alphamap_size;
}

/* reposition_up / reposition_down, move_structure_query.cpp:188-232 */
static uint64_t reposition_up(const movi_oracle_index* ix, uint64_t idx, int c) {
    if (idx == 0) return ix->r;
    int rc = raw_char(ix, idx);
    while (idx > 0 && rc != c) { idx -= 1; rc = raw_char(ix, idx); }
    return rc == c ? idx : ix->r;
}
static uint64_t reposition_down(const movi_oracle_index* ix, uint64_t idx, int c) {
    if (idx == ix->r - 1) return ix->r;
    int rc = raw_char(ix, idx);
    while (idx < ix->r - 1 && rc != c) { idx += 1; rc = raw_char(ix, idx); }
    return rc == c ? idx : ix->r;
}

static int sep_slot(const movi_oracle_index* ix, uint64_t row, uint64_t* slot) {
    uint64_t lo = 0, hi = ix->sep_map_size;
    while (lo < hi) {
        uint64_t mid = (lo + hi) / 2;
        if (ix->sep_map[2 * mid] < row) lo = mid + 1; else hi = mid;
    }
    if (lo < ix->sep_map_size && ix->sep_map[2 * lo] == row) { *slot = ix->sep_map[2 * lo + 1]; return 0; }
    /* std::unordered_map::operator[] would insert 0 */
    *slot = 0;
    return 0;
}

/* reposition_thresholds, move_structure_query.cpp:513-601.  Returns 1 = up, 0 = down, <0 error. */
static int reposition_thresholds(const movi_oracle_index* ix, uint64_t* idx, uint64_t offset, unsigned char r_char) {
    uint64_t saved = *idx;
    uint64_t ai = ix->alphamap[r_char];
    if (ix->separators) {
        if (ai == 0) return fail("[reposition thresholds] the alphabet index equal to 0 should not happen with separators.");
        ai -= 1;
    }
    int rl_char = raw_char(ix, *idx);
    uint64_t thr;
    if (*idx == ix->end_bwt_idx) {
        thr = ix->end_thr[ai];                                                    /* :534-535 */
    } else if (ix->separators && rl_char == SEPARATOR) {
        uint64_t slot; sep_slot(ix, *idx, &slot);
        thr = ix->sep_thr[slot * 4 + ai];                                         /* :540-541 */
    } else {
        uint64_t k = ix->separators ? alphamap_3[ix->alphamap[rl_char] - 1][ai]
                                    : alphamap_3[ix->alphamap[rl_char]][ai];      /* :545-556 */
        if (k == 3) return fail("[reposition thresholds] alphamap_3 is incorrect");
        thr = row_thr_bit(ix, *idx, (int)k) == 0 ? 0 : row_n(ix, *idx);           /* move_structure.cpp:306-308 */
    }
    if (offset >= thr) {
        *idx = reposition_down(ix, saved, r_char);
        if (*idx >= ix->r || raw_char(ix, *idx) != r_char) return fail("[reposition thresholds] r_char != alphabet[rlbwt[idx].get_c()] (down)");
        return 0;
    } else {
        *idx = reposition_up(ix, saved, r_char);
        if (*idx >= ix->r || raw_char(ix, *idx) != r_char) return fail("[reposition thresholds] r_char != alphabet[rlbwt[idx].get_c()] (up)");
        return 1;
    }
}

static inline uint16_t clamp16(uint64_t v) { return v > 65535 ? 65535 : (uint16_t)v; }   /* move_query.hpp:26-38 */

/* reposition_randomly, move_structure_query.cpp:604-688 (--rpml; the "random" direction is the half of the run the offset
 * lies in -- the call to the random generator is commented out in the reference, :607-608).  Returns 1 = up, 0 = down. */
static int reposition_randomly(const movi_oracle_index* ix, uint64_t* idx, uint64_t offset, unsigned char r_char) {
    const uint64_t saved = *idx;
    int dir = offset * 2 < row_n(ix, *idx) ? 1 : 0;                                   /* :608 */
    if (*idx == ix->r - 1) dir = 1;                                                   /* :616-618 */
    if (*idx == 0) dir = 0;                                                           /* :619-621 */
    int up;
    if ((dir == 1 && *idx > 0) || *idx == ix->r - 1) {                                /* :623 */
        up = 1;
        *idx = reposition_up(ix, saved, r_char);
        if (*idx >= ix->r) {                                                          /* :634-649 */
            up = 0;
            *idx = reposition_down(ix, saved, r_char);
            if (*idx == ix->r) return fail("[reposition randomly] Neither up or down repositioning works.");
        }
    } else {
        up = 0;
        *idx = reposition_down(ix, saved, r_char);
        if (*idx >= ix->r) {                                                          /* :661-674 */
            up = 1;
            *idx = reposition_up(ix, saved, r_char);
            if (*idx == ix->r) return fail("[reposition randomly] Neither up or down repositioning works.");
        }
    }
    if (*idx >= ix->r || raw_char(ix, *idx) != r_char) return fail("[reposition randomly] This should never happen.");
    return up;
}

static int query_pml_impl(const movi_oracle_index* ix, const char* read, int64_t len, uint16_t* out, int rpml);
/* query_pml, move_structure_query.cpp:234-474 (multi-class / logs / SA branches are off) */
int movi_oracle_query_pml(const movi_oracle_index* ix, const char* read, int64_t len, uint16_t* out) { return query_pml_impl(ix, read, len, out, 0); }
/* the same with --rpml: Case 2 repositions with reposition_randomly (:304-311) */
int movi_oracle_query_rpml(const movi_oracle_index* ix, const char* read, int64_t len, uint16_t* out) { return query_pml_impl(ix, read, len, out, 1); }
static int query_pml_impl(const movi_oracle_index* ix, const char* read, int64_t len, uint16_t* out, int rpml) {
    if (ix->no_thresholds) rpml = 1;                           /* move_structure_query.cpp:308-311 */
    int64_t pos = len - 1;
    uint64_t idx = ix->r - 1;
    uint64_t offset = row_n(ix, idx) - 1;
    uint64_t match_len = 0;
    uint64_t k = 0;
    while (pos > -1) {
        unsigned char rc = (unsigned char)read[pos];
        int row_ch = raw_char(ix, idx);
        if (!check_alphabet(ix, rc)) {
            match_len = 0;
        } else if (row_ch == (int)rc) {
            match_len += 1;
        } else {
            int up = rpml ? reposition_randomly(ix, &idx, offset, rc) : reposition_thresholds(ix, &idx, offset, rc);
            if (up < 0) return -1;
            match_len = 0;
            offset = up ? row_n(ix, idx) - 1 : 0;
        }
        out[k++] = clamp16(match_len);
        pos -= 1;
        if (lf_move(ix, &offset, &idx) < 0) return -1;
    }
    return 0;
}

/* ---------------------------------------------------------------- backward search */
typedef struct { uint64_t rs, os, re, oe; } interval_t;                           /* move_intervals.hpp:10-94 */
static inline int iv_empty(const interval_t* v) {
    return !((v->rs < v->re) || (v->rs == v->re && v->os <= v->oe));               /* :43-45 */
}
static inline void iv_make_empty(interval_t* v) { v->rs = 1; v->os = 0; v->re = 0; v->oe = 0; }   /* :36-41 */
static uint64_t iv_count(const movi_oracle_index* ix, const interval_t* v) {      /* :47-58 */
    if (v->rs == v->re) return v->oe - v->os + 1;
    uint64_t c = (row_n(ix, v->rs) - v->os) + (v->oe + 1);
    for (uint64_t k = v->rs + 1; k < v->re; k++) c += row_n(ix, k);
    return c;
}
/* initialize_backward_search, move_structure_search.cpp:261-293 (ftab off) */
static interval_t iv_init(const movi_oracle_index* ix, unsigned char c) {
    uint64_t k = ix->alphamap[c] + 1;
    interval_t v = {ix->first_runs[k], ix->first_offsets[k], ix->last_runs[k], ix->last_offsets[k]};
    return v;
}
/* update_interval, move_structure_search.cpp:48-61 */
static void update_interval(const movi_oracle_index* ix, interval_t* v, int c) {
    while (v->rs <= v->re && get_char(ix, v->rs) != c) {
        v->rs += 1; v->os = 0;
        if (v->rs >= ix->r) break;
    }
    while (v->re >= v->rs && get_char(ix, v->re) != c) {
        v->re -= 1; v->oe = row_n(ix, v->re) - 1;
        if (v->re == 0) break;
    }
}
/* backward_search_step(R, pos_on_r, interval), move_structure_search.cpp:311-333 */
static int bs_step(const movi_oracle_index* ix, const char* R, int64_t pos, interval_t* v) {
    if (pos <= 0) return fail("[backward search step] never called on position 0");
    unsigned char c = (unsigned char)R[pos - 1];
    if (!check_alphabet(ix, c)) { iv_make_empty(v); return 0; }
    update_interval(ix, v, c);
    if (!iv_empty(v)) {
        if (lf_move(ix, &v->os, &v->rs) < 0) return -1;
        if (lf_move(ix, &v->oe, &v->re) < 0) return -1;
    }
    return 0;
}

/* query_backward_search (search.cpp:340-352) + backward_search (:169-201) + count (move_intervals.hpp:47-58) */
int movi_oracle_query_count(const movi_oracle_index* ix, const char* read, int64_t len,
                            uint64_t* count, int32_t* pos_on_r) {
    if (len <= 0) { *count = 0; *pos_on_r = 0; return 0; }   /* reference: UB on empty reads */
    int64_t pos = len - 1;
    if (!check_alphabet(ix, (unsigned char)read[pos])) { *pos_on_r = (int32_t)(pos + 1); *count = 0; return 0; }
    interval_t v = iv_init(ix, (unsigned char)read[pos]);
    interval_t prev = v;
    while (pos > 0 && !iv_empty(&v)) {
        prev = v;
        if (bs_step(ix, read, pos, &v)) return -1;
        if (!iv_empty(&v)) pos -= 1;
    }
    const interval_t* j = iv_empty(&v) ? &prev : &v;
    *count = iv_count(ix, j);
    *pos_on_r = (int32_t)pos;
    return 0;
}

/* query_zml, move_structure_query.cpp:690-786 */
int64_t movi_oracle_query_zml(const movi_oracle_index* ix, const char* R, int64_t len, uint16_t* out) {
    int64_t pos = len - 1;
    uint64_t m = 0;
    int64_t k = 0;
    /* :696 reads R[pos] before testing pos >= 0; the value read at -1 cannot change the outcome */
    while (pos >= 0 && !check_alphabet(ix, (unsigned char)R[pos])) { out[k++] = 0; pos -= 1; }
    if (pos < 0) return k;
    interval_t v = iv_init(ix, (unsigned char)R[pos]);
    while (pos > 0) {
        if (bs_step(ix, R, pos, &v)) return -1;
        if (!iv_empty(&v)) {
            out[k++] = clamp16(m); pos -= 1; m += 1;
        } else {
            out[k++] = clamp16(m); pos -= 1; m = 0;
            while (!check_alphabet(ix, (unsigned char)R[pos]) && pos > 0) { out[k++] = clamp16(m); pos -= 1; }
            if (check_alphabet(ix, (unsigned char)R[pos])) v = iv_init(ix, (unsigned char)R[pos]);
        }
    }
    if (iv_empty(&v)) m = 0;
    out[k++] = clamp16(m);
    return k;
}

int movi_oracle_query_ml_batch(const movi_oracle_index* ix, int kind, const char* bases,
                               const uint64_t* offsets, uint64_t n_reads, uint16_t* out, int threads) {
    int err = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#endif
    #pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : 1)
    for (uint64_t i = 0; i < n_reads; i++) {
        int64_t L = (int64_t)(offsets[i + 1] - offsets[i]);
        int rc;
        if (kind == 0) rc = movi_oracle_query_pml(ix, bases + offsets[i], L, out + offsets[i]);
        else if (kind == 2) rc = movi_oracle_query_rpml(ix, bases + offsets[i], L, out + offsets[i]);
        else rc = movi_oracle_query_zml(ix, bases + offsets[i], L, out + offsets[i]) == L ? 0 : -1;
        if (rc) {
            #pragma omp atomic write
            err = 1;
        }
    }
    return err ? -1 : 0;
}

int movi_oracle_query_count_batch(const movi_oracle_index* ix, const char* bases,
                                  const uint64_t* offsets, uint64_t n_reads,
                                  uint64_t* counts, int32_t* pos_on_r, int threads) {
    int err = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#endif
    #pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : 1)
    for (uint64_t i = 0; i < n_reads; i++) {
        int64_t L = (int64_t)(offsets[i + 1] - offsets[i]);
        if (movi_oracle_query_count(ix, bases + offsets[i], L, &counts[i], &pos_on_r[i])) {
            #pragma omp atomic write
            err = 1;
        }
    }
    return err ? -1 : 0;
}

/* Classifier::classify, classifier.cpp:99-143 (the decision and the report columns) */
int movi_oracle_classify(const uint16_t* ml, uint64_t len, uint64_t bin_width, uint64_t max_value_thr,
                         uint64_t* bins_above, uint64_t* bins_below, uint64_t* sum_max, uint64_t* n_bins) {
    uint64_t above = 0, below = 0, sum = 0, nb = 0, start = 0, end = 0;
    while (start < len) {
        end = (start + bin_width < len) ? start + bin_width : len;
        if (len - end < bin_width) end = len;
        uint16_t mx = ml[start];
        for (uint64_t i = start; i < end; i++) if (ml[i] > mx) mx = ml[i];
        if (mx >= max_value_thr) above++; else below++;
        sum += mx; nb++;
        start += (end - start);
    }
    *bins_above = above; *bins_below = below; *sum_max = sum; *n_bins = nb;
    return (double)above / ((double)(above + below) + 0.0) > 0.50;
}

/* Analysis hook (not part of any parity check): rows[k] = the row the k-th LF step of query_pml starts from,
 * i.e. the one random row fetch of that base.  Used by tools/ to study the access distribution of a workload. */
int movi_oracle_trace_pml_rows(const movi_oracle_index* ix, const char* read, int64_t len, uint64_t* rows) {
    int64_t pos = len - 1;
    uint64_t idx = ix->r - 1, offset = row_n(ix, idx) - 1, k = 0;
    while (pos > -1) {
        unsigned char rc = (unsigned char)read[pos];
        if (check_alphabet(ix, rc) && raw_char(ix, idx) != (int)rc) {
            int up = reposition_thresholds(ix, &idx, offset, rc);
            if (up < 0) return -1;
            offset = up ? row_n(ix, idx) - 1 : 0;
        }
        rows[k++] = idx;
        pos -= 1;
        if (lf_move(ix, &offset, &idx) < 0) return -1;
    }
    return 0;
}

/* ================================================================ Movi Color
 * Multi-class classification with PMLs (`movi query --pml --multi-classify --no-prefetch`) on the default flat color table, and
 * `--sa-entries`.  Files of the index directory, written by the reference's `color` / `build-SA` actions:
 *   doc_sets_flat.bin   u64 size; u16 flat_colors[size]; MoveTally(5 B) doc_set_flat_inds[r]        move_structure_io.cpp:536-546, 587-607
 *   ref.fa.doc_offsets  text, one cumulative length per document                                     prepare_ref.cpp:121-128
 *   ref.fa.doc_ids      text, optional: taxon id per document                                        move_structure_io.cpp:643-700
 *   ssa.movi            u64 rate; u64 k; u64 sampled_SA[k]; u64 r; u64 all_p[r]                       move_structure_io.cpp:702-735 */
#include <math.h>

struct movi_oracle_colors {
    uint64_t flat_size; uint16_t* flat; uint8_t* inds; uint64_t r;
    uint32_t num_docs, num_species;
    uint32_t* to_taxon_id;            /* species -> taxon id (ascending: std::map order) */
    double* log_lens;                 /* log(total length of the species' documents) */
};

static int cmp_u32(const void* a, const void* b) { uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b; return x < y ? -1 : x > y; }

void movi_oracle_colors_free(movi_oracle_colors* c) {
    if (!c) return;
    free(c->flat); free(c->inds); free(c->to_taxon_id); free(c->log_lens); free(c);
}

/* load_color_table -> deserialize_doc_sets_flat + load_document_info (movi.cpp:121-152) */
int movi_oracle_colors_load(const char* index_dir, uint64_t r, movi_oracle_colors** out) {
    char path[4096];
    movi_oracle_colors* c = calloc(1, sizeof *c);
    if (!c) return fail("out of memory");
    c->r = r;
    snprintf(path, sizeof path, "%s/doc_sets_flat.bin", index_dir);
    FILE* f = fopen(path, "rb");
    if (!f) { movi_oracle_colors_free(c); return fail("[deserialize doc sets flat] Failed to open document sets flat file"); }
    if (rd(f, &c->flat_size, 8)) { fclose(f); movi_oracle_colors_free(c); return fail("doc_sets_flat.bin: truncated"); }
    c->flat = malloc(c->flat_size * 2 + 2);
    c->inds = malloc(r * 5 + 8);
    if (!c->flat || !c->inds || rd(f, c->flat, c->flat_size * 2) || rd(f, c->inds, r * 5)) { fclose(f); movi_oracle_colors_free(c); return fail("doc_sets_flat.bin: truncated"); }
    fclose(f);

    snprintf(path, sizeof path, "%s/ref.fa.doc_offsets", index_dir);
    f = fopen(path, "r");
    if (!f) { movi_oracle_colors_free(c); return fail("[load document info] doc_offsets file not found"); }
    uint64_t cap = 16, nd = 0, v;
    uint64_t* offs = malloc(cap * 8);
    while (fscanf(f, "%lu", &v) == 1) { if (nd == cap) { cap *= 2; offs = realloc(offs, cap * 8); } offs[nd++] = v; }
    fclose(f);
    c->num_docs = (uint32_t)nd;
    uint32_t* ids = malloc((nd + 1) * 4);
    uint64_t ni = 0;
    snprintf(path, sizeof path, "%s/ref.fa.doc_ids", index_dir);
    f = fopen(path, "r");
    if (f) {                                              /* taxon id per document */
        uint32_t* all = malloc(16 * 4); uint64_t acap = 16; unsigned t;
        while (fscanf(f, "%u", &t) == 1) { if (ni == acap) { acap *= 2; all = realloc(all, acap * 4); } all[ni++] = t; }
        fclose(f);
        free(ids); ids = all;
    } else {
        for (uint64_t i = 0; i < nd; i++) ids[i] = (uint32_t)i + 1;
        ni = nd;
    }
    /* "compress taxon id to 0 .. num_species - 1": ascending order of the distinct ids (a std::map) */
    uint32_t* sorted = malloc((ni + 1) * 4);
    memcpy(sorted, ids, ni * 4);
    qsort(sorted, ni, 4, cmp_u32);
    uint32_t ns = 0;
    for (uint64_t i = 0; i < ni; i++) if (i == 0 || sorted[i] != sorted[i - 1]) sorted[ns++] = sorted[i];
    c->num_species = ns;
    c->to_taxon_id = sorted;
    c->log_lens = calloc(ns + 1, sizeof(double));
    for (uint64_t i = 0; i < nd && i < ni; i++) {
        uint32_t* p = bsearch(&ids[i], sorted, ns, 4, cmp_u32);
        uint64_t doc_len = offs[i] - (i == 0 ? 0 : offs[i - 1]);
        c->log_lens[p - sorted] += (double)doc_len;
    }
    for (uint32_t i = 0; i < ns; i++) c->log_lens[i] = log(c->log_lens[i]);
    free(ids); free(offs);
    *out = c;
    return 0;
}

uint32_t movi_oracle_colors_num_species(const movi_oracle_colors* c) { return c->num_species; }
uint32_t movi_oracle_colors_taxon(const movi_oracle_colors* c, uint32_t species) { return species < c->num_species ? c->to_taxon_id[species] : 0; }

/* query_pml with is_multi_classify() (move_structure_query.cpp:251-260, 351-437): after the LF step of every base whose
 * match_len >= min_len, the documents of the color of the row the walk now stands on are counted; best / second best follow
 * in the order the documents are met.  raw[] = {best_doc, second_best_doc (65535 = none), sum_matching_lengths (u32 arithmetic)};
 * score[] = the two counters the final comparison reads (:452-462).  taxa[] = the two numbers of the output line (:439-470). */
int movi_oracle_query_multi_classify(const movi_oracle_index* ix, const movi_oracle_colors* col, const char* read, int64_t len,
                                     int min_len, int pvalue, uint16_t* out_ml, uint32_t raw[3], double score[2], uint32_t taxa[2]) {
    const double log4 = log(4);                                   /* move_structure.hpp:315 */
    uint32_t* cnts = calloc(col->num_species + 1, 4);
    double* scores = calloc(col->num_species + 1, sizeof(double));
    int64_t pos = len - 1;
    uint64_t idx = ix->r - 1, offset = row_n(ix, idx) - 1, match_len = 0, k = 0;
    uint16_t best = 65535, second = 65535;
    uint32_t sum_ml = 0;
    int rcode = 0;
    while (pos > -1) {
        unsigned char rc = (unsigned char)read[pos];
        int row_ch = raw_char(ix, idx);
        if (!check_alphabet(ix, rc)) {
            match_len = 0;
        } else if (row_ch == (int)rc) {
            match_len += 1;
        } else {
            int up = reposition_thresholds(ix, &idx, offset, rc);
            if (up < 0) { rcode = -1; break; }
            match_len = 0;
            offset = up ? row_n(ix, idx) - 1 : 0;
        }
        sum_ml += (uint32_t)match_len;
        if (out_ml) out_ml[k] = clamp16(match_len);
        k++;
        pos -= 1;
        if (lf_move(ix, &offset, &idx) < 0) { rcode = -1; break; }
        if (match_len >= (uint64_t)min_len) {
            const uint8_t* e = col->inds + idx * 5;               /* MoveTally::get, move_row.hpp:28-38 */
            uint64_t color_id = (uint64_t)ld32(e) | ((uint64_t)e[4] << 32);
            if (color_id >= col->flat_size) continue;             /* :395 */
            uint32_t set_size = col->flat[color_id];
            for (uint32_t j = 0; j < set_size; j++) {
                int doc = col->flat[color_id + 1 + j];
                if (!pvalue) {
                    cnts[doc]++;
                    if (doc != best) {
                        if (best == 65535 || cnts[doc] > cnts[best]) { second = best; best = (uint16_t)doc; }
                        else if (second == 65535 || cnts[doc] > cnts[second]) second = (uint16_t)doc;
                    }
                } else {
                    double val = (double)match_len - (col->log_lens[doc] / log4);
                    if (val >= 0) {
                        scores[doc] += val < 1. ? val : 1.;
                        if (doc != best) {
                            if (best == 65535 || scores[doc] > scores[best]) { second = best; best = (uint16_t)doc; }
                            else if (second == 65535 || scores[doc] > scores[second]) second = (uint16_t)doc;
                        }
                    }
                }
            }
        }
    }
    raw[0] = best; raw[1] = second; raw[2] = sum_ml;
    score[0] = score[1] = 0;
    taxa[0] = taxa[1] = 0;
    float pml_mean = (float)sum_ml / (float)(uint64_t)len;        /* static_cast<float>(sum) / length() (:443) */
    if (len == 0) pml_mean = NAN;                                  /* 0 / 0: the comparison below is false, best_doc decides */
    if (!(pml_mean < 0.4) && best != 65535) {                      /* UNCLASSIFIED_THRESHOLD, utils.hpp:169 */
        taxa[0] = col->to_taxon_id[best];
        if (second != 65535) {
            float best_cnt, second_cnt;
            if (min_len > 0) { best_cnt = (float)cnts[best]; second_cnt = (float)cnts[second]; }
            else             { best_cnt = (float)scores[best]; second_cnt = (float)scores[second]; }
            float diff = best_cnt - second_cnt;
            if (diff < 0.05 * best_cnt) taxa[1] = col->to_taxon_id[second];
        }
    }
    if (best != 65535) score[0] = pvalue ? scores[best] : (double)cnts[best];
    if (second != 65535) score[1] = pvalue ? scores[second] : (double)cnts[second];
    free(cnts); free(scores);
    return rcode;
}

/* --sa-entries: sampled suffix-array entries (build-SA) */
struct movi_oracle_ssa { uint64_t rate, k, r; uint64_t* sampled; uint64_t* all_p; };

void movi_oracle_ssa_free(movi_oracle_ssa* s) { if (s) { free(s->sampled); free(s->all_p); free(s); } }

/* deserialize_sampled_SA, move_structure_io.cpp:716-735 */
int movi_oracle_ssa_load(const char* index_dir, movi_oracle_ssa** out) {
    char path[4096];
    snprintf(path, sizeof path, "%s/ssa.movi", index_dir);
    FILE* f = fopen(path, "rb");
    if (!f) return fail("[deserialize sampled SA] Failed to open sampled SA entries file");
    movi_oracle_ssa* s = calloc(1, sizeof *s);
    int bad = rd(f, &s->rate, 8) || rd(f, &s->k, 8);
    if (!bad) { s->sampled = malloc(s->k * 8 + 8); bad = !s->sampled || rd(f, s->sampled, s->k * 8) || rd(f, &s->r, 8); }
    if (!bad) { s->all_p = malloc(s->r * 8 + 8); bad = !s->all_p || rd(f, s->all_p, s->r * 8); }
    fclose(f);
    if (bad) { movi_oracle_ssa_free(s); return fail("ssa.movi: truncated"); }
    *out = s;
    return 0;
}

/* get_SA_entries, move_structure.cpp:35-48 */
static int get_sa_entry(const movi_oracle_index* ix, const movi_oracle_ssa* s, uint64_t idx, uint64_t offset, uint64_t* out) {
    uint64_t abs_offset = s->all_p[idx] + offset, distance = 0;
    while (abs_offset % s->rate != 0) {
        if (lf_move(ix, &offset, &idx) < 0) return -1;
        abs_offset = s->all_p[idx] + offset;
        distance += 1;
    }
    *out = s->sampled[abs_offset / s->rate] + distance;
    return 0;
}

/* query_pml with is_get_sa_entries() (move_structure_query.cpp:353-357): one entry per base, same orientation as the lengths */
int movi_oracle_query_pml_sa(const movi_oracle_index* ix, const movi_oracle_ssa* s, const char* read, int64_t len,
                             uint16_t* out_ml, uint64_t* out_sa) {
    int64_t pos = len - 1;
    uint64_t idx = ix->r - 1, offset = row_n(ix, idx) - 1, match_len = 0, k = 0;
    while (pos > -1) {
        unsigned char rc = (unsigned char)read[pos];
        int row_ch = raw_char(ix, idx);
        if (!check_alphabet(ix, rc)) {
            match_len = 0;
        } else if (row_ch == (int)rc) {
            match_len += 1;
        } else {
            int up = reposition_thresholds(ix, &idx, offset, rc);
            if (up < 0) return -1;
            match_len = 0;
            offset = up ? row_n(ix, idx) - 1 : 0;
        }
        if (out_ml) out_ml[k] = clamp16(match_len);
        if (get_sa_entry(ix, s, idx, offset, &out_sa[k])) return -1;
        k++;
        pos -= 1;
        if (lf_move(ix, &offset, &idx) < 0) return -1;
    }
    return 0;
}

/* find_sampled_SA_entries (move_structure_build.cpp:1174-1212), the `build-SA` action: all_p = BWT offset of every run head; one LF
 * walk over the whole BWT from row 0, offset 0, SA value n - 1 downwards, keeping the value at every rate-th BWT offset. */
int movi_oracle_ssa_build(const movi_oracle_index* ix, uint64_t rate, movi_oracle_ssa** out) {
    if (rate == 0) return fail("sample rate 0");
    movi_oracle_ssa* s = calloc(1, sizeof *s);
    s->rate = rate; s->r = ix->r;
    s->all_p = malloc(ix->r * 8 + 8);
    uint64_t tot = 0;
    for (uint64_t i = 0; i < ix->r; i++) { s->all_p[i] = tot; tot += row_n(ix, i); }
    s->k = tot / rate + 1;
    s->sampled = calloc(s->k + 1, 8);
    uint64_t offset = 0, index = 0, sa = tot;
    for (uint64_t i = 0; i < tot; i++) {
        sa--;
        const uint64_t row_ind = s->all_p[index] + offset;
        if (row_ind % rate == 0) s->sampled[row_ind / rate] = sa;
        if (lf_move(ix, &offset, &index) < 0) { movi_oracle_ssa_free(s); return -1; }
    }
    *out = s;
    return 0;
}
uint64_t movi_oracle_ssa_rate(const movi_oracle_ssa* s) { return s->rate; }
uint64_t movi_oracle_ssa_samples(const movi_oracle_ssa* s) { return s->k; }
uint64_t movi_oracle_ssa_rows(const movi_oracle_ssa* s) { return s->r; }
const uint64_t* movi_oracle_ssa_sampled(const movi_oracle_ssa* s) { return s->sampled; }
const uint64_t* movi_oracle_ssa_all_p(const movi_oracle_ssa* s) { return s->all_p; }
