// dev_index.cuh -- device-resident form of a Movi index and the row decode.
//
// Device row format "R8": one 64-bit word per move row, 4 rows per 32-byte HBM sector.  Same fields as the reference's
// regular-thresholds MoveRow (include/move_row.hpp:131-142, include/move_row_configs.hpp:34-51), placed so that the query
// kernels -- bound by the ALU pipe, DESIGN.md section 7 -- decode them with the fewest instructions:
//     x (low 32)   : id bits 0..31
//     y bits  0..10: n (run length, <= 2047)            y bits 11..21: offset
//     y bits 22..24: id bits 32..34                      y bits 29..31: c (alphabet index; one shift, no mask)
//     y bits 25..28: threshold mask, bit 25 + a = "the threshold against read character a is n (else 0)" for the four
//                    alphabet indices a (with separators: a = code - 1).  The reference stores three bits and finds the one
//                    to use through alphamap_3[c][a] (src/utils.cpp:5-8, move_structure_query.cpp:545-556); that lookup is
//                    resolved once per row at upload.  (Rows assembled on the fly from the packed layouts carry the
//                    reference's three bits in 25..27 instead: make_r8_raw3 / row_thr_raw3.)
// Every index type is re-laid out into R8 on the device at upload time (convert.cuh): regular-thresholds by permuting the
// bits of its 8-byte rows in place, blocked-thresholds (6-byte rows + id_blocks) and sampled-thresholds (3-byte rows + id
// checkpoints) by also resolving the full destination id once per row instead of once per LF step.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace movi {

constexpr int LAYOUT_R8 = 0;      // 8-byte rows, full destination id in the row
constexpr int LAYOUT_B6 = 1;      // the reference's blocked-thresholds rows as they are on disk: 6 bytes per row + id_blocks
constexpr int LAYOUT_T4 = 2;      // sampled-thresholds: 4 bytes per row (the 3-byte row + its walk-back distance) + the id checkpoints

struct DevIndex {
    const uint2* rows;            // LAYOUT_R8: R8 rows
    const uint32_t* rows6;        // LAYOUT_B6: packed 6-byte rows viewed as 32-bit words (table padded by 8 bytes)
    const uint32_t* id_blocks;    // LAYOUT_B6: sigma x id_blocks_size checkpoints of the destination id (src/move_structure.cpp:95-102)
    uint32_t id_blocks_size;
    uint32_t block_shift;         // block_size == 1 << block_shift
    uint32_t layout;
    uint32_t nt;                  // 1: non-threshold index type (12-bit lengths, no threshold mask; PML repositions with the --rpml rule)
    const uint32_t* rows4;        // LAYOUT_T4: one 32-bit word per row (Rows4 below)
    const uint8_t* tally;         // LAYOUT_T4: sigma x tally_len MoveTally entries (5 bytes each), 4-byte aligned, padded by 8 bytes
    unsigned long long tally_len;
    uint32_t tally_cp;            // id checkpoint interval (tally_checkpoints)
    unsigned long long r;
    unsigned long long end_idx;   // row of '$'
    uint32_t n_last;              // n(r-1)
    uint32_t sep;                 // 1: alphabet is "%ACGT" (codes 1..4 legal), 0: codes 0..sigma-1
    uint32_t sigma;
    uint32_t sep_cnt;
    uint32_t end_thr[4];          // end_bwt_idx_thresholds, saturated to 32 bits
    const unsigned long long* sep_keys;   // sorted rows holding '%' (separators_thresholds_map keys)
    const uint16_t* sep_vals;             // 4 x u16 per key (ThresholdsRow), same order as sep_keys, + 1 entry = slot 0
    // backward-search seeds, index k = code + 1 (src/move_structure_search.cpp:285-292)
    unsigned long long first_runs[8], last_runs[8];
    uint32_t first_offsets[8], last_offsets[8];
    const unsigned long long* prefix;     // sampled run-length prefix sums: prefix[k] = sum_{j < k*PREFIX_STRIDE} n(j)
    uint8_t code_of[256];         // byte -> alphabet index, 0xFF = not a legal query character
};

constexpr int PREFIX_SHIFT = 4;   // one u64 prefix per 16 rows (0.5 B/row)
constexpr uint32_t CODE_ILLEGAL = 0xFFu;

constexpr int R8_MAX_ROW_BITS = 35;     // 32 + 3 id bits: 3.4e10 rows = 275 GB of rows, more than a B200 holds

// NT = the table came from a non-threshold index type ("regular", MODE 3: include/move_row_configs.hpp:20-31), whose run
// lengths and offsets have 12 bits.  Its R8 word has no threshold mask: n[0:12] | offset[12:24] | id[32:35] << 24 | c << 29.
template <bool NT = false> __host__ __device__ __forceinline__ uint32_t row_n(uint32_t y) { return NT ? (y & 0xFFFu) : (y & 0x7FFu); }
__host__ __device__ __forceinline__ uint32_t row_c(uint32_t y) { return y >> 29; }
template <bool NT = false> __host__ __device__ __forceinline__ uint32_t row_off(uint32_t y) { return NT ? ((y >> 12) & 0xFFFu) : ((y >> 11) & 0x7FFu); }
template <typename IdxT, bool NT = false>
__host__ __device__ __forceinline__ IdxT row_id(uint2 v) {
    if constexpr (sizeof(IdxT) == 8) return (IdxT)v.x | ((IdxT)((v.y >> (NT ? 24 : 22)) & 7u) << 32);
    else return (IdxT)v.x;
}
// the reference's "regular" row (n 0..11, c 13..15 | offset 16..27, id high 28..31) <-> R8 (NT)
__host__ __device__ __forceinline__ uint2 r8nt_from_file(uint2 f) {
    const uint32_t y = f.y;
    uint2 v; v.x = f.x;
    v.y = (y & 0xFFFu) | (((y >> 16) & 0xFFFu) << 12) | (((y >> 28) & 7u) << 24) | (((y >> 13) & 7u) << 29);
    return v;
}
__host__ __device__ __forceinline__ uint2 r8nt_to_file(uint2 v) {
    const uint32_t y = v.y;
    uint2 f; f.x = v.x;
    f.y = (y & 0xFFFu) | ((y >> 29) << 13) | (((y >> 12) & 0xFFFu) << 16) | (((y >> 24) & 7u) << 28);
    return f;
}
// alphamap_3 (src/utils.cpp:5-8): slot of read character b among the three thresholds of a row with character a
__host__ __device__ __forceinline__ uint32_t thr_slot(uint32_t a, uint32_t b) { return b - (b > a ? 1u : 0u); }
// threshold of a resident row against read character index ai (get_thresholds, move_structure.cpp:306-308): n or 0
__host__ __device__ __forceinline__ uint32_t row_thr(uint32_t y, uint32_t ai, uint32_t n) { return ((y >> (ai & 3u)) & (1u << 25)) ? n : 0u; }
// the same for a row assembled on the fly from a packed layout (three reference bits in 25..27); a = c - sep
__device__ __forceinline__ uint32_t row_thr_raw3(uint32_t y, uint32_t a, uint32_t ai, uint32_t n) {
    return ((y >> (25u + (thr_slot(a, ai) & 3u))) & 1u) ? n : 0u;
}
__host__ __device__ __forceinline__ uint32_t r8_word(unsigned long long id, uint32_t n, uint32_t off, uint32_t c, uint32_t bits25) {
    return (n & 0x7FFu) | ((off & 0x7FFu) << 11) | ((uint32_t)((id >> 32) & 7ull) << 22) | ((bits25 & 15u) << 25) | ((c & 7u) << 29);
}
// thr_bits = the reference's thresholds #0..#2 in bits 0..2; sep = 1 for the "%ACGT" alphabet
__host__ __device__ __forceinline__ uint2 make_r8(unsigned long long id, uint32_t n, uint32_t off, uint32_t c, uint32_t thr_bits, uint32_t sep) {
    const uint32_t a = c - sep;                     // ('%' rows: a wraps; their thresholds live in the side table)
    uint32_t mask = 0;
    for (uint32_t ai = 0; ai < 4; ai++)
        if (ai != a) mask |= ((thr_bits >> (thr_slot(a, ai) & 3u)) & 1u) << ai;
    uint2 v;
    v.x = (uint32_t)(id & 0xFFFFFFFFull);
    v.y = r8_word(id, n, off, c, mask);
    return v;
}
__host__ __device__ __forceinline__ uint2 make_r8_raw3(unsigned long long id, uint32_t n, uint32_t off, uint32_t c, uint32_t thr_bits) {
    uint2 v;
    v.x = (uint32_t)(id & 0xFFFFFFFFull);
    v.y = r8_word(id, n, off, c, thr_bits & 7u);
    return v;
}
// The reference's regular-thresholds row (MODE 6 bit fields: n 0..10, thr#1 11, thr#2 12, c 13..15, offset 16..26, thr#0 27,
// id high 28..31) <-> R8
__host__ __device__ __forceinline__ uint2 r8_from_file(uint2 f, uint32_t sep) {
    const uint32_t y = f.y;
    const uint32_t thr = ((y >> 27) & 1u) | ((y >> 10) & 6u);
    return make_r8((unsigned long long)f.x | ((unsigned long long)(y >> 28) << 32), y & 0x7FFu, (y >> 16) & 0x7FFu, (y >> 13) & 7u, thr, sep);
}
__host__ __device__ __forceinline__ uint2 r8_to_file(uint2 v, uint32_t sep) {
    const uint32_t y = v.y, c = row_c(y), a = c - sep, mask = (y >> 25) & 15u;
    uint32_t thr = 0;
    for (uint32_t ai = 0; ai < 4; ai++)
        if (ai != a) thr |= ((mask >> ai) & 1u) << (thr_slot(a, ai) & 3u);
    thr &= 7u;
    uint2 f;
    f.x = v.x;
    f.y = row_n(y) | (((thr >> 1) & 1u) << 11) | (((thr >> 2) & 1u) << 12) | (c << 13) | (row_off(y) << 16) | ((thr & 1u) << 27) |
          (((y >> 22) & 7u) << 28);
    return f;
}

__device__ __forceinline__ uint2 ldg64(const uint2* p) { return __ldg(p); }

// Run length of row k whatever the device layout (count query: sums over the rows an interval spans)
__device__ __forceinline__ uint32_t row_len_at(const DevIndex& ix, unsigned long long k) {
    if (ix.layout == LAYOUT_B6) return (uint32_t)reinterpret_cast<const uint16_t*>(ix.rows6)[3 * k + 1] & 0x3FFu;
    return ix.nt ? row_n<true>(ix.rows[k].y) : row_n<false>(ix.rows[k].y);
}

// LAYOUT_B6: the packed blocked-thresholds row (include/move_row.hpp:128-130,139-142; bit fields
// include/move_row_configs.hpp:76-104) fetched with two aligned 32-bit loads and re-expressed as an R8 word pair, with
// MoveStructure::get_id for BLOCKED_MODES (src/move_structure.cpp:95-102) applied: id22 + id_blocks[c][idx / block_size]
// + first_runs[c + 1] (the '$' row keeps its raw id).  idb points at the id_blocks copy to use (shared memory when it fits).
template <typename IdxT>
__device__ __forceinline__ uint2 fetch_row_b6(const uint32_t* __restrict__ w, IdxT idx, IdxT end_idx, const uint32_t* idb,
                                              uint32_t nb, uint32_t shift, const unsigned long long* first_runs) {
    const unsigned long long a = 6ull * (unsigned long long)idx;
    const uint32_t* p = w + (a >> 2);
    const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1);
    const uint32_t sh = (uint32_t)(a & 2u) * 8u;
    const uint32_t lo = __funnelshift_r(w0, w1, sh);          // f_id | f_n << 16
    const uint32_t f_off = (w1 >> sh) & 0xFFFFu;
    const uint32_t f_id = lo & 0xFFFFu, f_n = lo >> 16;
    const uint32_t n = f_n & 0x3FFu, off = f_off & 0x3FFu, c = (f_off >> 10) & 7u, thr = f_off >> 13;
    unsigned long long id = (unsigned long long)f_id | ((unsigned long long)(f_n >> 10) << 16);
    if (idx != end_idx) id += (unsigned long long)idb[c * nb + (uint32_t)((unsigned long long)idx >> shift)] + first_runs[c + 1];
    return make_r8_raw3(id, n, off, c, thr);
}

// threshold value of a '%' row (separators_thresholds[separators_thresholds_map[idx]].values[ai],
// src/move_structure_query.cpp:540-541); rows missing from the map read slot 0 like unordered_map::operator[]
__host__ __device__ __forceinline__ uint32_t sep_threshold(const DevIndex& ix, unsigned long long idx, uint32_t ai) {
    uint32_t lo = 0, hi = ix.sep_cnt;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (ix.sep_keys[mid] < idx) lo = mid + 1; else hi = mid;
    }
    if (lo < ix.sep_cnt && ix.sep_keys[lo] == idx) return ix.sep_vals[lo * 4 + ai];
    return ix.sep_vals[ix.sep_cnt * 4 + ai];   // entry sep_cnt mirrors slot 0 (set at upload)
}

// ---- sampled-thresholds (MODE 7): 3-byte rows {u8 n; u8 offset; u8 c} (include/move_row.hpp:122-126; bit fields
//      include/move_row_configs.hpp:120-136), MoveTally = {u32 right; u8 left} (include/move_row.hpp:13-39)
struct Rows3 {              // the packed table read through aligned 32-bit words (two loads per row)
    const uint32_t* w;
    unsigned long long end_idx;
    uint32_t nothr = 0;     // 1: MODE 5 "sampled" (no thresholds): two length bits and a 4-bit character in the third byte
                            //    (include/move_row_configs.hpp:107-118) instead of one and three + thresholds (:120-136)
    __device__ __forceinline__ uint32_t raw(unsigned long long i) const {
        const unsigned long long a = 3ull * i;
        const uint32_t* p = w + (a >> 2);
        return __funnelshift_r(__ldg(p), __ldg(p + 1), (uint32_t)(a & 3u) * 8u) & 0xFFFFFFu;
    }
    static __device__ __forceinline__ uint32_t n_of(uint32_t x) { return (x & 0xFFu) | (((x >> 17) & 1u) << 8); }
    static __device__ __forceinline__ uint32_t off_of(uint32_t x) { return ((x >> 8) & 0xFFu) | (((x >> 16) & 1u) << 8); }
    static __device__ __forceinline__ uint32_t c_of(uint32_t x) { return (x >> 18) & 7u; }
    static __device__ __forceinline__ uint32_t thr_of(uint32_t x) { return (x >> 21) & 7u; }
    __device__ __forceinline__ uint32_t dn(uint32_t x) const { return nothr ? ((x & 0xFFu) | (((x >> 18) & 3u) << 8)) : n_of(x); }
    __device__ __forceinline__ uint32_t doff(uint32_t x) const { return nothr ? (((x >> 8) & 0xFFu) | (((x >> 16) & 3u) << 8)) : off_of(x); }
    __device__ __forceinline__ uint32_t dc(uint32_t x) const { return nothr ? ((x >> 20) & 15u) : c_of(x); }
    __device__ __forceinline__ uint32_t n(unsigned long long i) const { return dn(raw(i)); }
    __device__ __forceinline__ uint32_t off(unsigned long long i) const { return doff(raw(i)); }
    // get_char (src/move_structure.cpp:288-293): '$' (here 16) for the end row
    __device__ __forceinline__ uint32_t ch(unsigned long long i) const { return i == end_idx ? 16u : dc(raw(i)); }
};

__device__ __forceinline__ unsigned long long tally_get(const uint8_t* tally, unsigned long long tally_len, uint32_t c, unsigned long long k) {
    const uint8_t* q = tally + (c * tally_len + k) * 5;
    return (unsigned long long)q[0] | ((unsigned long long)q[1] << 8) | ((unsigned long long)q[2] << 16) |
           ((unsigned long long)q[3] << 24) | ((unsigned long long)q[4] << 32);
}

// MoveStructure::get_id for TALLY_MODES (src/move_structure.cpp:103-219; the forward direction is always taken, :147-148):
// scan from idx to the next id checkpoint summing the lengths of the rows with idx's character, then walk back from the
// checkpoint's id.  R is any accessor with n(i), off(i), ch(i).  *bad is set where the reference throws (:196-199).
template <typename R>
__device__ __forceinline__ unsigned long long sampled_get_id(const R& rows, const uint8_t* tally, unsigned long long tally_len, uint32_t cp,
                                                             unsigned long long r, unsigned long long end_idx, unsigned long long idx,
                                                             uint32_t ci, bool* bad) {
    if (idx == end_idx) return 0;                                                        // :106-108
    if (idx == r - 1) return tally_get(tally, tally_len, ci, tally_len - 1);             // :114-117
    if (idx % cp == 0) return tally_get(tally, tally_len, ci, idx / cp);                 // :121-124
    const unsigned long long tally_b = idx / cp + 1;
    unsigned long long next_cp = tally_b * cp;
    if (next_cp >= r) next_cp = r - 1;                                                   // :135-139
    unsigned long long id = tally_get(tally, tally_len, ci, tally_b);
    unsigned long long rows_until = 0, last_id = r;
    const uint32_t my_ch = rows.ch(idx);
    for (unsigned long long i = idx; i < next_cp; i++)                                   // :165-171
        if (rows.ch(i) == my_ch) { rows_until += rows.n(i); last_id = i; }
    const bool same_as_cp = (my_ch == rows.ch(next_cp));
    if (last_id == idx && !same_as_cp) return id;                                        // :175-177
    uint32_t off = rows.off(next_cp);
    if (!same_as_cp) { rows_until -= rows.n(last_id); off = rows.off(last_id); }         // :190-193
    if (id >= r || off >= rows.n(id)) { *bad = true; return id < r ? id : 0; }           // :196-199 (reference throws)
    if (off >= rows_until) return id;                                                    // :200-201
    rows_until -= (off + 1);
    id -= 1;
    while (rows_until != 0) {                                                            // :207-215
        const uint32_t nn = rows.n(id);
        if (rows_until >= nn) { rows_until -= nn; id -= 1; } else rows_until = 0;
    }
    return id;
}


// ---- LAYOUT_T4: the memory-constrained device form of a sampled-thresholds index.  One aligned 32-bit word per row:
//     n[0:9) | offset[9:18) | c[18:21) | thresholds #0..#2 [21:24) | d[24:32)
// The first 24 bits are the reference's 3-byte row (include/move_row_configs.hpp:120-136) regrouped; d is computed once per row
// at upload: MoveStructure::get_id for TALLY_MODES (src/move_structure.cpp:103-219) ends in "the id sampled at the next
// checkpoint, minus the rows walked back from it", so
//     get_id(idx) = tally_ids[c][k(idx)] - d(idx),   k = idx / cp  at a checkpoint row, the last entry for the last row, idx / cp + 1 otherwise
// and the per-step scan of up to `cp` rows + walk-back (20 + ~5 dependent row reads) becomes one 5-byte table read.  d = 255 marks
// the rare row whose distance does not fit: there the full procedure runs on the resident rows.  4 + 5 * sigma / cp bytes per row
// (5 with four characters and checkpoints every 20 rows) against 8 for the expanded layout.
struct Rows4 {
    const uint32_t* w;
    unsigned long long end_idx;
    static __host__ __device__ __forceinline__ uint32_t pack(uint32_t n, uint32_t off, uint32_t c, uint32_t thr, uint32_t d) {
        return (n & 0x1FFu) | ((off & 0x1FFu) << 9) | ((c & 7u) << 18) | ((thr & 7u) << 21) | (d << 24);
    }
    static __device__ __forceinline__ uint32_t n_of(uint32_t x) { return x & 0x1FFu; }
    static __device__ __forceinline__ uint32_t off_of(uint32_t x) { return (x >> 9) & 0x1FFu; }
    static __device__ __forceinline__ uint32_t c_of(uint32_t x) { return (x >> 18) & 7u; }
    static __device__ __forceinline__ uint32_t thr_of(uint32_t x) { return (x >> 21) & 7u; }
    static __device__ __forceinline__ uint32_t d_of(uint32_t x) { return x >> 24; }
    __device__ __forceinline__ uint32_t n(unsigned long long i) const { return n_of(__ldg(w + i)); }
    __device__ __forceinline__ uint32_t off(unsigned long long i) const { return off_of(__ldg(w + i)); }
    __device__ __forceinline__ uint32_t ch(unsigned long long i) const { return i == end_idx ? 16u : c_of(__ldg(w + i)); }
};
constexpr uint32_t T4_FAR = 255u;

// one MoveTally entry through two aligned 32-bit loads (the table is 4-byte aligned and padded)
__device__ __forceinline__ unsigned long long tally_get_w(const uint8_t* tally, unsigned long long tally_len, uint32_t c, unsigned long long k) {
    const unsigned long long a = (c * tally_len + k) * 5ull;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(tally) + (a >> 2);
    const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1);
    const uint32_t lo = __funnelshift_r(w0, w1, (uint32_t)(a & 3u) * 8u);
    const uint32_t hi = (w1 >> ((uint32_t)((a + 4u) & 3u) * 8u)) & 0xFFu;
    return (unsigned long long)lo | ((unsigned long long)hi << 32);
}
// which checkpoint entry get_id starts from (src/move_structure.cpp:114-139)
__device__ __forceinline__ unsigned long long t4_tally_slot(unsigned long long idx, unsigned long long r, unsigned long long tally_len, uint32_t cp) {
    if (idx == r - 1) return tally_len - 1;
    const unsigned long long q = idx / cp;
    return idx - q * cp == 0 ? q : q + 1;
}
template <typename IdxT>
__device__ __forceinline__ IdxT t4_get_id(const DevIndex& ix, IdxT idx, IdxT end_idx, uint32_t x, bool* bad) {
    if (idx == end_idx) return 0;                                                                    // :106-108
    const uint32_t c = Rows4::c_of(x), d = Rows4::d_of(x);
    if (d == T4_FAR)
        return (IdxT)sampled_get_id(Rows4{ix.rows4, (unsigned long long)end_idx}, ix.tally, ix.tally_len, ix.tally_cp, ix.r,
                                    (unsigned long long)end_idx, (unsigned long long)idx, c, bad);
    IdxT k;
    if (sizeof(IdxT) == 4) {             // 32-bit division: tables below 2^32 rows
        const uint32_t i32 = (uint32_t)idx, q = i32 / ix.tally_cp;
        k = (i32 == (uint32_t)(ix.r - 1)) ? (IdxT)(ix.tally_len - 1) : (IdxT)(i32 - q * ix.tally_cp == 0 ? q : q + 1);
    } else k = (IdxT)t4_tally_slot((unsigned long long)idx, ix.r, ix.tally_len, ix.tally_cp);
    return (IdxT)(tally_get_w(ix.tally, ix.tally_len, c, (unsigned long long)k) - d);
}

}  // namespace movi
