// pml_kernel.cuh -- pseudo-matching lengths: many reads per warp, one read per lane.
//
// Restates, for a whole batch at once, the per-base loop of MoveStructure::query_pml
// (src/move_structure_query.cpp:234-474): compare the row character with the read base, on a
// mismatch reposition with the threshold bit (reposition_thresholds :513-601, reposition_up/down
// :188-232), emit the matching length (MoveQuery::add_ml, include/move_query.hpp:26-38), then
// LF_move (src/move_structure.cpp:59-87) with fast_forward (:524-545).
//
// Design (B200): the walk is a chain of dependent random 8-byte row fetches, so throughput comes from
// the number of independent reads in flight, not from any one read.  Every lane owns one read and runs a
// small state machine whose every iteration issues EXACTLY ONE row fetch for all lanes together:
//     ST_FF    row just reached by an LF jump or a fast-forward step: if offset >= n, step to the next
//              row (costs an iteration); otherwise process the next base against this row in place
//     ST_UP/ST_DOWN  scanning for the nearest row with the read's character after a mismatch
//     ST_NEW   fetch the read's [start, end) offsets (uses the same load slot as the row fetch)
// so fast-forward and reposition scans of one lane never stall the other 31 lanes (the CPU reference
// hides the same latency with 16 "strands" per thread and software prefetch,
// src/read_processor.cpp:641-730).  Finished lanes refill from a warp-local queue that is topped up
// from one global atomic counter in chunks, so reads of unequal length stay load balanced.
#pragma once
#include "dev_index.cuh"

namespace movi {

struct QueryArgs {
    const uint8_t* bases;                 // 8-byte aligned; read i = bases[offsets[i]-origin .. offsets[i+1]-origin)
    const unsigned long long* offsets;    // n_reads + 1
    unsigned long long origin;            // subtracted from every offset
    uint32_t n_reads;
    uint16_t* out_ml;                     // PML / ZML values at the same (offset - origin) positions
    unsigned long long* out_count;        // count query
    int32_t* out_pos;                     // count query
    unsigned long long* counter;          // work-queue head (zeroed before launch)
    unsigned int* status;                 // sticky error bits
    uint32_t ld_mode;                     // row fetch cache recipe (pml_kernel2.cuh: ld_row)
    uint32_t flags;                       // MOVI_GPU_FLAG_* of the call
    float keep_frac;                      // fraction of the row table whose fetches carry L2 evict_last (the rest: evict_first)
};

constexpr uint32_t ST_IDLE = 0, ST_NEW = 1, ST_FF = 2, ST_UP = 3, ST_DOWN = 4;
constexpr uint32_t QUEUE_CHUNK = 32;      // reads taken from the global counter per warp refill (small: short tail)
constexpr unsigned FULL = 0xffffffffu;

// Register-buffered output (OUTV == 1): the last `pend` (< 8) pushed values sit in the top `pend`
// 16-bit positions of (b0..b3); write them to out[op - pend + 1 .. op] with 2-byte stores.  Only used
// for the ragged head / tail group of a read; full groups go out as one 16-byte store.
__device__ __forceinline__ void flush_partial(uint16_t* out, uint32_t op, uint32_t pend,
                                              uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
    for (uint32_t t = 0; t < pend; t++) {
        const uint32_t pos = 8u - pend + t;
        const uint32_t ws = pos >> 1;
        const uint32_t w = ws == 0 ? b0 : ws == 1 ? b1 : ws == 2 ? b2 : b3;
        out[op - pend + 1 + t] = (uint16_t)((pos & 1u) ? (w >> 16) : (w & 0xFFFFu));
    }
}

// OUTV: 0 = one 2-byte store per base, 1 = 8 values buffered in registers, 16-byte stores
template <typename IdxT, int OUTV>
__global__ void __launch_bounds__(256)
movi_pml_kernel(const DevIndex ix, const QueryArgs q) {
    __shared__ uint8_t s_code[256];
    __shared__ uint32_t s_end_thr[4];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_code[i] = ix.code_of[i];
    if (threadIdx.x < 4) s_end_thr[threadIdx.x] = ix.end_thr[threadIdx.x];
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const IdxT r1 = (IdxT)(ix.r - 1);
    const IdxT end_idx = (IdxT)ix.end_idx;
    const uint2* __restrict__ rows = ix.rows;
    const uint2* __restrict__ bases8 = reinterpret_cast<const uint2*>(q.bases);
    const uint32_t sep = ix.sep;
    const uint32_t n_reads = q.n_reads;

    uint32_t q_next = 0, q_end = 0;
    bool exhausted = false;                 // warp uniform
    uint32_t st = ST_IDLE;
    IdxT idx = 0;                           // row index; read index while ST_NEW
    uint32_t off = 0, ml = 0;
    uint32_t rp = 0, op = 0, rem = 0;       // next base to read, next output slot, bases left
    uint32_t clo = 0, chi = 0;              // the 8 read bytes that contain bases[rp]
    bool need_chunk = false;
    uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0, pend = 0;   // OUTV==1: output shift register (8 x u16) + fill count

    for (;;) {
        // ---- hand reads to idle lanes (warp-aggregated; one global atomic per QUEUE_CHUNK reads)
        const bool want = (st == ST_IDLE) && !exhausted;
        const unsigned m = __ballot_sync(FULL, want);
        if (m) {
            if (q_next == q_end) {
                unsigned long long b = 0;
                if (lane == 0) b = atomicAdd(q.counter, (unsigned long long)QUEUE_CHUNK);
                b = __shfl_sync(FULL, b, 0);
                q_next = b < n_reads ? (uint32_t)b : n_reads;
                q_end = b + QUEUE_CHUNK < n_reads ? (uint32_t)(b + QUEUE_CHUNK) : n_reads;
                if (q_next == q_end) exhausted = true;
            }
            const uint32_t avail = q_end - q_next;
            const uint32_t rank = __popc(m & lt_mask);
            if (want && rank < avail) { idx = (IdxT)(q_next + rank); st = ST_NEW; }
            const uint32_t cnt = __popc(m);
            q_next += cnt < avail ? cnt : avail;
        }
        if (exhausted && __all_sync(FULL, st == ST_IDLE)) break;

        // ---- the two load slots of this iteration
        uint2 v = make_uint2(0u, 0u), w = make_uint2(0u, 0u);
        if (st != ST_IDLE) {
            const uint2* a = (st == ST_NEW) ? reinterpret_cast<const uint2*>(q.offsets + idx) : rows + idx;
            v = ldg64(a);
            if (st == ST_NEW || need_chunk) {
                const uint2* b = (st == ST_NEW) ? reinterpret_cast<const uint2*>(q.offsets + idx + 1) : bases8 + (rp >> 3);
                w = ldg64(b);
            }
        }

        if (st == ST_NEW) {
            const unsigned long long s = ((unsigned long long)v.y << 32 | v.x) - q.origin;
            const unsigned long long e = ((unsigned long long)w.y << 32 | w.x) - q.origin;
            const uint32_t len = (uint32_t)(e - s);
            if (len == 0) {
                st = ST_IDLE;
            } else {
                // query_pml :237-240: start at the last row, last offset
                rp = (uint32_t)e - 1; op = (uint32_t)s; rem = len;
                idx = r1; off = ix.n_last - 1; ml = 0;
                st = ST_FF; need_chunk = true; pend = 0;
            }
        } else if (st != ST_IDLE) {
            if (need_chunk) { clo = w.x; chi = w.y; need_chunk = false; }
            const uint32_t y = v.y;
            const uint32_t n = row_n(y);
            bool at_base = true;
            if (st == ST_FF && idx < r1 && off >= n) {       // fast_forward, one row per iteration
                off -= n; idx += 1; at_base = false;
            }
            if (at_base) {
                const uint32_t c = row_c(y);
                const uint32_t byte = __byte_perm(clo, chi, rp & 7u) & 0xFFu;
                const uint32_t code = s_code[byte];
                bool emit = false, bad = false;
                if (st == ST_FF) {
                    if (code == CODE_ILLEGAL) { ml = 0; emit = true; }          // :278-284
                    else if (c == code) { ml += 1; emit = true; }               // :285-296 (Case 1)
                    else {                                                      // :297-350 (Case 2)
                        uint32_t thr;
                        const uint32_t ai = code - sep;
                        if (idx == end_idx) thr = s_end_thr[ai & 3u];
                        else if (sep && c == 0) thr = sep_threshold(ix, (unsigned long long)idx, ai);
                        else thr = row_thr(y, ai, n);                                        // get_thresholds
                        if (off >= thr) { if (idx == r1) bad = true; else { idx += 1; st = ST_DOWN; } }
                        else            { if (idx == 0)  bad = true; else { idx -= 1; st = ST_UP; } }
                    }
                } else if (c == code) {                                         // scan hit: :313-327
                    off = (st == ST_UP) ? n - 1 : 0u; ml = 0; emit = true;
                } else if (st == ST_UP) {
                    if (idx == 0) bad = true; else idx -= 1;
                } else {
                    if (idx == r1) bad = true; else idx += 1;
                }
                if (bad) {            // reference: reposition returned r -> "This should not happen"
                    atomicOr(q.status, 1u);
                    st = ST_IDLE;
                }
                if (emit) {
                    const uint32_t val = ml < 65535u ? ml : 65535u;             // add_ml clamp
                    if (OUTV == 0) {
                        q.out_ml[op] = (uint16_t)val;
                    } else {
                        // shift register: the newest value enters at the top; after 8 pushes the oldest is at bit 0
                        b0 = __funnelshift_r(b0, b1, 16); b1 = __funnelshift_r(b1, b2, 16);
                        b2 = __funnelshift_r(b2, b3, 16); b3 = (b3 >> 16) | (val << 16);
                        pend += 1;
                        if ((op & 7u) == 7u || rem == 1u) {
                            if (pend == 8u) *reinterpret_cast<uint4*>(q.out_ml + (op & ~7u)) = make_uint4(b0, b1, b2, b3);
                            else flush_partial(q.out_ml, op, pend, b0, b1, b2, b3);
                            pend = 0;
                        }
                    }
                    op += 1; rem -= 1;
                    if (rem == 0) {
                        st = ST_IDLE;
                    } else {
                        if ((rp & 7u) == 0) need_chunk = true;
                        rp -= 1;
                        off += row_off(y);                                      // LF_move :66
                        idx = row_id<IdxT>(v);
                        st = ST_FF;
                    }
                }
            }
        }
    }
}

}  // namespace movi
