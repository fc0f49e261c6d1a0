// bs_kernel.cuh -- backward search over the move table: the count query and Ziv-Merhav lengths (ZML).
//
// Restates MoveStructure::query_backward_search / backward_search (src/move_structure_search.cpp:340-352,
// 169-201), query_zml (src/move_structure_query.cpp:690-786), backward_search_step (:311-333),
// update_interval (:48-61), initialize_backward_search (:261-293, ftab off) and MoveInterval::is_empty /
// count (include/move_intervals.hpp:43-58).
//
// Every read carries an interval with two ends (run_start:offset_start, run_end:offset_end).  Each end is a
// small walker: SCAN (update_interval: the start moves down / the end moves up to the nearest row whose
// character -- get_char, '$' for the end row -- equals the next base), then the LF jump, then FF
// (fast_forward), then READY.  One iteration of the warp loop issues one row fetch per walker for all 32
// lanes; the two walkers of a lane scan concurrently, which is equivalent to the reference's two sequential
// loops: when a matching row exists in [run_start, run_end] the start finds the first and the end finds the
// last one; when none exists the cursors cross and the interval is empty (only emptiness is observable then).
#pragma once
#include "dev_index.cuh"
#include "pml_kernel.cuh"

namespace movi {

constexpr uint32_t W_FF = 0, W_READY = 1, W_SCAN = 2;
constexpr uint32_t P_IDLE = 0, P_NEW = 1, P_SKIP = 2, P_RUN = 3, P_COUNT = 4;
constexpr int KIND_ZML = 0, KIND_COUNT = 1;

// sum of n(k) for k in [0, x): sampled prefix + at most 15 row reads
__device__ __forceinline__ unsigned long long rows_before(const DevIndex& ix, unsigned long long x) {
    unsigned long long s = ix.prefix[x >> PREFIX_SHIFT];
    for (unsigned long long k = (x >> PREFIX_SHIFT) << PREFIX_SHIFT; k < x; k++) s += row_len_at(ix, k);
    return s;
}

template <typename IdxT, int KIND>
__global__ void __launch_bounds__(256)
movi_bs_kernel(const DevIndex ix, const QueryArgs q) {
    __shared__ uint8_t s_code[256];
    __shared__ unsigned long long s_first_runs[8], s_last_runs[8];
    __shared__ uint32_t s_first_offsets[8], s_last_offsets[8];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_code[i] = ix.code_of[i];
    if (threadIdx.x < 8) {
        s_first_runs[threadIdx.x] = ix.first_runs[threadIdx.x]; s_last_runs[threadIdx.x] = ix.last_runs[threadIdx.x];
        s_first_offsets[threadIdx.x] = ix.first_offsets[threadIdx.x]; s_last_offsets[threadIdx.x] = ix.last_offsets[threadIdx.x];
    }
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const IdxT r1 = (IdxT)(ix.r - 1);
    const IdxT end_idx = (IdxT)ix.end_idx;
    const uint2* __restrict__ rows = ix.rows;
    const uint2* __restrict__ bases8 = reinterpret_cast<const uint2*>(q.bases);
    const uint32_t n_reads = q.n_reads;

    uint32_t q_next = 0, q_end = 0;
    bool exhausted = false;
    uint32_t pst = P_IDLE;
    uint32_t ri = 0;                        // read index
    IdxT rs = 0, re = 0;                    // interval ends (rows)
    uint32_t os = 0, oe = 0;                // interval ends (offsets)
    uint32_t sS = W_READY, sE = W_READY;
    bool movedS = false, movedE = false, inscan = false;
    IdxT prs = 0, pre = 0; uint32_t pos_ = 0, poe = 0;   // KIND_COUNT: interval before the current step
    uint32_t cur = 0, start = 0;            // next base to examine (index into bases), first base of the read
    uint32_t op = 0;                        // ZML: next output slot
    uint32_t m = 0;                         // ZML: current match length
    uint32_t code = 0;
    uint32_t clo = 0, chi = 0;
    bool need_chunk = false;

    for (;;) {
        const bool want = (pst == P_IDLE) && !exhausted;
        const unsigned bm = __ballot_sync(FULL, want);
        if (bm) {
            if (q_next == q_end) {
                unsigned long long b = 0;
                if (lane == 0) b = atomicAdd(q.counter, (unsigned long long)QUEUE_CHUNK);
                b = __shfl_sync(FULL, b, 0);
                q_next = b < n_reads ? (uint32_t)b : n_reads;
                q_end = b + QUEUE_CHUNK < n_reads ? (uint32_t)(b + QUEUE_CHUNK) : n_reads;
                if (q_next == q_end) exhausted = true;
            }
            const uint32_t avail = q_end - q_next;
            const uint32_t rank = __popc(bm & lt_mask);
            if (want && rank < avail) { ri = q_next + rank; pst = P_NEW; }
            const uint32_t cnt = __popc(bm);
            q_next += cnt < avail ? cnt : avail;
        }
        if (exhausted && __all_sync(FULL, pst == P_IDLE)) break;

        // ---- load slots: A = start row (or offsets[ri]), B = read chunk (or offsets[ri+1]), C = end row
        uint2 vS = make_uint2(0u, 0u), vE = make_uint2(0u, 0u), w = make_uint2(0u, 0u);
        if (pst == P_NEW) {
            vS = ldg64(reinterpret_cast<const uint2*>(q.offsets + ri));
            w = ldg64(reinterpret_cast<const uint2*>(q.offsets + ri + 1));
        } else if (pst != P_IDLE) {
            if (need_chunk) w = ldg64(bases8 + (cur >> 3));
            if (pst == P_RUN) { vS = ldg64(rows + rs); vE = ldg64(rows + re); }
        }

        if (pst == P_NEW) {
            const unsigned long long s = ((unsigned long long)vS.y << 32 | vS.x) - q.origin;
            const unsigned long long e = ((unsigned long long)w.y << 32 | w.x) - q.origin;
            start = (uint32_t)s; op = (uint32_t)s; m = 0;
            if (e == s) {                       // empty read: the reference has UB here; report 0 / nothing
                if (KIND == KIND_COUNT) { q.out_count[ri] = 0; q.out_pos[ri] = 0; }
                pst = P_IDLE;
            } else {
                cur = (uint32_t)e - 1; need_chunk = true; pst = P_SKIP;
            }
            continue;
        }
        if (pst == P_IDLE) continue;
        if (need_chunk) { clo = w.x; chi = w.y; need_chunk = false; }

        if (pst == P_SKIP) {
            // ZML: query_zml :696-704 and :756-762 -- illegal bases get 0; the first legal one seeds the interval.
            // count: query_backward_search :344-347 -- an illegal LAST base ends the query with 0.
            const uint32_t cd = s_code[__byte_perm(clo, chi, cur & 7u) & 0xFFu];
            if (cd == CODE_ILLEGAL) {
                if (KIND == KIND_COUNT) {
                    q.out_count[ri] = 0; q.out_pos[ri] = (int32_t)(cur - start + 1);
                    pst = P_IDLE;
                } else {
                    q.out_ml[op] = 0; op += 1;
                    if (cur == start) pst = P_IDLE;
                    else { if ((cur & 7u) == 0) need_chunk = true; cur -= 1; }
                }
            } else {
                const uint32_t k = cd + 1;       // initialize_backward_search :285-292
                rs = (IdxT)s_first_runs[k]; os = s_first_offsets[k];
                re = (IdxT)s_last_runs[k];  oe = s_last_offsets[k];
                sS = W_READY; sE = W_READY; inscan = false; m = 0;
                pst = P_RUN;
                // the interval now stands for base `cur`; the next base to examine is cur - 1
                if (cur == start) {
                    // nothing left to extend: ZML emits m = 0 (:763-766); count reports the seed interval
                    if (KIND == KIND_ZML) { q.out_ml[op] = 0; op += 1; pst = P_IDLE; }
                    else { prs = rs; pre = re; pos_ = os; poe = oe; cur = start; pst = P_COUNT; }
                    // count: pos_on_r == 0 here; handled in P_COUNT through `cur - start`
                } else {
                    if ((cur & 7u) == 0) need_chunk = true;
                    cur -= 1;
                }
            }
            if (pst != P_COUNT) continue;
        }

        if (pst == P_RUN) {
            const uint32_t yS = vS.y, yE = vE.y;
            // fast_forward of each end, one row per iteration (src/move_structure.cpp:524-545)
            if (sS == W_FF) { const uint32_t n = row_n(yS); if (rs < r1 && os >= n) { os -= n; rs += 1; } else sS = W_READY; }
            if (sE == W_FF) { const uint32_t n = row_n(yE); if (re < r1 && oe >= n) { oe -= n; re += 1; } else sE = W_READY; }

            bool fail = false, finished = false;
            if (sS == W_READY && sE == W_READY) {
                // both ends landed: the interval stands for base cur + 1 (or cur == start - 1 + 1 ...)
                if (inscan) {                    // previous extension succeeded
                    inscan = false;
                    if (KIND == KIND_ZML) { q.out_ml[op] = (uint16_t)(m < 65535u ? m : 65535u); op += 1; m += 1; }   // :716-719
                    if (cur == start) finished = true;          // the base just matched was the first one
                    else { if ((cur & 7u) == 0) need_chunk = true; cur -= 1; }
                }
                if (!finished) {
                    if (need_chunk) {
                        // the chunk holding the next base is not loaded yet: wait one iteration
                    } else {
                        code = s_code[__byte_perm(clo, chi, cur & 7u) & 0xFFu];
                        if (KIND == KIND_COUNT) { prs = rs; pre = re; pos_ = os; poe = oe; }   // prev_interval, :179
                        if (code == CODE_ILLEGAL) fail = true;  // backward_search_step :322-325
                        else {
                            sS = W_SCAN; sE = W_SCAN; movedS = false; movedE = false; inscan = true;
                        }
                    }
                }
            }
            if (inscan && !fail) {
                bool cross = false;
                if (sS == W_SCAN) {              // update_interval :48-54
                    const uint32_t ch = (rs == end_idx) ? 8u : row_c(yS);
                    if (ch == code) { if (movedS) os = 0; os += row_off(yS); rs = row_id<IdxT>(vS); sS = W_FF; }
                    else { rs += 1; movedS = true; }
                }
                if (sE == W_SCAN) {              // update_interval :55-61
                    const uint32_t ch = (re == end_idx) ? 8u : row_c(yE);
                    if (ch == code) { if (movedE) oe = row_n(yE) - 1; oe += row_off(yE); re = row_id<IdxT>(vE); sE = W_FF; }
                    else { if (re == 0) cross = true; else re -= 1; movedE = true; }
                }
                if (sS == W_SCAN && sE == W_SCAN && (cross || rs > re)) fail = true;   // no row with this character: empty
            }
            if (fail) {
                inscan = false;
                if (KIND == KIND_ZML) {
                    // :742-762: emit m for the current base, restart from the failing base
                    q.out_ml[op] = (uint16_t)(m < 65535u ? m : 65535u); op += 1; m = 0;
                    pst = P_SKIP;
                } else {
                    pst = P_COUNT;               // report the interval before this step; pos_on_r = cur + 1 - start
                    cur += 1;
                }
            } else if (finished) {
                if (KIND == KIND_ZML) {
                    q.out_ml[op] = (uint16_t)(m < 65535u ? m : 65535u); op += 1;     // :763-766
                    pst = P_IDLE;
                } else {
                    prs = rs; pre = re; pos_ = os; poe = oe;
                    pst = P_COUNT;               // cur == start: pos_on_r = 0
                }
            }
        }

        if (KIND == KIND_COUNT && pst == P_COUNT) {
            // MoveInterval::count (include/move_intervals.hpp:47-58) on (prs:pos_, pre:poe)
            unsigned long long cnt;
            if (prs == pre) cnt = (unsigned long long)poe - pos_ + 1;
            else {
                cnt = (unsigned long long)(row_n(rows[prs].y) - pos_) + (poe + 1);
                if (pre > prs + 1) cnt += rows_before(ix, (unsigned long long)pre) - rows_before(ix, (unsigned long long)prs + 1);
            }
            q.out_count[ri] = cnt;
            q.out_pos[ri] = (int32_t)(cur - start);
            pst = P_IDLE;
        }
    }
}

}  // namespace movi
