// bs_kernel.cuh -- backward search over the move table: the count query and Ziv-Merhav lengths (ZML).
//
// Restates query_backward_search / backward_search (src/move_structure_search.cpp:340-352, 169-201), query_zml
// (src/move_structure_query.cpp:690-786), backward_search_step (:311-333), update_interval (:48-61),
// initialize_backward_search (:261-293, ftab off), MoveInterval::is_empty / count (include/move_intervals.hpp:43-58).
//
// Every read carries an interval with two ends (run_start:offset_start, run_end:offset_end).  Each end is a small walker:
// SCAN (update_interval: the start moves down / the end moves up to the nearest row whose character -- get_char, '$' for the
// end row -- equals the next base), then the LF jump, then FF (fast_forward), then READY.  One iteration of the warp loop
// issues one row fetch per walker for all 32 lanes; the two walkers of a lane scan concurrently, which is equivalent to the
// reference's two sequential loops: when a matching row exists in [run_start, run_end] the start finds the first and the end
// finds the last one; when none exists the cursors cross and the interval is empty (only emptiness is observable then).
//   * the per-iteration step of the two interval ends is computed with predicates for all 32 lanes at once; only the rare
//     events (seeding an interval, a failed extension, the end of a read, the final count) are branches;
//   * read bytes are staged one aligned 32-byte sector at a time with cp.async, ZML values leave 16 at a time through
//     shared memory, and every iteration performs at most ONE emit (a failed extension or the end of a read is emitted by
//     the next iteration), so there is a single store site.
#pragma once
#include "kernel_common.cuh"

namespace movi {

constexpr uint32_t W_FF = 0, W_READY = 1, W_SCAN = 2;
constexpr int KIND_ZML = 0, KIND_COUNT = 1;

// sum of n(k) for k in [0, x): sampled prefix + at most 15 row reads
__device__ __forceinline__ unsigned long long rows_before(const DevIndex& ix, unsigned long long x) {
    unsigned long long s = ix.prefix[x >> PREFIX_SHIFT];
    for (unsigned long long k = (x >> PREFIX_SHIFT) << PREFIX_SHIFT; k < x; k++) s += row_len_at(ix, k);
    return s;
}

constexpr uint32_t B_IDLE = 0, B_NEW = 1, B_SKIP = 2, B_RUN = 3, B_FAIL = 4, B_FIN = 5, B_COUNT = 6;

// SECTOR (tables far beyond the L2, which are bound by the request rate of the memory system -- see pml_kernel.cuh): each
// interval end keeps the aligned 32-byte sector (4 rows) of its last fetch in registers; scan and fast-forward steps that
// stay inside it, and the very common case of both ends sitting in the same rows, cost no new request.
template <typename IdxT, int KIND, int LAYOUT, bool SECTOR = false, bool NT = false>
__global__ void __launch_bounds__(QUERY_THREADS, SECTOR ? 3 : 5)
movi_bs_kernel(const DevIndex ix, const QueryArgs q) {
    __shared__ uint8_t s_code[256];
    __shared__ uint32_t s_idb[LAYOUT == LAYOUT_B6 ? IDB_SMEM : 1];
    const uint32_t idb_n = ix.id_blocks_size * ix.sigma;
    const bool idb_in_smem = LAYOUT == LAYOUT_B6 && idb_n <= IDB_SMEM;
    if (LAYOUT == LAYOUT_B6 && idb_in_smem) for (uint32_t i = threadIdx.x; i < idb_n; i += blockDim.x) s_idb[i] = ix.id_blocks[i];
    const uint32_t* const idb = idb_in_smem ? s_idb : ix.id_blocks;
    __shared__ unsigned long long s_first_runs[8], s_last_runs[8];
    __shared__ uint32_t s_first_offsets[8], s_last_offsets[8];
    __shared__ __align__(32) uint8_t s_in[QUERY_THREADS * 32];
    __shared__ __align__(32) uint16_t s_out[KIND == KIND_ZML ? QUERY_THREADS * 16 : 16];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_code[i] = ix.code_of[i];
    if (threadIdx.x < 8) {
        s_first_runs[threadIdx.x] = ix.first_runs[threadIdx.x]; s_last_runs[threadIdx.x] = ix.last_runs[threadIdx.x];
        s_first_offsets[threadIdx.x] = ix.first_offsets[threadIdx.x]; s_last_offsets[threadIdx.x] = ix.last_offsets[threadIdx.x];
    }
    __syncthreads();

    const unsigned long long pol_keep = policy_evict_last();
    const unsigned long long pol_stream = policy_evict_first();
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const IdxT r1 = (IdxT)(ix.r - 1);
    const IdxT end_idx = (IdxT)ix.end_idx;
    const uint2* __restrict__ rows = ix.rows;
    const uint8_t* __restrict__ bases = q.bases;
    const uint2* __restrict__ offs8 = reinterpret_cast<const uint2*>(q.offsets);
    uint8_t* const in_slot = s_in + threadIdx.x * 32;
    const uint32_t sa_in = smem_addr(s_in) + threadIdx.x * 32u;
    uint16_t* const out_slot = s_out + (KIND == KIND_ZML ? threadIdx.x * 16 : 0);
    const uint32_t n_reads = q.n_reads;

    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t q_next = 0, q_end = 0, chunk = min(QUEUE_CHUNK, max(1u, n_reads / (2u * n_warps)));
    bool exhausted = false;
    uint32_t pst = B_IDLE;
    uint32_t ri = 0;
    IdxT rs = 0, re = 0;
    uint32_t os = 0, oe = 0;
    uint32_t sS = W_READY, sE = W_READY;
    bool movedS = false, movedE = false, inscan = false;
    IdxT prs = 0, pre = 0; uint32_t pos_ = 0, poe = 0;   // KIND_COUNT: the interval before the current step
    uint32_t cur = 0, start = 0, op = 0, m = 0, code = 0, pend = 0;
    bool need_chunk = false;

    uint4 loS = make_uint4(0, 0, 0, 0), hiS = loS, loE = loS, hiE = loS;      // SECTOR: rows 4*base .. 4*base + 3 of each end
    IdxT baseS = ~(IdxT)0, baseE = ~(IdxT)0;
    constexpr uint32_t REFILL_EVERY = 4;      // idle lanes are looked after every 4th iteration only (see pml_kernel.cuh)
    uint32_t tick = 0;
    for (;;) {
        const bool service = (tick++ & (REFILL_EVERY - 1u)) == 0u;
        const bool want = service && (pst == B_IDLE) && !exhausted;
        const unsigned bm = service ? __ballot_sync(FULL, want) : 0u;
        if (bm) {
            if (q_next == q_end) {
                // guided self-scheduling: take 32 reads per refill while plenty remain, fewer as the batch drains, so that
                // no warp sits on a private backlog while the lanes of other warps have already gone idle
                unsigned long long b = 0;
                if (lane == 0) b = atomicAdd(q.counter, (unsigned long long)chunk);
                b = __shfl_sync(FULL, b, 0);
                q_next = b < n_reads ? (uint32_t)b : n_reads;
                q_end = b + chunk < n_reads ? (uint32_t)(b + chunk) : n_reads;
                if (q_next == q_end) exhausted = true;
                const uint32_t left = n_reads - q_end;
                chunk = min(QUEUE_CHUNK, max(1u, left / (2u * n_warps)));
            }
            const uint32_t avail = q_end - q_next;
            const uint32_t rank = __popc(bm & lt_mask);
            if (want && rank < avail) {          // start a read where it is handed out (rare branch)
                ri = q_next + rank;
                const uint2 a = ldg64(offs8 + ri), b = ldg64(offs8 + ri + 1);
                const unsigned long long s = (((unsigned long long)a.y << 32) | a.x) - q.origin;
                const unsigned long long e = (((unsigned long long)b.y << 32) | b.x) - q.origin;
                start = (uint32_t)s; op = (uint32_t)s; m = 0; pend = 0; inscan = false;
                if (e >> 32) atomicOr(q.status, 4u);   // per-read cursors are 32-bit: a device call holds < 2^32 bases
                if (e == s || (e >> 32)) {          // empty read: outside the reference's domain (it indexes R[-1])
                    if (KIND == KIND_COUNT) { q.out_count[ri] = 0; q.out_pos[ri] = 0; }
                    pst = B_IDLE;
                } else {
                    cur = (uint32_t)e - 1; need_chunk = true; pst = B_SKIP;
                }
            }
            const uint32_t cnt = __popc(bm);
            q_next += cnt < avail ? cnt : avail;
        }
        if (service && exhausted && __all_sync(FULL, pst == B_IDLE)) break;

        // ---- loads
        const bool run = (pst == B_RUN);
        const bool chunk_user = run || (pst == B_SKIP);
        uint2 vS = make_uint2(0u, 0u), vE = make_uint2(0u, 0u);
        if (SECTOR && LAYOUT == LAYOUT_R8) {
            if (run) {
                const IdxT fbS = rs >> 2, fbE = re >> 2;
                if (fbS != baseS) {
                    const uint4* sp = reinterpret_cast<const uint4*>(rows) + 2 * (size_t)fbS;
                    ld_keep256(sp, pol_keep, loS, hiS); baseS = fbS;
                }
                if (fbE != baseE) {
                    if (fbE == baseS) { loE = loS; hiE = hiS; }
                    else { const uint4* sp = reinterpret_cast<const uint4*>(rows) + 2 * (size_t)fbE;
                           ld_keep256(sp, pol_keep, loE, hiE); }
                    baseE = fbE;
                }
                const uint4 hS = (rs & 2) ? hiS : loS, hE = (re & 2) ? hiE : loE;
                vS = (rs & 1) ? make_uint2(hS.z, hS.w) : make_uint2(hS.x, hS.y);
                vE = (re & 1) ? make_uint2(hE.z, hE.w) : make_uint2(hE.x, hE.y);
            }
        } else if (run) {
            if (LAYOUT == LAYOUT_R8) { vS = ld_keep(rows + rs, pol_keep); vE = ld_keep(rows + re, pol_keep); }
            else {
                // first_runs indexed by code + 1 is exactly what get_id needs (first_runs[c + 1])
                vS = fetch_row_b6<IdxT>(ix.rows6, rs, end_idx, idb, ix.id_blocks_size, ix.block_shift, s_first_runs);
                vE = fetch_row_b6<IdxT>(ix.rows6, re, end_idx, idb, ix.id_blocks_size, ix.block_shift, s_first_runs);
            }
        }
        const bool staged = chunk_user && need_chunk;
        if (__any_sync(FULL, staged)) {          // about half of the iterations have no lane that needs a new sector
            if (staged) {
                const uint8_t* src = bases + (cur & ~31u);
                cp_async16_stream(sa_in, src, pol_stream);            // streamed data: evict_first in L2 (pml_kernel.cuh)
                cp_async16_stream(sa_in + 16u, src + 16, pol_stream);
            }
            cp_async_wait_all();
            if (staged) need_chunk = false;
        }

        bool do_emit = false, last = false;
        uint32_t val = 0;

        // ---- rare states (branches): seed / skip illegal bases, deferred emits, final count
        if (pst == B_SKIP && !need_chunk && chunk_user) {
            // ZML: query_zml :696-704 and :756-762 -- illegal bases get 0, the first legal one seeds the interval.
            // count: query_backward_search :344-347 -- an illegal LAST base ends the query with 0.
            const uint32_t cd = s_code[in_slot[cur & 31u]];
            if (cd == CODE_ILLEGAL) {
                if (KIND == KIND_COUNT) {
                    q.out_count[ri] = 0; q.out_pos[ri] = (int32_t)(cur - start + 1);
                    pst = B_IDLE;
                } else {
                    do_emit = true; val = 0; last = (cur == start);
                    if (last) pst = B_IDLE;
                    else { need_chunk = (cur & 31u) == 0u; cur -= 1; }
                }
            } else {
                const uint32_t k = cd + 1;       // initialize_backward_search :285-292
                rs = (IdxT)s_first_runs[k]; os = s_first_offsets[k];
                re = (IdxT)s_last_runs[k];  oe = s_last_offsets[k];
                sS = W_READY; sE = W_READY; inscan = false; m = 0;
                if (cur == start) {              // nothing left to extend
                    if (KIND == KIND_ZML) pst = B_FIN;
                    else { prs = rs; pre = re; pos_ = os; poe = oe; pst = B_COUNT; }
                } else {
                    need_chunk = (cur & 31u) == 0u; cur -= 1; pst = B_RUN;
                }
            }
        } else if (pst == B_FAIL) {              // ZML :742-762: the failed extension closes the current match
            do_emit = true; val = m; m = 0; pst = B_SKIP;
        } else if (pst == B_FIN) {               // ZML :763-766
            do_emit = true; val = m; last = true; pst = B_IDLE;
        }

        // ---- the walking state: both interval ends step under predicates (lanes with !run change nothing)
        {
            const uint32_t yS = vS.y, yE = vE.y;
            const uint32_t nS = row_n<NT>(yS), nE = row_n<NT>(yE);
            // fast_forward of each end (src/move_structure.cpp:524-545)
            const bool ffS = run && (sS == W_FF), ffE = run && (sE == W_FF);
            const bool stepS = ffS && (rs < r1) && (os >= nS), stepE = ffE && (re < r1) && (oe >= nE);
            os -= stepS ? nS : 0u; rs += stepS ? 1 : 0; sS = (ffS && !stepS) ? W_READY : sS;
            oe -= stepE ? nE : 0u; re += stepE ? 1 : 0; sE = (ffE && !stepE) ? W_READY : sE;

            const bool both = run && (sS == W_READY) && (sE == W_READY);
            const bool succ = both && inscan;                       // the previous extension landed: :716-719
            const bool fin_now = succ && (cur == start);
            if (KIND == KIND_ZML) { if (succ) { do_emit = true; val = m; m += 1; } }
            inscan = inscan && !succ;
            const bool adv = succ && !fin_now;
            need_chunk = need_chunk || (adv && (cur & 31u) == 0u);
            cur -= adv ? 1u : 0u;

            const bool exam = both && !fin_now && !need_chunk;      // look at the next base (needs its sector staged)
            const uint32_t code_new = s_code[in_slot[cur & 31u]];
            code = exam ? code_new : code;
            if (KIND == KIND_COUNT) { if (exam || fin_now) { prs = rs; pre = re; pos_ = os; poe = oe; } }   // prev_interval, search.cpp:179
            const bool ill = exam && (code_new == CODE_ILLEGAL);    // backward_search_step :322-325
            const bool go = exam && !ill;
            sS = go ? W_SCAN : sS; sE = go ? W_SCAN : sE;
            movedS = movedS && !go; movedE = movedE && !go;
            inscan = inscan || go;

            // update_interval (search.cpp:48-61): start moves down, end moves up to the nearest row with the character
            const bool scS = run && (sS == W_SCAN), scE = run && (sE == W_SCAN);
            const uint32_t chS = (rs == end_idx) ? 8u : row_c(yS), chE = (re == end_idx) ? 8u : row_c(yE);
            const bool hitS = scS && (chS == code), hitE = scE && (chE == code);
            const bool mvS = scS && !hitS, mvE = scE && !hitE;
            os = hitS ? (movedS ? 0u : os) + row_off<NT>(yS) : os;
            oe = hitE ? (movedE ? nE - 1u : oe) + row_off<NT>(yE) : oe;
            const bool cross0 = mvE && (re == 0);
            rs = hitS ? row_id<IdxT, NT>(vS) : (mvS ? rs + 1 : rs);
            re = hitE ? row_id<IdxT, NT>(vE) : ((mvE && !cross0) ? re - 1 : re);
            sS = hitS ? W_FF : sS; sE = hitE ? W_FF : sE;
            movedS = movedS || mvS; movedE = movedE || mvE;
            // no row with this character between the two cursors: the interval is empty
            const bool cross = mvS && mvE && (cross0 || rs > re);
            const bool fail = ill || cross;
            inscan = inscan && !fail;
            if (KIND == KIND_ZML) pst = fail ? B_FAIL : (fin_now ? B_FIN : pst);
            else { cur += fail ? 1u : 0u; pst = (fail || fin_now) ? B_COUNT : pst; }   // pos_on_r = cur + 1 - start on a failure
        }

        if (KIND == KIND_ZML) {
            if (do_emit) {
                out_slot[op & 15u] = (uint16_t)(val < 65535u ? val : 65535u);
                pend += 1;
                if ((op & 15u) == 15u || last) {
                    uint16_t* g = q.out_ml + (op & ~15u);
                    if (pend == 16u) {
                        const uint4 lo4 = *reinterpret_cast<const uint4*>(out_slot);
                        const uint4 hi4 = *reinterpret_cast<const uint4*>(out_slot + 8);
                        st_stream128(g, lo4, pol_stream);
                        st_stream128(g + 8, hi4, pol_stream);
                    } else {
                        uint32_t lo = (op & 15u) + 1u - pend;
                        const uint32_t hi1 = (op & 15u) + 1u;
                        if ((lo & 1u) && lo < hi1) { g[lo] = out_slot[lo]; lo += 1; }
                        if ((lo & 2u) && lo + 2 <= hi1) { *reinterpret_cast<uint32_t*>(g + lo) = *reinterpret_cast<const uint32_t*>(out_slot + lo); lo += 2; }
                        if ((lo & 4u) && lo + 4 <= hi1) { *reinterpret_cast<uint2*>(g + lo) = *reinterpret_cast<const uint2*>(out_slot + lo); lo += 4; }
                        if (lo + 8 <= hi1) { *reinterpret_cast<uint4*>(g + lo) = *reinterpret_cast<const uint4*>(out_slot + lo); lo += 8; }
                        if (lo + 4 <= hi1) { *reinterpret_cast<uint2*>(g + lo) = *reinterpret_cast<const uint2*>(out_slot + lo); lo += 4; }
                        if (lo + 2 <= hi1) { *reinterpret_cast<uint32_t*>(g + lo) = *reinterpret_cast<const uint32_t*>(out_slot + lo); lo += 2; }
                        if (lo < hi1) g[lo] = out_slot[lo];
                    }
                    pend = 0;
                }
                op += 1;
            }
        } else if (pst == B_COUNT) {
            // MoveInterval::count (include/move_intervals.hpp:47-58) on (prs:pos_, pre:poe)
            unsigned long long cnt;
            if (prs == pre) cnt = (unsigned long long)poe - pos_ + 1;
            else {
                cnt = (unsigned long long)(row_len_at(ix, (unsigned long long)prs) - pos_) + (poe + 1);
                if (pre > prs + 1) cnt += rows_before(ix, (unsigned long long)pre) - rows_before(ix, (unsigned long long)prs + 1);
            }
            q.out_count[ri] = cnt;
            q.out_pos[ri] = (int32_t)(cur - start);
            pst = B_IDLE;
        }
    }
}

}  // namespace movi
