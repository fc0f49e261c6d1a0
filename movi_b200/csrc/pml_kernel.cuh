// pml_kernel.cuh -- pseudo-matching lengths: one read per lane, a persistent grid, one row fetch per warp iteration.
//
// Restates, for a whole batch at once, the per-base loop of MoveStructure::query_pml (src/move_structure_query.cpp:234-474):
// compare the row character with the read base, on a mismatch reposition with the threshold bit (reposition_thresholds
// :513-601, reposition_up/down :188-232), emit the matching length (MoveQuery::add_ml, include/move_query.hpp:26-38), then
// LF_move (src/move_structure.cpp:59-87) with fast_forward (:524-545).
//
// The walk is a chain of dependent random row fetches, so throughput comes from the number of independent reads in flight.
// Every lane owns one read and runs a small state machine whose every iteration issues exactly one row fetch for all lanes:
//     ST_FF        row reached by an LF jump or a fast-forward step: offset >= n ? step to idx + 1 : process the base here
//     ST_UP/DOWN   scanning for the nearest row with the read's character after a mismatch
// so the fast-forward / reposition scan of one lane never stalls the other 31 (the CPU reference hides the same latency with
// 16 "strands" per thread and software prefetch, src/read_processor.cpp:641-730).  All transitions are predicates / selects:
// one instruction stream per warp.  Finished lanes refill from a warp-local queue topped up from one global atomic (guided
// self-scheduling), a read is started inside that rarely taken branch.  Streaming: each lane stages its read bytes one
// aligned 32-byte sector at a time with cp.async (behind a warp vote) and its results 16 at a time in shared memory, written
// with two 16-byte stores, so every sector of the two streams crosses the memory system once.  History of the formulations
// that led here (branch-per-state, register-buffered output, batched service rings, ...): git log + profiles/r01_explore.jsonl.
#pragma once
#include "kernel_common.cuh"

namespace movi {

// SECTOR: keep the aligned 32-byte sector (4 rows) of the last row fetch in registers and serve fast-forward / scan steps
// that stay inside it from there.  Costs ~10 instructions per iteration, saves ~20 % of the row requests: a loss on
// L2-sized tables (issue bound), a gain on tables far beyond the L2, which are bound by the request rate of the memory
// system (profiles/r01_big_tables.jsonl) -- selected by table size in movi_gpu.cu.
// (Tried and dropped: fetching the row with cp.async.cg into shared memory instead of a load through L1 -- 5x slower,
//  profiles/r01_explore.jsonl "9:..."; and the row fetches need their L1: with the shared-memory carve-out at its maximum
//  the same kernel runs 2.8x slower, which is also what limits the per-lane staging buffers to 32 + 32 bytes.)
// RPML: `movi query --rpml` -- Case 2 repositions with reposition_randomly (src/move_structure_query.cpp:604-688) instead of
// the thresholds: up when the offset lies in the first half of the run (the reference's random draw is commented out, :607-608),
// forced at the two ends of the table, and if the scan leaves the table without finding the character it starts again from
// the same row in the other direction (:634-649, :661-674; states ST_UPF / ST_DOWNF = "already turned round").
constexpr uint32_t ST_UPF = 5, ST_DOWNF = 6;

// NT: the table of a non-threshold index type (dev_index.cuh): 12-bit field accessors; always queried with RPML.
// OUTW: matching lengths staged per lane before they are written (16 = one 32-byte sector, 32 = one 64-byte DRAM atom).  Tables
// far beyond the L2 turn the whole L2 over between two flushes of a lane, so a 32-byte flush reaches DRAM as a partial atom;
// with 32 values the write is a whole atom and there are half as many of them (the request rate is what bounds such tables).
template <typename IdxT, bool SEP, int MINB, int LAYOUT, bool SECTOR = false, bool RPML = false, bool NT = false, int OUTW = 16>
__global__ void __launch_bounds__(QUERY_THREADS, MINB)
movi_pml_kernel(const DevIndex ix, const QueryArgs q) {
    __shared__ uint8_t s_code[256];
    __shared__ uint32_t s_end_thr[4];
    __shared__ uint32_t s_idb[LAYOUT == LAYOUT_B6 ? IDB_SMEM : 1];
    __shared__ unsigned long long s_first[8];
    const uint32_t idb_n = ix.id_blocks_size * ix.sigma;
    const bool idb_in_smem = LAYOUT == LAYOUT_B6 && idb_n <= IDB_SMEM;
    if (LAYOUT == LAYOUT_B6) {
        if (idb_in_smem) for (uint32_t i = threadIdx.x; i < idb_n; i += blockDim.x) s_idb[i] = ix.id_blocks[i];
        if (threadIdx.x < 8) s_first[threadIdx.x] = ix.first_runs[threadIdx.x];
    }
    const uint32_t* const idb = idb_in_smem ? s_idb : ix.id_blocks;
    __shared__ __align__(32) uint8_t s_in[QUERY_THREADS * 32];      // per lane: the 32-byte sector of its read that holds bases[rp]
    __shared__ __align__(32) uint16_t s_out[QUERY_THREADS * OUTW];  // per lane: the sector (OUTW = 16) / atom (32) of its output being filled
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_code[i] = ix.code_of[i];
    if (threadIdx.x < 4) s_end_thr[threadIdx.x] = ix.end_thr[threadIdx.x];
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
    uint16_t* const out_slot = s_out + threadIdx.x * OUTW;            // (generic pointer: only the rare ragged flush uses it)
    constexpr uint32_t OM = OUTW - 1;
    const uint32_t sa_in = smem_addr(s_in) + threadIdx.x * 32u;       // shared-window addresses of this lane's slots
    const uint32_t sa_out = smem_addr(s_out) + threadIdx.x * (2u * OUTW);
    const uint32_t sa_code = smem_addr(s_code);
    constexpr uint32_t sep = SEP ? 1u : 0u;
    constexpr uint32_t REFILL_EVERY = 4;
    const uint32_t n_reads = q.n_reads;

    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t q_next = 0, q_end = 0, chunk = min(QUEUE_CHUNK, max(1u, n_reads / (2u * n_warps)));
    bool exhausted = false;
    uint32_t st = ST_IDLE;
    IdxT idx = 0;
    uint32_t off = 0, ml = 0, rp = 0, op = 0, rem = 0, pend = 0;
    bool need_chunk = false;

    uint32_t tick = 0;
    uint4 sec_lo = make_uint4(0, 0, 0, 0), sec_hi = make_uint4(0, 0, 0, 0);   // SECTOR: rows 4*sec_base .. 4*sec_base + 3
    IdxT sec_base = ~(IdxT)0;
    IdxT saved = 0;                                                           // RPML: the row the reposition started from
    for (;;) {
        // idle lanes are looked after every REFILL_EVERY-th iteration only: a finished lane waits <= 3 of the ~240
        // iterations of a 150-base read, and the other iterations skip the vote / queue bookkeeping
        const bool service = (tick++ & (REFILL_EVERY - 1u)) == 0u;
        const bool want = service && (st == ST_IDLE) && !exhausted;
        const unsigned m = service ? __ballot_sync(FULL, want) : 0u;
        if (m) {
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
            const uint32_t rank = __popc(m & lt_mask);
            if (want && rank < avail) {                                   // start a read (query_pml :237-240)
                const uint32_t ri = q_next + rank;
                const uint2 a = ldg64(offs8 + ri), b = ldg64(offs8 + ri + 1);
                const unsigned long long s = (((unsigned long long)a.y << 32) | a.x) - q.origin;
                const unsigned long long e = (((unsigned long long)b.y << 32) | b.x) - q.origin;
                rem = (uint32_t)(e - s);
                if (e >> 32) { atomicOr(q.status, 4u); rem = 0; }          // per-read cursors are 32-bit: a device call holds < 2^32 bases
                rp = (uint32_t)e - 1; op = (uint32_t)s;
                idx = r1; off = ix.n_last - 1; ml = 0; pend = 0;
                st = rem ? ST_FF : ST_IDLE;                                // (an empty read leaves the lane idle: it asks again)
                need_chunk = true;
            }
            const uint32_t cnt = __popc(m);
            q_next += cnt < avail ? cnt : avail;
        }
        if (service && exhausted && __all_sync(FULL, st == ST_IDLE)) break;

        // ---- loads: row fetch for every walking lane; a lane that ran out of staged read bytes pulls the next sector
        const bool proc = (st != ST_IDLE);
        // (lanes that are not walking fetch row 0: harmless and keeps the fetch out of a divergent branch)
        const IdxT fidx = proc ? idx : (IdxT)0;
        uint2 v;
        uint32_t x4 = 0;                                              // LAYOUT_T4: the row word (its top byte is needed at the LF step)
        if (LAYOUT == LAYOUT_R8 && SECTOR) {
            const IdxT fb = fidx >> 2;
            if (proc && fb != sec_base) {
                const uint4* sp = reinterpret_cast<const uint4*>(rows) + 2 * (size_t)fb;
                ld_keep256(sp, pol_keep, sec_lo, sec_hi);
                sec_base = fb;
            }
            const uint4 h = (fidx & 2) ? sec_hi : sec_lo;
            v = (fidx & 1) ? make_uint2(h.z, h.w) : make_uint2(h.x, h.y);
        }
        else if (LAYOUT == LAYOUT_R8) v = ld_keep(rows + fidx, pol_keep);
        else if (LAYOUT == LAYOUT_B6) v = fetch_row_b6<IdxT>(ix.rows6, fidx, end_idx, idb, ix.id_blocks_size, ix.block_shift, s_first);
        else {              // LAYOUT_T4: the 4-byte row; its destination id is resolved below, only when an LF step needs it
            asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(x4) : "l"(ix.rows4 + fidx), "l"(pol_keep));
            v = make_r8_raw3(0, Rows4::n_of(x4), Rows4::off_of(x4), Rows4::c_of(x4), Rows4::thr_of(x4));
        }
        if (__any_sync(FULL, proc && need_chunk)) {
            if (proc && need_chunk) {
                const uint8_t* src = bases + (rp & ~31u);
                cp_async16_stream(sa_in, src, pol_stream); cp_async16_stream(sa_in + 16u, src + 16, pol_stream);
            }
            cp_async_wait_all();
            need_chunk = need_chunk && !proc;
        }

        // ---- one predicate-driven step of the state machine for every lane
        const uint32_t y = v.y;
        const uint32_t n = row_n<NT>(y);
        const uint32_t c = row_c(y);
        const uint32_t code = lds_u8(sa_code + lds_u8(sa_in + (rp & 31u)));
        const bool is_ff = (st == ST_FF);
        const bool is_scan = (st >= ST_UP);
        const bool scan_up = RPML ? ((st & 1u) != 0u) : (st == ST_UP);           // (ST_UP = 3, ST_UPF = 5)
        const bool hit = (c == code);
        const bool ffmove = is_ff && (idx < r1) && (off >= n);            // fast_forward, move_structure.cpp:524-545
        const bool at_base = is_ff && !ffmove;
        const bool base_mis = at_base && !hit && (code != CODE_ILLEGAL);   // Case 2, query.cpp:297-350
        const bool scan_hit = is_scan && hit;
        const bool scan_move = is_scan && !hit;
        const bool emit = (at_base && !base_mis) || scan_hit;

        const uint32_t ai = code - sep;
        uint32_t thr = (LAYOUT == LAYOUT_R8) ? row_thr(y, ai, n) : row_thr_raw3(y, c - sep, ai, n);
        if (!RPML && base_mis & ((idx == end_idx) | (SEP & (c == 0)))) {            // '$' row / '%' rows: rare, skipped warp-wide
            thr = (idx == end_idx) ? s_end_thr[ai & 3u] : sep_threshold(ix, (unsigned long long)idx, ai);
        }
        const bool down = RPML ? !((idx == r1) || (idx > 0 && off * 2u < n)) : (off >= thr);
        const bool mv_down = (scan_move && !scan_up) || (base_mis && down);
        const bool mv_up = (scan_move && scan_up) || (base_mis && !down);
        bool bad = (mv_up && idx == 0) || (mv_down && idx == r1);

        ml = (at_base && hit) ? ml + 1 : (emit ? 0u : ml);
        off = ffmove ? off - n : (scan_hit ? (scan_up ? n - 1 : 0u) : off);

        const bool fin = emit && (rem == 1u);
        const bool lf = emit && !fin;
        if (emit) {
            sts_u16(sa_out + 2u * (op & OM), ml < 65535u ? ml : 65535u);   // add_ml clamp
            pend += 1;
            if ((op & OM) == OM || fin) {
                uint16_t* g = q.out_ml + (op & ~OM);
                if (pend == (uint32_t)OUTW) {
#pragma unroll
                    for (uint32_t k = 0; k < (uint32_t)OUTW / 8u; k++) st_stream128(g + 8u * k, lds_128(sa_out + 16u * k), pol_stream);
                } else {
                    // ragged head / tail of a read: only our own slots [lo, hi1) of the staged group, widest aligned pieces first
                    uint32_t lo = (op & OM) + 1u - pend;
                    const uint32_t hi1 = (op & OM) + 1u;                   // one past the last slot
                    while (lo < hi1) {
                        if (!(lo & 7u) && lo + 8u <= hi1) { *reinterpret_cast<uint4*>(g + lo) = *reinterpret_cast<const uint4*>(out_slot + lo); lo += 8u; }
                        else if (!(lo & 3u) && lo + 4u <= hi1) { *reinterpret_cast<uint2*>(g + lo) = *reinterpret_cast<const uint2*>(out_slot + lo); lo += 4u; }
                        else if (!(lo & 1u) && lo + 2u <= hi1) { *reinterpret_cast<uint32_t*>(g + lo) = *reinterpret_cast<const uint32_t*>(out_slot + lo); lo += 2u; }
                        else { g[lo] = out_slot[lo]; lo += 1u; }
                    }
                }
                pend = 0;
            }
        }
        op += emit ? 1u : 0u;
        rem -= emit ? 1u : 0u;
        need_chunk = need_chunk || (lf && (rp & 31u) == 0u);
        rp -= lf ? 1u : 0u;
        off += lf ? row_off<NT>(y) : 0u;                                        // LF_move, move_structure.cpp:66
        IdxT lf_id = row_id<IdxT, NT>(v);
        bool bad_id = false;
        if (LAYOUT == LAYOUT_T4) {
            // sampled-thresholds: get_id = checkpoint scan + walk-back (src/move_structure.cpp:103-219), per LF step
            if (lf) lf_id = t4_get_id<IdxT>(ix, idx, end_idx, x4, &bad_id);
        }
        if (RPML) saved = base_mis ? idx : saved;
        idx = lf ? lf_id : (ffmove || mv_down) ? idx + 1 : (mv_up ? idx - 1 : idx);
        st = fin ? ST_IDLE : lf ? ST_FF : base_mis ? (down ? ST_DOWN : ST_UP) : st;
        if (RPML) {
            if (bad && st <= ST_DOWN && (mv_up ? saved < r1 : saved > 0)) {     // first scan left the table: turn round
                idx = mv_up ? saved + 1 : saved - 1;
                st = mv_up ? ST_DOWNF : ST_UPF;
                bad = false;
            }
        }
        if (bad || bad_id) { atomicOr(q.status, bad_id ? 2u : 1u); st = ST_IDLE; }
    }
}

}  // namespace movi
