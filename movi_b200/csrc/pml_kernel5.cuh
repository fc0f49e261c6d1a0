// pml_kernel5.cuh -- PML kernel, variant 8: variant 5 (pml_kernel4.cuh) with batched streaming service.
//
// Variant 5 spends ~35 of its ~150 instructions per warp iteration on service code that one or two lanes need (stage the
// next 32 bytes of a read, write out 16 results, the ragged ends of a read) but the whole warp issues.  Here every lane
// owns a ring of two input sectors and a ring of 32 results, and a service event -- forced by the one lane that cannot go
// on -- is used by every lane of the warp: the lane that ran dry stages its sector and all the others prefetch theirs; the
// lane whose ring is full writes out and so do all the others.  Events come round every 30-50 iterations instead of in
// most of them.  The ragged ends are written with 4-byte pieces in a short loop instead of seven predicated pieces.
// Same state machine, same results (tests/test_gpu_parity.py runs every variant against the oracle).
//
// MEASURED (profiles/r01_explore.jsonl, "8:..."): -8 % instructions and 28 instead of 25 lanes per instruction, but the two
// rings cost 33 KB of shared memory per block; six blocks per SM then leave 23 KB of L1 and the row fetches starve
// (5.6 ms per c2 batch instead of 1.97), and at the three blocks per SM that keep the L1 at 124 KB it reaches 2.30 ms.
// Not the default; kept as the record of that experiment (pml_variant 8).
#pragma once
#include "pml_kernel4.cuh"

namespace movi {

__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

template <typename IdxT, bool SEP, int MINB, int LAYOUT, bool SECTOR = false>
__global__ void __launch_bounds__(PML3_THREADS, MINB)
movi_pml_kernel5(const DevIndex ix, const QueryArgs q) {
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
    __shared__ __align__(32) uint8_t s_in[PML3_THREADS * 64];      // per lane: ring of two 32-byte sectors of its read (slot = sector & 1)
    __shared__ __align__(32) uint16_t s_out[PML3_THREADS * 32];    // per lane: ring of 32 results (slot = output index & 31)
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
    const uint32_t sa_in = smem_addr(s_in) + threadIdx.x * 64u;       // shared-window addresses of this lane's rings
    const uint32_t sa_out = smem_addr(s_out) + threadIdx.x * 64u;
    const uint32_t sa_code = smem_addr(s_code);
    constexpr uint32_t sep = SEP ? 1u : 0u;
    constexpr uint32_t REFILL_EVERY = 4;
    const uint32_t n_reads = q.n_reads;

    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t q_next = 0, q_end = 0, chunk = min(QUEUE_CHUNK, max(1u, n_reads / (2u * n_warps)));
    bool exhausted = false;
    uint32_t st = ST_IDLE;
    IdxT idx = 0;
    uint32_t off = 0, ml = 0, rp = 0, op = 0, rem = 0;
    uint32_t fl = 0;          // results [fl, op) are still in the ring
    uint32_t in_lo = 0;       // sectors in_lo .. rp >> 5 of the read bytes are staged (at most two)

    // write results [fl, target) of this lane from the ring to global memory: whole aligned 16-value groups with two
    // 16-byte stores, the ragged ends of a read (they share their sector with the neighbouring read) piecewise
    auto flush_to = [&](uint32_t target) {
        while (fl < target) {
            const uint32_t gend = min((fl | 15u) + 1u, target);
            uint16_t* g = q.out_ml + fl;
            const uint32_t sl = sa_out + 2u * (fl & 31u);
            if (gend - fl == 16u) {
                const uint4 lo = lds_128(sl), hi = lds_128(sl + 16u);
                st_stream128(g, lo, pol_stream); st_stream128(g + 8, hi, pol_stream);
            } else {
                uint32_t k = 0; const uint32_t cnt = gend - fl;
                if ((fl & 1u) && k < cnt) { g[0] = (uint16_t)lds_u16(sl); k = 1; }
                for (; k + 2 <= cnt; k += 2) *reinterpret_cast<uint32_t*>(g + k) = lds_u32(sl + 2u * k);
                if (k < cnt) g[k] = (uint16_t)lds_u16(sl + 2u * k);
            }
            fl = gend;
        }
    };

    uint32_t tick = 0;
    uint4 sec_lo = make_uint4(0, 0, 0, 0), sec_hi = make_uint4(0, 0, 0, 0);   // SECTOR: rows 4*sec_base .. 4*sec_base + 3
    IdxT sec_base = ~(IdxT)0;
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
                rp = (uint32_t)e - 1; op = (uint32_t)s;
                idx = r1; off = ix.n_last - 1; ml = 0;
                fl = op; in_lo = (rp >> 5) + 1u;                           // nothing pending, nothing staged
                st = rem ? ST_FF : ST_IDLE;                                // (an empty read leaves the lane idle: it asks again)
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
        if (LAYOUT == LAYOUT_R8 && SECTOR) {
            const IdxT fb = fidx >> 2;
            if (proc && fb != sec_base) {
                const uint4* sp = reinterpret_cast<const uint4*>(rows) + 2 * (size_t)fb;
                sec_lo = ld_keep128(sp, pol_keep); sec_hi = ld_keep128(sp + 1, pol_keep);
                sec_base = fb;
            }
            const uint4 h = (fidx & 2) ? sec_hi : sec_lo;
            v = (fidx & 1) ? make_uint2(h.z, h.w) : make_uint2(h.x, h.y);
        }
        else if (LAYOUT == LAYOUT_R8) v = ld_keep(rows + fidx, pol_keep);
        else if (LAYOUT == LAYOUT_B6) v = fetch_row_b6<IdxT>(ix.rows6, fidx, end_idx, idb, ix.id_blocks_size, ix.block_shift, s_first);
        else {              // LAYOUT_T3: the 3-byte row without its destination id (resolved below, only when an LF step needs it)
            const uint32_t x = Rows3{ix.rows3, (unsigned long long)end_idx}.raw((unsigned long long)fidx);
            v = make_r8_raw3(0, Rows3::n_of(x), Rows3::off_of(x), Rows3::c_of(x), Rows3::thr_of(x));
        }
        // ---- service events.  Both are forced by one lane and then served for every lane of the warp that can use them, so
        // that they come round every few dozen iterations instead of in most of them:
        //   input:  a lane stepped below its staged sectors -> every walking lane stages what it is missing of its current
        //           sector and the one below it (never below the start of its read);
        //   output: a lane's ring is full -> every walking lane writes out its complete 16-value groups.
        if (__any_sync(FULL, proc && (rp >> 5) < in_lo)) {
            if (proc) {
                const uint32_t cur = rp >> 5, first_sec = (rp - rem + 1u) >> 5;
                if (cur < in_lo) {
                    const uint8_t* src = bases + ((size_t)cur << 5); const uint32_t dst = sa_in + ((cur & 1u) << 5);
                    cp_async16_stream(dst, src, pol_stream); cp_async16_stream(dst + 16u, src + 16, pol_stream);
                    in_lo = cur;
                }
                if (in_lo == cur && cur > first_sec) {
                    const uint8_t* src = bases + ((size_t)(cur - 1u) << 5); const uint32_t dst = sa_in + (((cur - 1u) & 1u) << 5);
                    cp_async16_stream(dst, src, pol_stream); cp_async16_stream(dst + 16u, src + 16, pol_stream);
                    in_lo = cur - 1u;
                }
            }
            cp_async_wait_all();
        }
        if (__any_sync(FULL, proc && (op - fl) >= 32u)) {
            if (proc) flush_to(op & ~15u);
        }

        // ---- one predicate-driven step of the state machine for every lane
        const uint32_t y = v.y;
        const uint32_t n = row_n(y);
        const uint32_t c = row_c(y);
        const uint32_t code = lds_u8(sa_code + lds_u8(sa_in + (rp & 63u)));
        const bool is_ff = (st == ST_FF);
        const bool is_scan = (st >= ST_UP);
        const bool scan_up = (st == ST_UP);
        const bool hit = (c == code);
        const bool ffmove = is_ff && (idx < r1) && (off >= n);            // fast_forward, move_structure.cpp:524-545
        const bool at_base = is_ff && !ffmove;
        const bool base_mis = at_base && !hit && (code != CODE_ILLEGAL);   // Case 2, query.cpp:297-350
        const bool scan_hit = is_scan && hit;
        const bool scan_move = is_scan && !hit;
        const bool emit = (at_base && !base_mis) || scan_hit;

        const uint32_t ai = code - sep;
        uint32_t thr = (LAYOUT == LAYOUT_R8) ? row_thr(y, ai, n) : row_thr_raw3(y, c - sep, ai, n);
        if (base_mis & ((idx == end_idx) | (SEP & (c == 0)))) {            // '$' row / '%' rows: rare, skipped warp-wide
            thr = (idx == end_idx) ? s_end_thr[ai & 3u] : sep_threshold(ix, (unsigned long long)idx, ai);
        }
        const bool down = off >= thr;
        const bool mv_down = (scan_move && !scan_up) || (base_mis && down);
        const bool mv_up = (scan_move && scan_up) || (base_mis && !down);
        const bool bad = (mv_up && idx == 0) || (mv_down && idx == r1);

        ml = (at_base && hit) ? ml + 1 : (emit ? 0u : ml);
        off = ffmove ? off - n : (scan_hit ? (scan_up ? n - 1 : 0u) : off);

        const bool fin = emit && (rem == 1u);
        const bool lf = emit && !fin;
        if (emit) sts_u16(sa_out + 2u * (op & 31u), ml < 65535u ? ml : 65535u);   // add_ml clamp
        op += emit ? 1u : 0u;
        rem -= emit ? 1u : 0u;
        rp -= lf ? 1u : 0u;
        off += lf ? row_off(y) : 0u;                                        // LF_move, move_structure.cpp:66
        IdxT lf_id = row_id<IdxT>(v);
        bool bad_id = false;
        if (LAYOUT == LAYOUT_T3) {
            // sampled-thresholds: get_id = checkpoint scan + walk-back (src/move_structure.cpp:103-219), per LF step
            if (lf) lf_id = (IdxT)sampled_get_id(Rows3{ix.rows3, (unsigned long long)end_idx}, ix.tally, ix.tally_len, ix.tally_cp, ix.r,
                                                 (unsigned long long)end_idx, (unsigned long long)idx, c, &bad_id);
        }
        idx = lf ? lf_id : (ffmove || mv_down) ? idx + 1 : (mv_up ? idx - 1 : idx);
        st = fin ? ST_IDLE : lf ? ST_FF : base_mis ? (down ? ST_DOWN : ST_UP) : st;
        if (__any_sync(FULL, fin)) {                 // end of a read: everything it still holds, ragged tail included
            if (fin) flush_to(op);
        }
        if (bad || bad_id) { atomicOr(q.status, bad_id ? 2u : 1u); st = ST_IDLE; }
    }
}

}  // namespace movi
