// pml_kernel3.cuh -- PML kernel, variant 4: convergent state machine + sector-granular streaming through shared memory.
//
// The ncu captures of variants 1 and 2 on the c2 workload (profiles/r01_*.md) showed ~74 bytes of DRAM read per base
// where one 64-byte atom per base (the row fetch) is the floor, and 2.1x write amplification: every lane touched its
// read bytes 8 bytes at a time and its output 8 bytes at a time, tens of microseconds apart, while the row fetches
// turn the whole 126 MB L2 over every ~35 us -- so each 64-byte atom of the read / output streams was fetched or
// written back several times.  Here every lane stages
//   * its read bytes one aligned 32-byte sector at a time, global -> shared with cp.async (no registers, L1 bypass),
//   * its PML values 16 at a time (one aligned 32-byte sector) in shared memory, written out with two 16-byte stores,
// so each sector of the two streams crosses the memory system once.  Per-base cost drops as well: the base is one
// LDS.U8 instead of a byte permute, the emit is one STS.U16 instead of a register shift chain.
//
// State machine and reference citations: see pml_kernel.cuh / pml_kernel2.cuh (identical semantics).
#pragma once
#include "dev_index.cuh"
#include "pml_kernel.cuh"
#include "pml_kernel2.cuh"

namespace movi {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int PML3_THREADS = 256;

// SEP: the index has separators ("%ACGT"; threshold slots shift by one, '%' rows use a side table) -- compiled out otherwise
// MINB: resident blocks per SM the register allocation must allow.  6 (40 registers) is the default; 8 (32 registers,
// 2048 walks per SM, one spilled word) was measured slower both on the L2-sized c2 table (issue bound) and on 8-64 GB
// tables (bound by the memory system's random-request rate, not by exposed latency): profiles/r01_big_tables.jsonl.
// LAYOUT: LAYOUT_R8 (one 8-byte fetch per step), LAYOUT_B6 (6-byte blocked rows + id_blocks, 25 % less HBM) or
// LAYOUT_T3 (3-byte sampled rows + id checkpoints, half the HBM; the destination id costs a scan + walk per LF step).
constexpr int IDB_SMEM = 1024;       // id_blocks entries kept in shared memory (covers r <= 2.7e8 rows at block_size 2^20)

template <typename IdxT, bool SEP, int MINB, int LAYOUT>
__global__ void __launch_bounds__(PML3_THREADS, MINB)
movi_pml_kernel3(const DevIndex ix, const QueryArgs q) {
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
    __shared__ __align__(32) uint8_t s_in[PML3_THREADS * 32];      // per lane: the 32-byte sector of its read that holds bases[rp]
    __shared__ __align__(32) uint16_t s_out[PML3_THREADS * 16];    // per lane: the 32-byte sector of its output being filled
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_code[i] = ix.code_of[i];
    if (threadIdx.x < 4) s_end_thr[threadIdx.x] = ix.end_thr[threadIdx.x];
    __syncthreads();

    const unsigned long long pol_keep = policy_evict_last();
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const IdxT r1 = (IdxT)(ix.r - 1);
    const IdxT end_idx = (IdxT)ix.end_idx;
    const uint2* __restrict__ rows = ix.rows;
    const uint8_t* __restrict__ bases = q.bases;
    const uint2* __restrict__ offs8 = reinterpret_cast<const uint2*>(q.offsets);
    uint8_t* const in_slot = s_in + threadIdx.x * 32;
    uint16_t* const out_slot = s_out + threadIdx.x * 16;
    constexpr uint32_t sep = SEP ? 1u : 0u;
    const uint32_t n_reads = q.n_reads;

    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t q_next = 0, q_end = 0, chunk = min(QUEUE_CHUNK, max(1u, n_reads / (2u * n_warps)));
    bool exhausted = false;
    uint32_t st = ST_IDLE;
    IdxT idx = 0;
    uint32_t off = 0, ml = 0, rp = 0, op = 0, rem = 0, pend = 0;
    bool need_chunk = false;

    for (;;) {
        const bool want = (st == ST_IDLE) && !exhausted;
        const unsigned m = __ballot_sync(FULL, want);
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
            if (want && rank < avail) { idx = (IdxT)(q_next + rank); st = ST_NEW; }
            const uint32_t cnt = __popc(m);
            q_next += cnt < avail ? cnt : avail;
        }
        if (exhausted && __all_sync(FULL, st == ST_IDLE)) break;

        // ---- loads: row fetch for every walking lane; a lane that ran out of staged read bytes pulls the next sector
        const bool is_new = (st == ST_NEW);
        const bool proc = (st >= ST_FF);
        // (lanes that are not walking fetch row 0: harmless and keeps the fetch out of a divergent branch)
        const IdxT fidx = proc ? idx : (IdxT)0;
        uint2 v;
        if (LAYOUT == LAYOUT_R8) v = ld_keep(rows + fidx, pol_keep);
        else if (LAYOUT == LAYOUT_B6) v = fetch_row_b6<IdxT>(ix.rows6, fidx, end_idx, idb, ix.id_blocks_size, ix.block_shift, s_first);
        else {              // LAYOUT_T3: the 3-byte row without its destination id (resolved below, only when an LF step needs it)
            const uint32_t x = Rows3{ix.rows3, (unsigned long long)end_idx}.raw((unsigned long long)fidx);
            v = make_r8_raw3(0, Rows3::n_of(x), Rows3::off_of(x), Rows3::c_of(x), Rows3::thr_of(x));
        }
        if (proc && need_chunk) {
            const uint8_t* src = bases + (rp & ~31u);
            cp_async16(in_slot, src);
            cp_async16(in_slot + 16, src + 16);
        }
        if (is_new) {       // rare per lane: start a read (query_pml :237-240)
            const uint2 a = ldg64(offs8 + idx), b = ldg64(offs8 + idx + 1);
            const unsigned long long s = (((unsigned long long)a.y << 32) | a.x) - q.origin;
            const unsigned long long e = (((unsigned long long)b.y << 32) | b.x) - q.origin;
            rem = (uint32_t)(e - s);
            rp = (uint32_t)e - 1; op = (uint32_t)s;
            idx = r1; off = ix.n_last - 1; ml = 0; pend = 0;
            st = rem ? ST_FF : ST_IDLE;
            need_chunk = true;
        }
        cp_async_wait_all();
        need_chunk = need_chunk && !proc;

        // ---- one predicate-driven step of the state machine for every lane
        const uint32_t y = v.y;
        const uint32_t n = row_n(y);
        const uint32_t c = row_c(y);
        const uint32_t code = s_code[in_slot[rp & 31u]];
        const bool is_ff = proc && (st == ST_FF);
        const bool is_scan = proc && (st >= ST_UP);
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
        if (emit) {
            out_slot[op & 15u] = (uint16_t)(ml < 65535u ? ml : 65535u);    // add_ml clamp
            pend += 1;
            if ((op & 15u) == 15u || fin) {
                uint16_t* g = q.out_ml + (op & ~15u);
                if (pend == 16u) {
                    const uint4 lo = *reinterpret_cast<const uint4*>(out_slot);
                    const uint4 hi = *reinterpret_cast<const uint4*>(out_slot + 8);
                    *reinterpret_cast<uint4*>(g) = lo;
                    *reinterpret_cast<uint4*>(g + 8) = hi;
                } else {
                    // ragged head / tail of a read: only our own slots [lo, hi] of the sector, widest aligned pieces first
                    uint32_t lo = (op & 15u) + 1u - pend;
                    const uint32_t hi1 = (op & 15u) + 1u;                  // one past the last slot
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
        }
        op += emit ? 1u : 0u;
        rem -= emit ? 1u : 0u;
        need_chunk = need_chunk || (lf && (rp & 31u) == 0u);
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
        if (bad || bad_id) { atomicOr(q.status, bad_id ? 2u : 1u); st = ST_IDLE; }
    }
}

}  // namespace movi
