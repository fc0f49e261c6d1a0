// pml_kernel2.cuh -- convergent PML kernel (variant 2).
//
// Same algorithm and state machine as pml_kernel.cuh (one read per lane, one row fetch per iteration for the
// whole warp), re-written so that the 32 lanes of a warp execute ONE instruction stream per iteration:
// every transition of the state machine is computed with predicates / selects instead of nested branches.
// The round-1 profile of variant 1 (profiles/r01_pml_v1_c2.md) showed 12 of 32 lanes active per issued
// instruction because fast-forward, scan and emit paths were serialised against each other.
//
// Other changes against variant 1:
//   * L2 policy: row fetches carry an evict_last hint, the streamed read bytes / read offsets / PML stores carry
//     evict_first + L1 no-allocate, so the 450 MB of per-batch streaming does not flush the row table out of L2;
//   * output: 4 values (8 bytes) buffered per lane in a 2-register shift register, one 8-byte store per 4 bases;
//     ragged head / tail groups of a read go out as at most three 2-byte stores.
#pragma once
#include "dev_index.cuh"
#include "pml_kernel.cuh"

namespace movi {

__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
// evict_last for a fixed `frac` of the addresses (the hardware picks them by address), evict_first for the others: when the
// row table is larger than what the L2 can hold, protecting the part that fits beats letting all of it thrash
__device__ __forceinline__ unsigned long long policy_keep_fraction(float frac) {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(p) : "f"(frac));
    return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint2 ld_keep(const uint2* p, unsigned long long pol) {
    uint2 v;
    asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
}
// row fetch with a run-time selectable cache recipe (exploration knob "row_ld_mode"; the switch is warp uniform)
__device__ __forceinline__ uint2 ld_row(const uint2* p, unsigned long long pol, uint32_t mode) {
    uint2 v;
    switch (mode) {
        default:
        case 0: asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol)); break;
        case 1: asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.L2::128B.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol)); break;
        case 2: asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.L2::256B.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol)); break;
        case 3: asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.L2::64B.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol)); break;
        case 4: asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol)); break;
        case 5: v = __ldg(p); break;
        case 6: asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::128B.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol)); break;
        case 7: asm volatile("ld.global.nc.L2::128B.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); break;
    }
    return v;
}
__device__ __forceinline__ uint2 ld_stream(const uint2* p, unsigned long long pol) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_stream64(void* p, uint32_t a, uint32_t b, unsigned long long pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" :: "l"(p), "r"(a), "r"(b), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream16(void* p, uint32_t a, unsigned long long pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.u16 [%0], %1, %2;" :: "l"(p), "h"((unsigned short)a), "l"(pol) : "memory");
}

template <typename IdxT, bool HINTS>
__global__ void __launch_bounds__(256)
movi_pml_kernel2(const DevIndex ix, const QueryArgs q) {
    __shared__ uint8_t s_code[256];
    __shared__ uint32_t s_end_thr[4];
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
    const uint2* __restrict__ bases8 = reinterpret_cast<const uint2*>(q.bases);
    const uint2* __restrict__ offs8 = reinterpret_cast<const uint2*>(q.offsets);
    const uint32_t sep = ix.sep;
    const uint32_t n_reads = q.n_reads;

    uint32_t q_next = 0, q_end = 0;
    bool exhausted = false;
    uint32_t st = ST_IDLE;
    IdxT idx = 0;
    uint32_t off = 0, ml = 0, rp = 0, op = 0, rem = 0, clo = 0, chi = 0;
    bool need_chunk = false;
    uint32_t b0 = 0, b1 = 0, pend = 0;

    for (;;) {
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

        // ---- loads: the row fetch (and, when a new 8-byte group of read bytes is needed, the read chunk) go out
        //      first for all walking lanes; lanes that start a read fetch its [start, end) offsets meanwhile.
        const bool is_new = (st == ST_NEW);
        const bool proc = (st >= ST_FF);
        uint2 v = make_uint2(0u, 0u), w = make_uint2(0u, 0u);
        if (proc) {
            if (HINTS) v = ld_row(rows + idx, pol_keep, q.ld_mode); else v = ldg64(rows + idx);
            if (need_chunk) { if (HINTS) w = ld_stream(bases8 + (rp >> 3), pol_stream); else w = ldg64(bases8 + (rp >> 3)); }
        }
        if (is_new) {       // rare per lane: start a read (query_pml :237-240)
            const uint2 a = HINTS ? ld_stream(offs8 + idx, pol_stream) : ldg64(offs8 + idx);
            const uint2 b = HINTS ? ld_stream(offs8 + idx + 1, pol_stream) : ldg64(offs8 + idx + 1);
            const unsigned long long s = (((unsigned long long)a.y << 32) | a.x) - q.origin;
            const unsigned long long e = (((unsigned long long)b.y << 32) | b.x) - q.origin;
            rem = (uint32_t)(e - s);
            rp = (uint32_t)e - 1; op = (uint32_t)s;
            idx = r1; off = ix.n_last - 1; ml = 0; pend = 0;
            st = rem ? ST_FF : ST_IDLE;
            need_chunk = true;
        }

        // ---- one state-machine step for every lane, predicate driven (lanes with !proc change nothing)
        if (proc && need_chunk) { clo = w.x; chi = w.y; }
        need_chunk = need_chunk && !proc;
        const uint32_t y = v.y;
        const uint32_t n = row_n(y);
        const uint32_t c = row_c(y);
        const uint32_t code = s_code[__byte_perm(clo, chi, rp & 7u) & 0xFFu];
        const bool is_ff = proc && (st == ST_FF);
        const bool is_scan = proc && (st >= ST_UP);
        const bool scan_up = (st == ST_UP);
        const bool hit = (c == code);
        const bool ffmove = is_ff && (idx < r1) && (off >= n);            // fast_forward :524-545
        const bool at_base = is_ff && !ffmove;
        const bool base_mis = at_base && !hit && (code != CODE_ILLEGAL);   // Case 2 :297-350
        const bool scan_hit = is_scan && hit;
        const bool scan_move = is_scan && !hit;
        const bool emit = (at_base && !base_mis) || scan_hit;

        // threshold decision (reposition_thresholds :513-601); the '$' row and '%' rows are rare special cases
        const uint32_t ai = code - sep;
        uint32_t thr = row_thr(y, ai, n);
        if (base_mis & ((idx == end_idx) | ((sep != 0) & (c == 0)))) {     // one predicate: skipped by the whole warp
            thr = (idx == end_idx) ? s_end_thr[ai & 3u] : sep_threshold(ix, (unsigned long long)idx, ai);
        }
        const bool down = off >= thr;
        const bool mv_down = (scan_move && !scan_up) || (base_mis && down);
        const bool mv_up = (scan_move && scan_up) || (base_mis && !down);
        const bool bad = (mv_up && idx == 0) || (mv_down && idx == r1);    // reference: reposition returned r

        // matching length (Case 1 :285-296, illegal :278-284, after reposition :313-327)
        ml = (at_base && hit) ? ml + 1 : (emit ? 0u : ml);
        off = ffmove ? off - n : (scan_hit ? (scan_up ? n - 1 : 0u) : off);

        const bool fin = emit && (rem == 1u);
        const bool lf = emit && !fin;
        if (emit) {
            const uint32_t val = ml < 65535u ? ml : 65535u;                 // add_ml clamp
            b0 = __funnelshift_r(b0, b1, 16); b1 = (b1 >> 16) | (val << 16);
            pend += 1;
            if ((op & 3u) == 3u || fin) {
                if (pend == 4u) {
                    if (HINTS) st_stream64(q.out_ml + (op & ~3u), b0, b1, pol_stream);
                    else *reinterpret_cast<uint2*>(q.out_ml + (op & ~3u)) = make_uint2(b0, b1);
                } else {
                    // the last `pend` values sit in the top 16-bit positions: bring them down, then <= 3 short stores
                    const unsigned long long both = (((unsigned long long)b1 << 32) | b0) >> (16u * (4u - pend));
                    uint16_t* t = q.out_ml + (op + 1u - pend);
                    t[0] = (uint16_t)both;
                    if (pend >= 2u) t[1] = (uint16_t)(both >> 16);
                    if (pend == 3u) t[2] = (uint16_t)(both >> 32);
                }
                pend = 0;
            }
        }
        op += emit ? 1u : 0u;
        rem -= emit ? 1u : 0u;
        need_chunk = need_chunk || (lf && (rp & 7u) == 0u);
        rp -= lf ? 1u : 0u;
        off += lf ? row_off(y) : 0u;                                        // LF_move :66
        idx = lf ? row_id<IdxT>(v) : (ffmove || mv_down) ? idx + 1 : (mv_up ? idx - 1 : idx);
        st = fin ? ST_IDLE : lf ? ST_FF : base_mis ? (down ? ST_DOWN : ST_UP) : st;
        if (bad) { atomicOr(q.status, 1u); st = ST_IDLE; }
    }
}

}  // namespace movi
