// ext_kernel.cuh -- the PML walk with a side task at every position it settles on (Movi Color):
//
//   EXT_COLOR   `movi query --pml --multi-classify`: after the LF step of every base whose matching length reaches --min-len,
//               the documents of the colour of the row the walk now stands on are counted per read, best and second best
//               following in the order the documents are met (MoveStructure::query_pml, src/move_structure_query.cpp:251-260,
//               373-437; the flat colour table of src/move_structure_io.cpp:513-548, 587-607).
//   EXT_PVALUE  the same with `--pvalue-scoring`: min(match_len - log(len(doc)) / log 4, 1) added to a double score (:419-433).
//   EXT_SA      `--sa-entries`: MoveStructure::get_SA_entries (src/move_structure.cpp:35-48) for the position of every base.
//
// One read per lane and one row fetch per step, like movi_pml_kernel; but every lane is independent here (work is claimed with
// one atomic per read, no warp-level queue, no shared-memory staging): ExtLane::step is plain C++ on plain pointers, compiled
// for the host as well -- tests/native/ext_emulate.cu runs the very same source against the oracle and the reference's outputs
// without a GPU, and the -m gpu tests then only have to confirm the launch.  The walk itself follows pml_kernel.cuh state for
// state (ST_FF / ST_UP / ST_DOWN, thresholds, the '$' and '%' rows), on the expanded 8-byte rows of a threshold index.
//
// Differences from the plain PML walk that the two side tasks need:
//   * the reference runs LF_move after EVERY base, the last one included (query.cpp:358-361), and counts the colour of the row it
//     lands on: a read ends one settled position after its last base ("tail");
//   * get_SA_entries walks LF steps from the base's position P until it meets a sampled BWT offset, d steps away, and returns
//     sampled + d.  Its first step is the walk's own LF step, so the side walk starts from the next settled position with d = 1;
//     and while the walk is not repositioned, the entry of the next base is the same sample at distance d - 1: entry - 1, for
//     free, until d reaches 0 (exactly what the reference computes, which applies no modulo either).
#pragma once
#include <cstring>
#include "../../include/movi_gpu.h"
#include "kernel_common.cuh"

namespace movi {

constexpr int EXT_COLOR = 1, EXT_PVALUE = 2, EXT_SA = 4;
constexpr uint32_t DOC_NONE = 0xFFFFu;        // std::numeric_limits<uint16_t>::max() (query.cpp:262-263)
constexpr uint32_t ST_SA = 7;                 // the lane is on its get_SA_entries side walk

struct ExtArgs {
    const uint8_t* bases;                     // read i = bases[offsets[i] - origin .. offsets[i + 1] - origin)
    const unsigned long long* offsets;
    unsigned long long origin;
    uint32_t n_reads;
    uint16_t* out_ml;                         // matching lengths at the same positions (may be null)
    unsigned long long* counter;              // next read to claim (zeroed before launch)
    unsigned int* status;                     // sticky error bits (1: a scan left the table; 4: 2^32 bases or more in one call)
    // EXT_COLOR
    const uint8_t* color_inds;                // r x MoveTally {u32 right; u8 left}: offset of the row's colour in `flat`
    const uint16_t* flat;                     // flat colours: size, documents, size, documents, ...
    unsigned long long flat_size;
    uint32_t num_species, min_len;
    const double* lk;                         // EXT_PVALUE: log_lens[doc] / log4
    void* cnt;                                // per-lane counters (u32) or scores (double): document d of lane l at [d * cnt_stride + l]
    unsigned long long cnt_stride;
    MoviGpuColorResult* out_color;            // per read
    // EXT_SA
    const unsigned long long* all_p;          // BWT offset of every run head (r)
    const unsigned long long* sampled;        // sampled_SA_entries
    unsigned long long rate;
    unsigned long long* out_sa;               // per base, same positions as out_ml
};

template <typename IdxT, bool SEP, int EXT>
struct ExtLane {
    uint32_t st = ST_IDLE;
    IdxT idx = 0;
    uint32_t off = 0, ml = 0, rp = 0, op = 0, rem = 0, ri = 0;
    // colours
    bool cpend = false;                       // the position the LF step in flight leads to is to be counted
    uint32_t cml = 0;                         // matching length of the base that LF step belongs to
    uint32_t best = DOC_NONE, second = DOC_NONE, sum_ml = 0;
    // SA entries
    bool sa_valid = false, sa_search = false;
    unsigned long long sa_val = 0, sa_dist = 0, sa_d = 0;
    uint32_t sa_slot = 0;
    IdxT idx2 = 0;
    uint32_t off2 = 0;

    __host__ __device__ void raise(const ExtArgs& a, unsigned int bits) {
#ifdef __CUDA_ARCH__
        atomicOr(a.status, bits);
#else
        *a.status |= bits;
#endif
        st = ST_IDLE;
    }

    __host__ __device__ void write_result(const ExtArgs& a, unsigned long long slot) {
        if (!(EXT & EXT_COLOR)) return;
        MoviGpuColorResult res;
        res.best_doc = best; res.second_best_doc = second; res.sum_matching_lengths = sum_ml; res.reserved = 0;
        res.best_raw = 0; res.second_raw = 0;
        if (EXT & EXT_PVALUE) {
            const double* sc = static_cast<const double*>(a.cnt);
            double b = best != DOC_NONE ? sc[best * a.cnt_stride + slot] : 0., s = second != DOC_NONE ? sc[second * a.cnt_stride + slot] : 0.;
            static_assert(sizeof(double) == sizeof(res.best_raw), "scores travel as their bits");
            memcpy(&res.best_raw, &b, 8); memcpy(&res.second_raw, &s, 8);
        } else {
            const uint32_t* cn = static_cast<const uint32_t*>(a.cnt);
            if (best != DOC_NONE) res.best_raw = cn[best * a.cnt_stride + slot];
            if (second != DOC_NONE) res.second_raw = cn[second * a.cnt_stride + slot];
        }
        a.out_color[ri] = res;
    }

    // claim read `r`: query_pml's start state (query.cpp:236-263)
    __host__ __device__ void start(const DevIndex& ix, const ExtArgs& a, uint32_t r, unsigned long long slot) {
        ri = r;
        const unsigned long long s = a.offsets[r] - a.origin, e = a.offsets[r + 1] - a.origin;
        rem = (uint32_t)(e - s);
        if (e >> 32) { raise(a, 4u); rem = 0; }
        rp = (uint32_t)e - 1; op = (uint32_t)s;
        idx = (IdxT)(ix.r - 1); off = ix.n_last - 1; ml = 0;
        cpend = false; best = DOC_NONE; second = DOC_NONE; sum_ml = 0;
        sa_valid = false; sa_search = false;
        if (EXT & EXT_COLOR) {
            if (EXT & EXT_PVALUE) { double* sc = static_cast<double*>(a.cnt); for (uint32_t d = 0; d < a.num_species; d++) sc[d * a.cnt_stride + slot] = 0.; }
            else { uint32_t* cn = static_cast<uint32_t*>(a.cnt); for (uint32_t d = 0; d < a.num_species; d++) cn[d * a.cnt_stride + slot] = 0u; }
        }
        if (rem) st = ST_FF;
        else { write_result(a, slot); st = ST_IDLE; }        // an empty read: "id,0,0" (best_doc stays at its start value)
    }

    // the documents of the colour of row idx (query.cpp:373-437)
    __host__ __device__ void count_colors(const ExtArgs& a, unsigned long long slot) {
        const uint8_t* e = a.color_inds + 5ull * (unsigned long long)idx;         // MoveTally::get, include/move_row.hpp:28-38
        const unsigned long long color_id = (unsigned long long)e[0] | ((unsigned long long)e[1] << 8) | ((unsigned long long)e[2] << 16) |
                                            ((unsigned long long)e[3] << 24) | ((unsigned long long)e[4] << 32);
        if (color_id >= a.flat_size) return;                                      // :395 (doc sets dropped by compression)
        const uint32_t set_size = a.flat[color_id];
        for (uint32_t j = 0; j < set_size; j++) {
            const uint32_t doc = a.flat[color_id + 1 + j];
            if (EXT & EXT_PVALUE) {
                double* sc = static_cast<double*>(a.cnt);
                const double val = (double)cml - a.lk[doc];                        // :421
                if (val >= 0) {
                    const double now = sc[doc * a.cnt_stride + slot] + (val < 1. ? val : 1.);
                    sc[doc * a.cnt_stride + slot] = now;
                    if (doc != best) {
                        if (best == DOC_NONE || now > sc[best * a.cnt_stride + slot]) { second = best; best = doc; }
                        else if (second == DOC_NONE || now > sc[second * a.cnt_stride + slot]) second = doc;
                    }
                }
            } else {
                uint32_t* cn = static_cast<uint32_t*>(a.cnt);
                const uint32_t now = cn[doc * a.cnt_stride + slot] + 1u;            // :407-416
                cn[doc * a.cnt_stride + slot] = now;
                if (doc != best) {
                    if (best == DOC_NONE || now > cn[best * a.cnt_stride + slot]) { second = best; best = doc; }
                    else if (second == DOC_NONE || now > cn[second * a.cnt_stride + slot]) second = doc;
                }
            }
        }
    }

    // the base at rp is decided at (idx, off), v = that row: add_ml, the side tasks, LF_move (query.cpp:351-361)
    __host__ __device__ void emit(const ExtArgs& a, uint2 v, bool repositioned) {
        sum_ml += ml;                                                             // uint32 arithmetic like the reference's (:265, :351)
        if (a.out_ml) a.out_ml[op] = (uint16_t)(ml < 65535u ? ml : 65535u);       // add_ml clamp
        if (EXT & EXT_SA) {
            if (sa_valid && !repositioned && sa_dist > 0) { sa_val -= 1; sa_dist -= 1; a.out_sa[op] = sa_val; }
            else {
                const unsigned long long abs_off = a.all_p[idx] + off;            // get_SA_entries, move_structure.cpp:36-47
                if (abs_off % a.rate == 0) { sa_val = a.sampled[abs_off / a.rate]; sa_dist = 0; sa_valid = true; a.out_sa[op] = sa_val; }
                else { sa_search = true; sa_valid = false; sa_slot = op; sa_d = 1; }
            }
        }
        if (EXT & EXT_COLOR) { cpend = ml >= a.min_len; cml = ml; }
        op += 1; rem -= 1;
        if (rem) rp -= 1;
        off += row_off(v.y);                                                      // LF_move, move_structure.cpp:59-87
        idx = row_id<IdxT>(v);
        st = ST_FF;
    }

    // one row fetch
    __host__ __device__ void step(const DevIndex& ix, const ExtArgs& a, unsigned long long slot) {
        const IdxT r1 = (IdxT)(ix.r - 1);
        const IdxT end_idx = (IdxT)ix.end_idx;
        constexpr uint32_t sep = SEP ? 1u : 0u;
        if ((EXT & EXT_SA) && st == ST_SA) {
            const uint2 w = ix.rows[idx2];
            const uint32_t n2 = row_n(w.y);
            if (idx2 < r1 && off2 >= n2) { off2 -= n2; idx2 += 1; return; }      // fast_forward of the side walk
            const unsigned long long abs_off = a.all_p[idx2] + off2;
            if (abs_off % a.rate == 0) {
                sa_val = a.sampled[abs_off / a.rate] + sa_d;
                a.out_sa[sa_slot] = sa_val;
                sa_dist = sa_d; sa_valid = true;
                st = ST_FF;                                                       // back to the walk, which stands where it was
                return;
            }
            sa_d += 1;
            off2 += row_off(w.y); idx2 = row_id<IdxT>(w);
            return;
        }
        const uint2 v = ix.rows[idx];
        const uint32_t y = v.y, n = row_n(y), c = row_c(y);
        if (st == ST_FF) {
            if (idx < r1 && off >= n) { off -= n; idx += 1; return; }            // fast_forward, move_structure.cpp:524-545
            // the walk stands on (idx, off): where the reference is right after LF_move
            if ((EXT & EXT_COLOR) && cpend) { count_colors(a, slot); cpend = false; }
            if ((EXT & EXT_SA) && sa_search) { sa_search = false; idx2 = idx; off2 = off; st = ST_SA; return; }
            if (rem == 0) { write_result(a, slot); st = ST_IDLE; return; }      // that was the LF step of the read's last base
            const uint32_t code = ix.code_of[a.bases[rp]];
            if (code == CODE_ILLEGAL) { ml = 0; emit(a, v, false); return; }     // query.cpp:278-285
            if (c == code) { ml += 1; emit(a, v, false); return; }               // Case 1
            // Case 2: reposition_thresholds (query.cpp:513-601)
            const uint32_t ai = code - sep;
            uint32_t thr = row_thr(y, ai, n);
            if (idx == end_idx || (SEP && c == 0)) thr = idx == end_idx ? ix.end_thr[ai & 3u] : sep_threshold(ix, (unsigned long long)idx, ai);
            const bool down = off >= thr;
            if (down ? idx == r1 : idx == 0) { raise(a, 1u); return; }
            st = down ? ST_DOWN : ST_UP;
            idx = down ? idx + 1 : idx - 1;
            return;
        }
        // ST_UP / ST_DOWN: reposition_up / reposition_down (query.cpp:188-232)
        const uint32_t code = ix.code_of[a.bases[rp]];
        const bool up = st == ST_UP;
        if (c == code) { off = up ? n - 1 : 0u; ml = 0; emit(a, v, true); return; }
        if (up ? idx == 0 : idx == r1) { raise(a, 1u); return; }
        idx = up ? idx - 1 : idx + 1;
    }
};

template <typename IdxT, bool SEP, int EXT>
__global__ void __launch_bounds__(QUERY_THREADS)
movi_ext_kernel(const DevIndex ix, const ExtArgs a) {
    ExtLane<IdxT, SEP, EXT> lane;
    const unsigned long long slot = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    for (;;) {
        if (lane.st == ST_IDLE) {
            const unsigned long long r = atomicAdd(a.counter, 1ull);
            if (r >= a.n_reads) break;
            lane.start(ix, a, (uint32_t)r, slot);
            continue;
        }
        lane.step(ix, a, slot);
    }
}

}  // namespace movi
