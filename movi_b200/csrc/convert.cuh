// convert.cuh -- upload-time re-layout kernels: on-disk rows -> R8 device rows.
//
// Each kernel decodes one on-disk layout exactly as the reference's per-MODE getters do and resolves
// MoveStructure::get_id (src/move_structure.cpp:91-286) ONCE per row:
//   blocked-thresholds (MODE 8): id22 + id_blocks[c][idx / block_size] + first_runs[c + 1]   (:95-102)
//   sampled-thresholds (MODE 7): forward scan to the next id checkpoint + backward walk        (:103-219)
// so that the query kernels do one 8-byte row fetch per LF step whatever the on-disk layout was.
#pragma once
#include "dev_index.cuh"

namespace movi {

// ---- MODE 8: 6-byte rows {u16 id; u16 n; u16 offset} (include/move_row.hpp:128-130,139-142;
//      bit fields include/move_row_configs.hpp:76-104)
// ---- MODE 6: the 8-byte rows are already in HBM (copied or adopted); permute their bit fields into R8 in place
__global__ void relayout_regular_kernel(uint2* __restrict__ rows, unsigned long long r, uint32_t sep) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < r; i += stride) rows[i] = r8_from_file(rows[i], sep);
}
// undo of relayout_regular_kernel, in place (an adopted table is handed back to its owner the way it came when the upload fails)
__global__ void relayout_regular_undo_kernel(uint2* __restrict__ rows, unsigned long long r, uint32_t sep) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < r; i += stride) rows[i] = r8_to_file(rows[i], sep);
}
// every row's character must be one of the index's sigma characters (the side tables are indexed by it): counts the offenders
__global__ void check_row_chars_kernel(const uint2* __restrict__ rows, unsigned long long r, unsigned long long end_idx, uint32_t sigma,
                                       unsigned int* __restrict__ status) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < r; i += stride) bad |= (i != end_idx) && (row_c(rows[i].y) >= sigma);
    if (bad) atomicOr(status, 8u);
}
// ---- MODE 3 ("regular", no thresholds): the same, into the 12-bit form of R8
__global__ void relayout_regular_nt_kernel(uint2* __restrict__ rows, unsigned long long r) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < r; i += stride) rows[i] = r8nt_from_file(rows[i]);
}
__global__ void relayout_nt_to_file_kernel(const uint2* __restrict__ rows, uint2* __restrict__ out, unsigned long long r) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < r; i += stride) out[i] = r8nt_to_file(rows[i]);
}
// the inverse, into a separate buffer (generators / tests that hand a device table to the CPU oracle)
__global__ void relayout_to_file_kernel(const uint2* __restrict__ rows, uint2* __restrict__ out, unsigned long long r, uint32_t sep) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < r; i += stride) out[i] = r8_to_file(rows[i], sep);
}

__global__ void convert_blocked_kernel(const uint16_t* __restrict__ raw, uint2* __restrict__ out,
                                       unsigned long long r, unsigned long long end_idx, uint32_t sep, uint32_t nothr,
                                       const uint32_t* __restrict__ id_blocks, unsigned long long id_blocks_size,
                                       unsigned long long block_size, const unsigned long long* __restrict__ first_runs) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < r; i += stride) {
        const uint32_t f_id = raw[3 * i], f_n = raw[3 * i + 1], f_off = raw[3 * i + 2];
        const uint32_t n = f_n & 0x3FFu;
        const uint32_t off = f_off & 0x3FFu;
        const uint32_t c = (f_off >> 10) & 7u;
        // MODE 8: thresholds #0..#2 = offset bits 13..15.  MODE 2 ("blocked", no thresholds): offset bits 14..15 are two more
        // id bits (include/move_row.hpp:267-286, move_row_configs.hpp:54-73)
        const uint32_t thr = nothr ? 0u : (f_off >> 13) & 7u;
        unsigned long long id = (unsigned long long)f_id | ((unsigned long long)((f_n >> 10) & 0x3Fu) << 16);
        if (nothr) id |= (unsigned long long)((f_off >> 14) & 3u) << 22;
        if (i != end_idx) id += (unsigned long long)id_blocks[c * id_blocks_size + i / block_size] + first_runs[c + 1];
        out[i] = make_r8(id, n, off, c, thr, sep);
    }
}

// ---- MODE 7 -> R8: one thread per row runs sampled_get_id (dev_index.cuh) on the packed table
__global__ void convert_sampled_kernel(const uint32_t* __restrict__ raw_words, uint2* __restrict__ out,
                                       unsigned long long r, unsigned long long end_idx, uint32_t sep, uint32_t nothr,
                                       const uint8_t* __restrict__ tally, unsigned long long tally_len, uint32_t cp,
                                       unsigned int* __restrict__ status) {
    const Rows3 R{raw_words, end_idx, nothr};
    unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; idx < r; idx += stride) {
        const uint32_t x = R.raw(idx);
        const uint32_t ci = R.dc(x);
        bool bad = false;
        const unsigned long long id = sampled_get_id(R, tally, tally_len, cp, r, end_idx, idx, ci, &bad);
        if (bad) atomicOr(status, 2u);
        out[idx] = make_r8(id, R.dn(x), R.doff(x), ci, nothr ? 0u : Rows3::thr_of(x), sep);
    }
}

// ---- MODE 7 -> T4 (dev_index.cuh): the 3-byte row + the walk-back distance of its destination id from the checkpoint entry
__global__ void convert_sampled_t4_kernel(const uint32_t* __restrict__ raw_words, uint32_t* __restrict__ out, unsigned long long r,
                                          unsigned long long end_idx, const uint8_t* __restrict__ tally, unsigned long long tally_len, uint32_t cp,
                                          unsigned int* __restrict__ status) {
    const Rows3 R{raw_words, end_idx, 0u};
    unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; idx < r; idx += stride) {
        const uint32_t x = R.raw(idx);
        const uint32_t ci = Rows3::c_of(x);
        bool bad = false;
        uint32_t d = 0;
        if (idx != end_idx) {
            const unsigned long long id = sampled_get_id(R, tally, tally_len, cp, r, end_idx, idx, ci, &bad);
            const unsigned long long from = tally_get(tally, tally_len, ci, t4_tally_slot(idx, r, tally_len, cp));
            d = (from >= id && from - id < T4_FAR) ? (uint32_t)(from - id) : T4_FAR;
        }
        if (bad) atomicOr(status, 2u);
        out[idx] = Rows4::pack(Rows3::n_of(x), Rows3::off_of(x), ci, Rows3::thr_of(x), d);
    }
}

// ---- sampled prefix sums of run lengths (for MoveInterval::count in O(1): include/move_intervals.hpp:47-58)
// pass 1: group sums (one u32 per 2^PREFIX_SHIFT rows) and per-block totals
constexpr int SCAN_BLOCK = 1024;
__global__ void prefix_group_sums_kernel(const DevIndex ix, unsigned long long r,
                                         unsigned long long n_groups, unsigned long long* __restrict__ group_sum,
                                         unsigned long long* __restrict__ block_sum) {
    __shared__ unsigned long long s[SCAN_BLOCK / 32];
    const unsigned long long g = blockIdx.x * (unsigned long long)SCAN_BLOCK + threadIdx.x;
    unsigned long long sum = 0;
    if (g < n_groups) {
        const unsigned long long b = g << PREFIX_SHIFT;
        for (int k = 0; k < (1 << PREFIX_SHIFT); k++) if (b + k < r) sum += row_len_at(ix, b + k);
        group_sum[g] = sum;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned long long v = threadIdx.x < SCAN_BLOCK / 32 ? s[threadIdx.x] : 0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) block_sum[blockIdx.x] = v;
    }
}
// pass 2: exclusive scan of the block totals by a single block (n_blocks is r / 16384: small)
__global__ void prefix_scan_blocks_kernel(unsigned long long* __restrict__ block_sum, unsigned long long n_blocks) {
    __shared__ unsigned long long s[SCAN_BLOCK];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (unsigned long long base = 0; base < n_blocks; base += SCAN_BLOCK) {
        const unsigned long long i = base + threadIdx.x;
        const unsigned long long v = i < n_blocks ? block_sum[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
            unsigned long long t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n_blocks) block_sum[i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) carry += s[threadIdx.x];
        __syncthreads();
    }
}
// pass 3: exclusive scan inside each block + block offset, in place (group_sum -> prefix)
__global__ void prefix_finish_kernel(unsigned long long* __restrict__ group_sum, unsigned long long n_groups,
                                     const unsigned long long* __restrict__ block_off) {
    __shared__ unsigned long long s[SCAN_BLOCK];
    const unsigned long long g = blockIdx.x * (unsigned long long)SCAN_BLOCK + threadIdx.x;
    const unsigned long long v = g < n_groups ? group_sum[g] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
        unsigned long long t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    if (g < n_groups) group_sum[g] = block_off[blockIdx.x] + s[threadIdx.x] - v;
}

}  // namespace movi
