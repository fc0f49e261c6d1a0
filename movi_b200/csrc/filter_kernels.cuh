// filter_kernels.cuh -- device-side emission of the `--filter` output: the reads Classifier::classify keeps (FOUND, or NOT_PRESENT
// with --invert) echoed as ">" id "\n" sequence "\n" (output_read, reference src/utils.cpp:291-294 via src/movi.cpp:41-47), assembled
// in HBM from the parsed chunk, so that only the kept text crosses PCIe and the host formats nothing.
//
//   filter_sizes_kernel   bytes of every read's echo (0 for a read that is dropped) from its two bin counters
//   scan_u64_kernel       exclusive prefix sums of those sizes (one block; a chunk holds at most a few million reads)
//   filter_emit_kernel    one warp per kept read copies id and sequence to their place
// The per-read pieces are plain functions, compiled for the host as well: tests/native/ext_emulate.cu runs them against
// movi_host_classify_bins (the host restatement of the reference's writer) in the CPU test run.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace movi {

// found = above / (above + below + 0.0) > 0.50 (src/classifier.cpp:126): for counters below 2^32 the rounded quotient cannot reach
// 0.5 from above unless above > below, and 0 / 0 (an empty read) is NaN, which compares false -- so: above > below
__host__ __device__ __forceinline__ bool filter_keeps(uint32_t above, uint32_t below, int invert) {
    const bool found = above > below;
    return invert ? !found : found;
}
__host__ __device__ __forceinline__ unsigned long long filter_echo_bytes(unsigned long long id_len, unsigned long long len) {
    return id_len + len + 3ull;                                       // '>' id '\n' sequence '\n'
}
// lane `lane` of `stride` lanes writes its share of read i's echo at dst
__host__ __device__ __forceinline__ void filter_emit_read(uint8_t* __restrict__ dst, const uint8_t* __restrict__ id, unsigned long long id_len,
                                                          const uint8_t* __restrict__ seq, unsigned long long len, unsigned lane, unsigned stride) {
    if (lane == 0) { dst[0] = '>'; dst[1 + id_len] = '\n'; dst[2 + id_len + len] = '\n'; }
    for (unsigned long long k = lane; k < id_len; k += stride) dst[1 + k] = id[k];
    uint8_t* body = dst + 2 + id_len;
    for (unsigned long long k = lane; k < len; k += stride) body[k] = seq[k];
}

__global__ void filter_sizes_kernel(const uint32_t* __restrict__ above, const uint32_t* __restrict__ below, int invert,
                                    const unsigned long long* __restrict__ offsets, const unsigned long long* __restrict__ id_offsets,
                                    unsigned long long n_reads, unsigned long long* __restrict__ sizes) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < n_reads; i += stride)
        sizes[i] = filter_keeps(above[i], below[i], invert) ? filter_echo_bytes(id_offsets[i + 1] - id_offsets[i], offsets[i + 1] - offsets[i]) : 0ull;
}

// v[0 .. n) -> exclusive prefix sums in place, v[n] = the total.  One block of 1024 threads, every thread a contiguous share.
__global__ void __launch_bounds__(1024) scan_u64_kernel(unsigned long long* __restrict__ v, unsigned long long n) {
    __shared__ unsigned long long part[1024];
    const unsigned t = threadIdx.x;
    const unsigned long long per = (n + 1023ull) / 1024ull;
    const unsigned long long b = min(n, per * t), e = min(n, b + per);
    unsigned long long s = 0;
    for (unsigned long long i = b; i < e; i++) s += v[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (unsigned k = 0; k < 1024; k++) { const unsigned long long x = part[k]; part[k] = run; run += x; }
        v[n] = run;
    }
    __syncthreads();
    unsigned long long run = part[t];
    for (unsigned long long i = b; i < e; i++) { const unsigned long long x = v[i]; v[i] = run; run += x; }
}

__global__ void filter_emit_kernel(const unsigned long long* __restrict__ pos, unsigned long long n_reads,
                                   const uint8_t* __restrict__ bases, const unsigned long long* __restrict__ offsets,
                                   const uint8_t* __restrict__ ids, const unsigned long long* __restrict__ id_offsets,
                                   uint8_t* __restrict__ text) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long w = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
    const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (; w < n_reads; w += n_warps) {
        if (pos[w + 1] == pos[w]) continue;                           // dropped (an echo is never empty: it has at least 3 bytes)
        filter_emit_read(text + pos[w], ids + id_offsets[w], id_offsets[w + 1] - id_offsets[w], bases + offsets[w], offsets[w + 1] - offsets[w], lane, 32u);
    }
}

}  // namespace movi
