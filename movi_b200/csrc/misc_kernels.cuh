// misc_kernels.cuh -- device-side classification from PMLs and the random-access roofline probe.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace movi {

// Classifier::classify (src/classifier.cpp:99-143): bins of bin_width over the matching lengths (a short
// tail is merged into the previous bin), bin maximum >= max_value_thr counts as "above".  One warp per read.
__global__ void classify_kernel(const uint16_t* __restrict__ ml, const unsigned long long* __restrict__ offsets, unsigned long long origin,
                                unsigned long long n_reads, uint32_t bin_width, uint32_t thr,
                                uint32_t* __restrict__ above_out, uint32_t* __restrict__ below_out,
                                unsigned long long* __restrict__ sum_out) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long w = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
    const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (; w < n_reads; w += n_warps) {
        const unsigned long long s = offsets[w] - origin, e = offsets[w + 1] - origin;
        const unsigned long long len = e - s;
        uint32_t above = 0, below = 0;
        unsigned long long sum = 0;
        unsigned long long start = 0;
        while (start < len) {
            unsigned long long end = (start + bin_width < len) ? start + bin_width : len;
            if (len - end < bin_width) end = len;
            uint32_t mx = 0;
            for (unsigned long long i = start + lane; i < end; i += 32) mx = max(mx, (uint32_t)ml[s + i]);
            for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (mx >= thr) above++; else below++;
            sum += mx;
            start = end;
        }
        if (lane == 0) { above_out[w] = above; below_out[w] = below; sum_out[w] = sum; }
    }
}

// Device-side .bpf emission (output_base_stats, binary branch: reference src/utils.cpp:202-246): the records
//     { u16 id_len; char id[id_len]; u64 L; u16 ml[L] }
// of reads [0, n_reads) back to back, i.e. exactly the bytes the reference appends to <prefix>.pml.bpf after the 12-byte BPFHeader.
// No scan is needed for the record positions: id_offsets and read offsets are prefix sums already, so record i starts at byte
// 10 * i + (id_offsets[i] - id_offsets[0]) + 2 * (offsets[i] - offsets[0]).  One warp per read.  The payload lands at an
// arbitrary byte address (odd when id_len is odd): interior 4-byte words of the destination are assembled from the two aligned
// source words that hold their bytes (funnel shift), the edges of a record go byte by byte.
__global__ void bpf_assemble_kernel(const uint16_t* __restrict__ ml, const unsigned long long* __restrict__ offsets, unsigned long long origin,
                                    unsigned long long n_reads, const uint8_t* __restrict__ ids, const unsigned long long* __restrict__ id_offsets,
                                    uint8_t* __restrict__ image) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long w = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
    const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long off0 = offsets[0], id0 = id_offsets[0];
    for (; w < n_reads; w += n_warps) {
        const unsigned long long s = offsets[w], L = offsets[w + 1] - s;
        const unsigned long long ia = id_offsets[w];
        const uint32_t idl = (uint32_t)(id_offsets[w + 1] - ia);
        uint8_t* rec = image + 10ull * w + (ia - id0) + 2ull * (s - off0);
        for (uint32_t b = lane; b < 10u + idl; b += 32u) {                      // u16 id_len | id | u64 L
            uint8_t v;
            if (b < 2u) v = (uint8_t)(idl >> (8u * b));
            else if (b < 2u + idl) v = ids[ia - id0 + (b - 2u)];
            else v = (uint8_t)(L >> (8u * (b - 2u - idl)));
            rec[b] = v;
        }
        uint8_t* pay = rec + 10u + idl;                                         // 2 * L bytes of matching lengths
        const uint8_t* src = reinterpret_cast<const uint8_t*>(ml + (s - origin));
        const unsigned long long nbytes = 2ull * L;
        const unsigned long long head = min((unsigned long long)((4u - ((uintptr_t)pay & 3u)) & 3u), nbytes);   // bytes up to the first aligned word
        const unsigned long long words = (nbytes - head) >> 2;
        const unsigned long long tail0 = head + (words << 2);
        if (lane < head) pay[lane] = src[lane];
        if (lane < nbytes - tail0) pay[tail0 + lane] = src[tail0 + lane];
        uint32_t* dst = reinterpret_cast<uint32_t*>(pay + head);
        const uint8_t* sp = src + head;                                         // source of the first aligned word; sp + 4k for word k
        const uint32_t sh = 8u * (uint32_t)((uintptr_t)sp & 3u);
        const uint32_t* sw = reinterpret_cast<const uint32_t*>((uintptr_t)sp & ~(uintptr_t)3);
        for (unsigned long long k = lane; k < words; k += 32u) {
            const uint32_t lo = sw[k], hi = sh ? sw[k + 1] : 0u;                // (hi stays inside the ml buffer: it is padded by 64 bytes)
            dst[k] = __funnelshift_r(lo, hi, sh);
        }
    }
}

// u16 -> u8 re-pack of a slice of matching lengths (movi_gpu_query_ml8: every read of the batch is <= 255 bases, so no
// value can exceed 255; clamped anyway).  16 values per thread: two 16-byte loads, one 16-byte store.
__global__ void narrow_ml_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, unsigned long long n_groups) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < n_groups; i += stride) {
        const uint4 a = in[2 * i], b = in[2 * i + 1];
        uint4 o;
        o.x = __byte_perm(__vminu2(a.x, 0x00FF00FFu), __vminu2(a.y, 0x00FF00FFu), 0x6420);
        o.y = __byte_perm(__vminu2(a.z, 0x00FF00FFu), __vminu2(a.w, 0x00FF00FFu), 0x6420);
        o.z = __byte_perm(__vminu2(b.x, 0x00FF00FFu), __vminu2(b.y, 0x00FF00FFu), 0x6420);
        o.w = __byte_perm(__vminu2(b.z, 0x00FF00FFu), __vminu2(b.w, 0x00FF00FFu), 0x6420);
        out[i] = o;
    }
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

__global__ void probe_fill_kernel(uint2* __restrict__ t, unsigned long long n) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) { unsigned long long h = mix64(i + 1); t[i] = make_uint2((uint32_t)h, (uint32_t)(h >> 32)); }
}

// Random-access probe: every thread walks `steps` pseudo-random rows of the table.  DEP = true makes each
// address depend on the previous load (the shape of an LF walk); DEP = false issues independent loads
// (the practical ceiling of the memory system for 32-byte-sector gathers).
template <int BYTES, bool DEP>
__global__ void gather_probe_kernel(const uint2* __restrict__ table, unsigned long long n_rows, uint32_t steps,
                                    unsigned long long* __restrict__ sink) {
    unsigned long long x = mix64(blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 0x9E3779B97F4A7C15ULL);
    unsigned long long acc = 0;
    for (uint32_t s = 0; s < steps; s++) {
        unsigned long long idx = __umul64hi(x, n_rows);
        unsigned long long got;
        if (BYTES == 8) {
            const uint2 v = __ldg(table + idx);
            got = ((unsigned long long)v.y << 32) | v.x;
        } else if (BYTES == 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(table + (idx & ~1ull)));
            got = (((unsigned long long)v.y << 32) | v.x) ^ (((unsigned long long)v.w << 32) | v.z);
        } else {
            const uint4* p = reinterpret_cast<const uint4*>(table + (idx & ~3ull));
            unsigned long long a0, a1, a2, a3;
            asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a0), "=l"(a1), "=l"(a2), "=l"(a3) : "l"(p));
            got = a0 ^ a1 ^ a2 ^ a3;
        }
        acc += got;
        x = DEP ? mix64(x ^ got) : x * 6364136223846793005ULL + 1442695040888963407ULL;
    }
    if (acc == 0x1234567ull) *sink = acc;
}

}  // namespace movi
