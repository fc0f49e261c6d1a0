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
