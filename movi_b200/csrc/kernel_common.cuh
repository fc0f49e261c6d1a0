// kernel_common.cuh -- what the query kernels share: the launch arguments, the walker states, L2 cache policies and the
// shared-memory / cp.async helpers of the sector-granular streaming (read bytes in, matching lengths out).
#pragma once
#include "dev_index.cuh"

namespace movi {

struct QueryArgs {
    const uint8_t* bases;                 // 32-byte aligned; read i = bases[offsets[i]-origin .. offsets[i+1]-origin)
    const unsigned long long* offsets;    // n_reads + 1
    unsigned long long origin;            // subtracted from every offset
    uint32_t n_reads;
    uint16_t* out_ml;                     // PML / ZML values at the same (offset - origin) positions
    unsigned long long* out_count;        // count query
    int32_t* out_pos;                     // count query
    unsigned long long* counter;          // work-queue head (zeroed before launch)
    unsigned int* status;                 // sticky error bits
    uint32_t flags;                       // MOVI_GPU_FLAG_* of the call
};

constexpr uint32_t ST_IDLE = 0, ST_FF = 2, ST_UP = 3, ST_DOWN = 4;
constexpr uint32_t QUEUE_CHUNK = 32;      // reads taken from the global counter per warp refill (small: short tail)
constexpr unsigned FULL = 0xffffffffu;
constexpr int QUERY_THREADS = 256;
constexpr int IDB_SMEM = 1024;            // id_blocks entries kept in shared memory (covers r <= 2.7e8 rows at block_size 2^20)

// L2 policies: row fetches evict_last, the streamed data (read bytes in, lengths out) evict_first + L1 no-allocate, so that the
// hundreds of MB that stream through per batch displace as little of the row table as possible (without any hint on the rows
// the c2 kernel takes 2.69 ms instead of 1.97; the hints on the streams are worth another 2.5 %: profiles/r01_explore.jsonl)
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint2 ld_keep(const uint2* p, unsigned long long pol) {
    uint2 v;
    asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint4 ld_keep128(const uint4* p, unsigned long long pol) {
    uint4 v;
    asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
// one aligned 32-byte sector (4 rows) in ONE request: a 256-bit load (sm_100+).  L2::64B: on a miss the L2 of this part brings in
// the whole 128-byte line (4 sectors) by default, whatever the width of the load -- 4.0 DRAM sectors per random load measured on a
// 36 GB table, 2.0 with this qualifier (profiles/r02_load_flavours_ncu.csv; cudaLimitMaxL2FetchGranularity changes nothing).  This
// fetch serves tables far beyond the L2, where the other sectors of the line are never used before they are evicted.
__device__ __forceinline__ void ld_keep256(const void* p, unsigned long long pol, uint4& lo, uint4& hi) {
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(pol));
    lo = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
    hi = make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)d, (uint32_t)(d >> 32));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ uint4 lds_128(uint32_t a) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void cp_async16_stream(uint32_t s, const void* gmem, unsigned long long pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(s), "l"(gmem), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream128(void* p, uint4 v, unsigned long long pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

}  // namespace movi
