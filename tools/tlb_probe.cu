// tlb_probe.cu -- what bounds random row fetches on tables far beyond the L2?  (measurement tool, not product)
//
// Round 1 found a flat ceiling of ~3.7e10 random loads/s from 1 GB to 40 GB whatever the load width or occupancy
// (profiles/r01_probe.jsonl), against 1.24e11 on a 124 MB table.  Two candidate causes predict different outcomes:
//   * address translation (every access misses the SM-local TLB, reach 128 x 2 MB): a SECOND access to the same 2 MB page is
//     nearly free, and an SM that confines itself to a slice <= 256 MB of the table runs at the DRAM rate;
//   * DRAM row activations: a second access is free only inside the same DRAM page (~1 KB), slicing by SM changes nothing.
// Modes (one JSON line each):
//   single            random 8-byte loads over the whole table
//   pair:<mask>       loads at A and A ^ mask (bytes)          -> pairs/s
//   slice:<MB>:<g>    every SM group g (1 = per SM, 2 = per TPC pair, ...) confines itself to its own slice of <MB> MB
//   bslice:<MB>       slices chosen by blockIdx instead of %smid (control: blocks of one SM hit different slices)
//   atomic:<n>        atomicAdd with return on n counters chosen at random (queue-tail contention)
//   queue:<n>         claim a slot in one of n rings (atomicAdd) and write a 32-byte record there (walk hand-over cost)
// usage: tlb_probe <table GB> <mode> [<mode> ...]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <cuda_runtime.h>

typedef unsigned long long u64;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); exit(2); } } while (0)

__device__ __forceinline__ u64 mix64(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }

__global__ void fill(u64* t, u64 n) {
    u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    for (; i < n; i += (u64)gridDim.x * blockDim.x) t[i] = mix64(i + 1);
}

// MODE 0: whole table; 1: pair with xor mask; 2: slice by smid / group; 3: slice by blockIdx
template <int MODE>
__global__ void gather(const u64* __restrict__ t, u64 n_words, u64 slice_words, unsigned group, u64 mask_words, unsigned steps, u64* sink) {
    u64 x = mix64(blockIdx.x * (u64)blockDim.x + threadIdx.x + 0x9E3779B97F4A7C15ULL);
    u64 base = 0, span = n_words;
    if (MODE == 2) { base = (u64)(smid() / group) * slice_words; span = slice_words; }
    if (MODE == 3) { base = (u64)(blockIdx.x % 148u) * slice_words; span = slice_words; }
    u64 acc = 0;
    for (unsigned s = 0; s < steps; s++) {
        const u64 i = base + __umul64hi(x, span);
        acc += __ldg(t + i);
        if (MODE == 1) acc += __ldg(t + (i ^ mask_words));
        x = x * 6364136223846793005ULL + 1442695040888963407ULL;
    }
    if (acc == 0x1234567ull) *sink = acc;
}

// Load flavours of the 32-byte sector fetch (what does each one make the L2 bring in from DRAM?  run under ncu and read
// dram__bytes_read per launch): 0 = 8-byte __ldg; 1 = 256-bit load, no hints; 2 = 256-bit + L1::evict_last + L2 evict_last policy
// (the query kernels' fetch); 3 = 2 + L2::64B prefetch size; 4 = two 128-bit loads + the same hints; 5 = 256-bit L1::no_allocate;
// 6 = 256-bit with an L2 evict_first policy; 7 = 8-byte load with the hints of 2; 8 = 256-bit ld.global.cg (no .nc);
// 9 = 256-bit + L2 evict_normal policy; 10 = 8-byte + L2::64B; 11 = 256-bit L2::128B
template <int F>
__global__ void flavour(const u64* __restrict__ t, u64 n_words, unsigned steps, u64* sink) {
    u64 x = mix64(blockIdx.x * (u64)blockDim.x + threadIdx.x + 0x9E3779B97F4A7C15ULL);
    u64 pol_last, pol_first, pol_normal;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_normal));
    u64 acc = 0;
    for (unsigned s = 0; s < steps; s++) {
        const u64 i = __umul64hi(x, n_words) & ~3ull;          // aligned 32-byte sector
        const u64* p = t + i;
        u64 a = 0, b = 0, c = 0, d = 0;
        if (F == 0) a = __ldg(p);
        if (F == 1) asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
        if (F == 2) asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(pol_last));
        if (F == 3) asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(pol_last));
        if (F == 4) {
            asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(a), "=l"(b) : "l"(p), "l"(pol_last));
            asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(c), "=l"(d) : "l"(p + 2), "l"(pol_last));
        }
        if (F == 5) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
        if (F == 6) asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(pol_first));
        if (F == 7) asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(a) : "l"(p), "l"(pol_last));
        if (F == 8) asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
        if (F == 9) asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(pol_normal));
        if (F == 10) asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(a) : "l"(p));
        if (F == 11) asm volatile("ld.global.nc.L2::128B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
        acc += a ^ b ^ c ^ d;
        x = x * 6364136223846793005ULL + 1442695040888963407ULL;
    }
    if (acc == 0x1234567ull) *sink = acc;
}

__global__ void atomics(u64* counters, unsigned n, unsigned steps, u64* sink) {
    u64 x = mix64(blockIdx.x * (u64)blockDim.x + threadIdx.x + 77);
    u64 acc = 0;
    for (unsigned s = 0; s < steps; s++) {
        const unsigned d = (unsigned)__umul64hi(x, (u64)n);
        acc += atomicAdd(counters + d * 16, 1ull);          // counters 128 bytes apart
        x = x * 6364136223846793005ULL + 1442695040888963407ULL;
    }
    if (acc == 0x1234567ull) *sink = acc;
}

// hand-over: slot = atomicAdd(tail[d]); ring[d][slot % cap] = 32-byte record.  AGG: lanes of a warp going to the same ring share one atomic.
template <bool AGG>
__global__ void queue_push(u64* tails, uint4* rings, unsigned n, unsigned cap, unsigned steps, u64* sink) {
    u64 x = mix64(blockIdx.x * (u64)blockDim.x + threadIdx.x + 99);
    u64 acc = 0;
    for (unsigned s = 0; s < steps; s++) {
        const unsigned d = (unsigned)__umul64hi(x, (u64)n);
        u64 slot;
        if (AGG) {
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const unsigned leader = __ffs(peers) - 1, lane = threadIdx.x & 31u;
            u64 b = 0;
            if (lane == leader) b = atomicAdd(tails + d * 16, (u64)__popc(peers));
            b = __shfl_sync(0xffffffffu, b, leader);
            slot = b + __popc(peers & ((1u << lane) - 1u));
        } else slot = atomicAdd(tails + d * 16, 1ull);
        uint4* p = rings + ((u64)d * cap + (slot % cap)) * 2;
        p[0] = make_uint4((unsigned)x, (unsigned)(x >> 32), s, d);
        p[1] = make_uint4(d, s, (unsigned)slot, 7u);
        acc += slot;
        x = x * 6364136223846793005ULL + 1442695040888963407ULL;
    }
    if (acc == 0x1234567ull) *sink = acc;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: tlb_probe <table GB> <mode> ...\n"); return 1; }
    const double gb = atof(argv[1]);
    const u64 n_words = ((u64)(gb * (double)(1ull << 30)) / 8) & ~0xFFFFull;
    u64* t = nullptr; u64* sink = nullptr;
    CK(cudaMalloc(&t, n_words * 8)); CK(cudaMalloc(&sink, 8));
    fill<<<148 * 8, 256>>>(t, n_words);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const unsigned blocks = 148 * 8, threads = 256;
    for (int a = 2; a < argc; a++) {
        std::string m(argv[a]);
        unsigned steps = 256;
        double per_step = 1.0;
        u64 slice_words = 0, mask_words = 0; unsigned group = 1, nq = 0;
        int mode = -1;                // 0..3 gather modes, 4 atomics, 5 queue, 6 queue aggregated
        if (m == "single") mode = 0;
        else if (m.rfind("pair:", 0) == 0) { mode = 1; mask_words = strtoull(m.c_str() + 5, nullptr, 0) / 8; per_step = 2.0; }
        else if (m.rfind("slice:", 0) == 0) {
            mode = 2; const char* p = m.c_str() + 6; slice_words = (u64)(atof(p) * (1 << 20)) / 8;
            const char* c = strchr(p, ':'); group = c ? (unsigned)atoi(c + 1) : 1u;
            if (slice_words * ((147u / group) + 1) > n_words) { printf("{\"mode\": \"%s\", \"skipped\": \"table too small\"}\n", m.c_str()); continue; }
        }
        else if (m.rfind("bslice:", 0) == 0) {
            mode = 3; slice_words = (u64)(atof(m.c_str() + 7) * (1 << 20)) / 8;
            if (slice_words * 148 > n_words) { printf("{\"mode\": \"%s\", \"skipped\": \"table too small\"}\n", m.c_str()); continue; }
        }
        else if (m.rfind("gran:", 0) == 0) {           // cudaLimitMaxL2FetchGranularity for the launches that follow
            const size_t g = (size_t)atoi(m.c_str() + 5); size_t now = 0;
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
            cudaDeviceGetLimit(&now, cudaLimitMaxL2FetchGranularity);
            printf("{\"mode\": \"%s\", \"set\": \"%s\", \"now\": %zu}\n", m.c_str(), cudaGetErrorString(e), now);
            continue;
        }
        else if (m.rfind("flavour:", 0) == 0) { mode = 100 + atoi(m.c_str() + 8); }
        else if (m.rfind("atomic:", 0) == 0) { mode = 4; nq = (unsigned)atoi(m.c_str() + 7); }
        else if (m.rfind("queue:", 0) == 0) { mode = 5; nq = (unsigned)atoi(m.c_str() + 6); }
        else if (m.rfind("queueagg:", 0) == 0) { mode = 6; nq = (unsigned)atoi(m.c_str() + 9); }
        else { fprintf(stderr, "unknown mode %s\n", m.c_str()); continue; }
        const unsigned cap = 1u << 12;              // 4096 records of 32 bytes per ring
        auto run = [&](unsigned st) {
            switch (mode) {
                case 0: gather<0><<<blocks, threads>>>(t, n_words, 0, 1, 0, st, sink); break;
                case 1: gather<1><<<blocks, threads>>>(t, n_words, 0, 1, mask_words, st, sink); break;
                case 2: gather<2><<<blocks, threads>>>(t, n_words, slice_words, group, 0, st, sink); break;
                case 3: gather<3><<<blocks, threads>>>(t, n_words, slice_words, 1, 0, st, sink); break;
                case 4: atomics<<<blocks, threads>>>(t, nq, st, sink); break;
                case 5: queue_push<false><<<blocks, threads>>>(t, reinterpret_cast<uint4*>(t + 16 * 4096), nq, cap, st, sink); break;
                case 6: queue_push<true><<<blocks, threads>>>(t, reinterpret_cast<uint4*>(t + 16 * 4096), nq, cap, st, sink); break;
                case 100: flavour<0><<<blocks, threads>>>(t, n_words, st, sink); break;
                case 101: flavour<1><<<blocks, threads>>>(t, n_words, st, sink); break;
                case 102: flavour<2><<<blocks, threads>>>(t, n_words, st, sink); break;
                case 103: flavour<3><<<blocks, threads>>>(t, n_words, st, sink); break;
                case 104: flavour<4><<<blocks, threads>>>(t, n_words, st, sink); break;
                case 105: flavour<5><<<blocks, threads>>>(t, n_words, st, sink); break;
                case 106: flavour<6><<<blocks, threads>>>(t, n_words, st, sink); break;
                case 107: flavour<7><<<blocks, threads>>>(t, n_words, st, sink); break;
                case 108: flavour<8><<<blocks, threads>>>(t, n_words, st, sink); break;
                case 109: flavour<9><<<blocks, threads>>>(t, n_words, st, sink); break;
                case 110: flavour<10><<<blocks, threads>>>(t, n_words, st, sink); break;
                case 111: flavour<11><<<blocks, threads>>>(t, n_words, st, sink); break;
            }
        };
        if (mode >= 4 && mode < 100) CK(cudaMemset(t, 0, 16 * 4096 * 8));
        run(steps / 8);
        CK(cudaEventRecord(e0));
        run(steps);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)blocks * threads * steps;
        printf("{\"kind\": \"tlb_probe\", \"table_gb\": %.2f, \"mode\": \"%s\", \"ms\": %.3f, \"steps_per_s\": %.4g, \"loads_per_s\": %.4g}\n",
               gb, m.c_str(), ms, ops / (ms * 1e-3), ops * per_step / (ms * 1e-3));
        fflush(stdout);
        if (mode >= 4 && mode < 100) { fill<<<148 * 8, 256>>>(t, n_words); CK(cudaDeviceSynchronize()); }
    }
    return 0;
}
