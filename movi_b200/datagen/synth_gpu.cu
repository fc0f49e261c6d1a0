// synth_gpu.cu -- human-scale synthetic move tables generated directly in HBM (BENCH DATA GENERATOR, not product).
//
// BASELINE.json configs[3] asks for a row table "sized to fit 180 GB HBM".  A 32 x 1 Gbp text cannot be suffix-sorted on
// the bench host, and a table of 10^10 rows cannot be built on the CPU and pushed through PCIe within a benchmark run, so
// the table is synthesised on the device, following SURVEY 8(d)/Appendix B: ANY sequence of run heads (adjacent runs
// differ) and run lengths defines a valid LF permutation, LF(i) = C[c] + rank_c(i), which is all the query needs.
//
//   run i:  c_i = 2*(i & 1) + (hash(i) & 1)      -> even runs are A/C, odd runs are G/T: adjacent runs always differ
//           n_i = 1 + geometric-ish length from hash(i)  (mean ~ 8, max 2047); one run E is the '$' run (n = 1, c field 0)
//   pass 1: block-wise exclusive scans of (n, n*[c==0], .., n*[c==3]) -> start position p_i of every run and rank_c(i)
//   pass 2: lf_i = C[c_i] + rank_{c_i}(i)   (lf = 0 for '$'),  id_i = the run that contains lf_i (binary search over p),
//           offset_i = lf_i - p_{id_i};  threshold bits = hash bits, forced "down" / "up" near the table ends so that a
//           reposition always finds the character in the chosen direction
// Rows are written in the device format of movi_b200/csrc/dev_index.cuh (= the reference's regular-thresholds MoveRow,
// include/move_row.hpp:131-142) and handed to the query context with movi_gpu_adopt_device_rows().
// Reads that "match" are made by LF-walking the finished table (emit the row character, LF, repeat).
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <cuda_runtime.h>

#include "../../include/movi_gpu.h"
#include "../csrc/dev_index.cuh"          // the device row format the adopted table is rewritten into (make_reads walks it)

namespace {

typedef unsigned long long u64;

__host__ __device__ inline u64 mix(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
struct Gen { u64 r, end_idx, seed; uint32_t profile; };   // profile 0: short runs (mean ~8), 1: heavy-tailed runs (mean 55), 2: even runs (mean 56)

__host__ __device__ inline uint32_t run_c(const Gen& g, u64 i) { return i == g.end_idx ? 0u : (uint32_t)(2 * (i & 1) + (mix(i ^ g.seed) & 1)); }
__host__ __device__ inline uint32_t run_n(const Gen& g, u64 i) {
    if (i == g.end_idx) return 1u;
    const u64 h = mix(i * 0x9E3779B97F4A7C15ULL + g.seed);
    // mixture: mostly short runs, a tail of long ones (mean ~ 8)
    const uint32_t k = (uint32_t)(h & 0xFF);
    if (g.profile == 2)                                // n/r = 56, run lengths 48..64: the LF image of a run spans at most two rows, so a
        return 48u + (uint32_t)((h >> 8) % 17);        // walk fast-forwards 0.5 rows per base (suffix-sorted indexes: 0.22 - 0.27; profile 1: 2.0)
    if (g.profile == 1) {                              // n/r ~ 40, like the c2 pangenome index (n/r = 41)
        uint32_t m = 1 + (uint32_t)((h >> 8) % 64);
        if (k < 32) m = 64 + (uint32_t)((h >> 20) % 192);
        if (k < 3) m = 256 + (uint32_t)((h >> 32) % 1024);
        return m;
    }
    uint32_t n = 1 + (uint32_t)((h >> 8) % 6);
    if (k < 40) n = 8 + (uint32_t)((h >> 20) % 24);
    if (k < 4) n = 64 + (uint32_t)((h >> 32) % 448);
    return n;
}

constexpr int SB = 1024;              // rows per scan block (one thread per row)
struct Sum5 { u64 v[5]; };            // total, per character

__device__ inline Sum5 contrib(const Gen& g, u64 i) {
    Sum5 s; const uint32_t n = run_n(g, i), c = run_c(g, i);
    s.v[0] = n;
    for (int k = 0; k < 4; k++) s.v[k + 1] = (i != g.end_idx && c == (uint32_t)k) ? n : 0;
    return s;
}

// pass 1a: per-block totals
__global__ void block_totals(Gen g, u64 n_blocks, Sum5* __restrict__ tot) {
    __shared__ u64 sh[5][SB / 32];
    const u64 i = blockIdx.x * (u64)SB + threadIdx.x;
    Sum5 s{}; if (i < g.r) s = contrib(g, i);
    for (int k = 0; k < 5; k++) {
        u64 v = s.v[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) { u64 v = 0; for (int w = 0; w < SB / 32; w++) v += sh[threadIdx.x][w]; tot[blockIdx.x].v[threadIdx.x] = v; }
}
// pass 1b: exclusive scan of the block totals (single block, sequential over chunks)
__global__ void scan_totals(Sum5* __restrict__ tot, u64 n_blocks, Sum5* __restrict__ grand) {
    __shared__ u64 s[SB]; __shared__ u64 carry;
    for (int k = 0; k < 5; k++) {
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (u64 base = 0; base < n_blocks; base += SB) {
            const u64 i = base + threadIdx.x;
            const u64 v = i < n_blocks ? tot[i].v[k] : 0;
            s[threadIdx.x] = v; __syncthreads();
            for (int o = 1; o < SB; o <<= 1) { u64 t = threadIdx.x >= o ? s[threadIdx.x - o] : 0; __syncthreads(); s[threadIdx.x] += t; __syncthreads(); }
            if (i < n_blocks) tot[i].v[k] = carry + s[threadIdx.x] - v;
            __syncthreads();
            if (threadIdx.x == SB - 1) carry += s[threadIdx.x];
            __syncthreads();
        }
        if (threadIdx.x == 0) grand->v[k] = carry;
        __syncthreads();
    }
}
// pass 1c: start position of every run -> p[i]; LF target of every run head -> lf[i] (stored in the row array for now)
__global__ void positions_and_lf(Gen g, const Sum5* __restrict__ tot, const Sum5* __restrict__ grand, u64* __restrict__ p, u64* __restrict__ lf) {
    __shared__ u64 s[SB];
    const u64 i = blockIdx.x * (u64)SB + threadIdx.x;
    Sum5 mine{}; if (i < g.r) mine = contrib(g, i);
    const uint32_t c = i < g.r ? run_c(g, i) : 0;
    u64 excl[5];
    for (int k = 0; k < 5; k++) {
        s[threadIdx.x] = mine.v[k]; __syncthreads();
        for (int o = 1; o < SB; o <<= 1) { u64 t = threadIdx.x >= o ? s[threadIdx.x - o] : 0; __syncthreads(); s[threadIdx.x] += t; __syncthreads(); }
        excl[k] = tot[blockIdx.x].v[k] + s[threadIdx.x] - mine.v[k];
        __syncthreads();
    }
    if (i < g.r) {
        p[i] = excl[0];
        u64 C = 1;                                   // F column: '$' first, then the characters in order
        for (uint32_t k = 0; k < c; k++) C += grand->v[k + 1];
        lf[i] = (i == g.end_idx) ? 0 : C + excl[c + 1];
    }
    if (i == g.r) p[i] = excl[0];                    // sentinel p[r] = n (one extra slot)
}
__device__ inline u64 run_of(const u64* __restrict__ p, u64 r, u64 pos) {   // largest i with p[i] <= pos
    u64 lo = 0, hi = r;
    while (hi - lo > 1) { const u64 mid = (lo + hi) >> 1; if (p[mid] <= pos) lo = mid; else hi = mid; }
    return lo;
}
// pass 2: rows in place (row i currently holds lf_i)
__global__ void make_rows(Gen g, const u64* __restrict__ p, uint2* __restrict__ rows, u64 guard) {
    u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; i < g.r; i += stride) {
        const u64 lf = reinterpret_cast<const u64*>(rows)[i];
        const u64 id = run_of(p, g.r, lf);
        const uint32_t off = (uint32_t)(lf - p[id]);
        const uint32_t c = run_c(g, i), n = run_n(g, i);
        uint32_t thr = (uint32_t)(mix(i + 0x1234567ULL + g.seed) >> 40) & 7u;
        if (i < guard) thr = 0;                      // too close to the top: every character is guaranteed only below
        if (i + guard >= g.r) thr = 7;               // too close to the bottom: only above
        uint2 v;
        v.x = (uint32_t)(id & 0xFFFFFFFFull);
        v.y = (n & 0x7FFu) | (((thr >> 1) & 1u) << 11) | (((thr >> 2) & 1u) << 12) | ((c & 7u) << 13) |
              ((off & 0x7FFu) << 16) | ((thr & 1u) << 27) | ((uint32_t)((id >> 32) & 0xFull) << 28);
        rows[i] = v;
    }
}
// matching reads: LF walks over the finished table with substitutions; every `rand_every`-th read is uniform random
__global__ void make_reads(Gen g, const uint2* __restrict__ rows, uint8_t* __restrict__ bases, u64 n_reads, uint32_t len, uint32_t sub_per_1024,
                           uint32_t rand_every) {
    u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const char ACGT[4] = {'A', 'C', 'G', 'T'};
    for (; t < n_reads; t += stride) {
        u64 h = mix(t * 0xD1B54A32D192ED03ULL + g.seed + 99);
        u64 idx = h % g.r; if (idx == g.end_idx) idx = (idx + 1) % g.r;
        uint32_t off = 0;
        uint8_t* out = bases + t * (u64)len;
        const bool rnd = rand_every && (t % rand_every == rand_every - 1);
        for (uint32_t k = 0; k < len; k++) {
            h = mix(h + k);
            uint2 v = rows[idx];
            uint32_t c = movi::row_c(v.y);
            if (idx == g.end_idx) { idx = (idx + 1 < g.r) ? idx + 1 : 0; off = 0; v = rows[idx]; c = movi::row_c(v.y); }   // never emit '$'
            uint32_t e = c;
            if (rnd || (h & 1023u) < sub_per_1024) e = (uint32_t)((h >> 12) & 3u);
            out[len - 1 - k] = (uint8_t)ACGT[e & 3u];
            // LF_move (reference src/move_structure.cpp:59-87)
            off += movi::row_off(v.y);
            idx = movi::row_id<u64>(v);
            for (;;) { const uint32_t n = movi::row_n(rows[idx].y); if (idx < g.r - 1 && off >= n) { off -= n; idx++; } else break; }
        }
    }
}

// Reads that behave like reads of a real pangenome index: the generator runs the QUERY's own walk (query_pml, reference
// src/move_structure_query.cpp:234-350) and writes down the characters it meets.  Where it decides on a substitution it picks a
// different character and repositions exactly as the query will (threshold bit -> up / down, scan to the nearest run of that
// character, :513-601), then carries on from there -- so after every error the read keeps matching along the new path and
// its PMLs climb again, as they do after a SNP on a suffix-sorted index (whose thresholds are LCP-guided).  With plain
// LF-walk reads (make_reads above) a read never re-synchronises on a random table: its first error leaves the walk on an
// unrelated path and the rest of the read runs in the worst-case reposition regime.  The first `lead` processed bases are
// uniform random: they scatter the walks, which all start at the last row like every query, over the whole table.
__global__ void make_reads_resync(Gen g, const uint2* __restrict__ rows, uint8_t* __restrict__ bases, u64 n_reads, uint32_t len,
                                  uint32_t sub_per_1024, uint32_t rand_every, uint32_t lead, u64 read_seed) {
    u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const char ACGT[4] = {'A', 'C', 'G', 'T'};
    const u64 r1 = g.r - 1;
    const uint32_t n_last = movi::row_n(rows[r1].y);
    for (; t < n_reads; t += stride) {
        u64 h = mix(t * 0xD1B54A32D192ED03ULL + g.seed + 99 + read_seed * 0x9E3779B97F4A7C15ULL);
        u64 idx = r1;
        uint32_t off = n_last - 1;
        uint8_t* out = bases + t * (u64)len;
        const bool rnd = rand_every && (t % rand_every == rand_every - 1);
        for (uint32_t k = 0; k < len; k++) {
            h = mix(h + k);
            uint2 v = rows[idx];
            const uint32_t c = movi::row_c(v.y) & 3u;         // the '$' row decodes as alphabet[0] in PML
            uint32_t e = c;
            if (rnd || k < lead) e = (uint32_t)((h >> 12) & 3u);
            else if ((h & 1023u) < sub_per_1024) e = (c + 1u + (uint32_t)((h >> 12) % 3u)) & 3u;
            if (e != c) {
                const uint32_t n = movi::row_n(v.y);
                const uint32_t thr = (idx == g.end_idx) ? 0u : ((v.y >> e) & (1u << 25)) ? n : 0u;     // row_thr of dev_index.cuh
                const bool down = off >= thr;
                u64 j = idx; bool found = false; uint2 w = v;
                for (uint32_t s = 0; s < 4096; s++) {
                    if (down ? j == r1 : j == 0) break;
                    j = down ? j + 1 : j - 1;
                    w = rows[j];
                    if ((movi::row_c(w.y) & 7u) == e) { found = true; break; }
                }
                if (found) { idx = j; v = w; off = down ? 0u : movi::row_n(w.y) - 1u; }
                else e = c;                                   // the scan would leave the table: keep the matching character
            }
            out[len - 1 - k] = (uint8_t)ACGT[e];
            off += movi::row_off(v.y);                        // LF_move + fast_forward (src/move_structure.cpp:59-87, 524-545)
            idx = movi::row_id<u64>(v);
            for (;;) { const uint32_t n = movi::row_n(rows[idx].y); if (idx < r1 && off >= n) { off -= n; idx++; } else break; }
        }
    }
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "synth_gpu: %s: %s\n", #call, cudaGetErrorString(e_)); return 2; } } while (0)

}  // namespace

extern "C" {

// Build a synthetic table of r rows on the current device of `ctx` and make it the context's index.
// Returns 0 on success; *n_out = BWT length.
int movi_synth_index(movi_gpu_ctx* ctx, uint64_t r, uint64_t seed, uint32_t profile, uint64_t* n_out, MoviGpuIndexDesc* desc_out) {
    if (!ctx || r < 1024) return 1;
    Gen g{r, mix(seed) % (r / 2) + r / 4, seed, profile};
    const u64 n_blocks = (r + 1 + SB - 1) / SB;
    uint2* rows = nullptr; u64* p = nullptr; Sum5* tot = nullptr; Sum5* grand = nullptr;
    CK(cudaMalloc(&rows, (r + 4) * sizeof(uint2)));
    CK(cudaMalloc(&p, (r + 1) * sizeof(u64)));
    CK(cudaMalloc(&tot, n_blocks * sizeof(Sum5)));
    CK(cudaMalloc(&grand, sizeof(Sum5)));
    block_totals<<<(unsigned)n_blocks, SB>>>(g, n_blocks, tot);
    scan_totals<<<1, SB>>>(tot, n_blocks, grand);
    positions_and_lf<<<(unsigned)n_blocks, SB>>>(g, tot, grand, p, reinterpret_cast<u64*>(rows));
    make_rows<<<148 * 16, 256>>>(g, p, rows, 64);
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    Sum5 hg;
    CK(cudaMemcpy(&hg, grand, sizeof(hg), cudaMemcpyDeviceToHost));
    if (n_out) *n_out = hg.v[0];

    // side tables (reference find_base_interval_data, src/move_structure_build.cpp:694-731): F-intervals of the characters
    MoviGpuIndexDesc d{};
    static uint64_t alphamap[256];
    for (int b = 0; b < 256; b++) alphamap[b] = 256;
    const char* al = "ACGT";
    for (int k = 0; k < 4; k++) { alphamap[(unsigned char)al[k]] = k; d.alphabet[k] = (uint8_t)al[k]; }
    d.mode = MOVI_MODE_REGULAR_THRESHOLDS; d.sigma = 4; d.length = hg.v[0]; d.r = r; d.end_bwt_idx = g.end_idx;
    d.alphamap = alphamap;
    for (int k = 0; k < 4; k++) d.end_bwt_idx_thresholds[k] = 0;      // '$' row: always reposition down (it sits mid-table)
    auto locate = [&](u64 pos, uint64_t* run, uint64_t* off) -> int {      // host binary search with single-element copies
        u64 lo = 0, hi = r;
        while (hi - lo > 1) { u64 mid = (lo + hi) >> 1, v; if (cudaMemcpy(&v, p + mid, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return 2; if (v <= pos) lo = mid; else hi = mid; }
        u64 v; if (cudaMemcpy(&v, p + lo, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return 2;
        *run = lo; *off = pos - v; return 0;
    };
    u64 first = 1;                                   // position 0 of F is '$'
    d.first_runs[0] = d.first_offsets[0] = d.last_runs[0] = d.last_offsets[0] = 0;
    for (int k = 0; k < 4; k++) {
        const u64 cnt = hg.v[k + 1];
        if (locate(first, &d.first_runs[k + 1], &d.first_offsets[k + 1])) return 2;
        if (locate(first + cnt - 1, &d.last_runs[k + 1], &d.last_offsets[k + 1])) return 2;
        first += cnt;
    }
    cudaFree(p); cudaFree(tot); cudaFree(grand);
    if (desc_out) { *desc_out = d; desc_out->rows = nullptr; }
    const int rc = movi_gpu_adopt_device_rows(ctx, &d, rows);
    if (rc) cudaFree(rows);                          // ownership only moves on success
    return rc;
}

// Fill d_bases (n_reads * len bytes, device) with reads that match the table of a context built by movi_synth_index.
int movi_synth_reads(const void* d_rows, uint64_t r, uint64_t seed, void* d_bases, uint64_t n_reads, uint32_t len,
                     uint32_t sub_per_1024, uint32_t rand_every) {
    Gen g{r, mix(seed) % (r / 2) + r / 4, seed, 0};   // the walk only reads the finished rows
    make_reads<<<148 * 8, 128>>>(g, static_cast<const uint2*>(d_rows), static_cast<uint8_t*>(d_bases), n_reads, len, sub_per_1024, rand_every);
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    return 0;
}

// The same with reads that re-synchronise after every error (make_reads_resync); read_seed selects the batch.
int movi_synth_reads_resync(const void* d_rows, uint64_t r, uint64_t seed, void* d_bases, uint64_t n_reads, uint32_t len,
                            uint32_t sub_per_1024, uint32_t rand_every, uint32_t lead, uint64_t read_seed) {
    Gen g{r, mix(seed) % (r / 2) + r / 4, seed, 0};
    make_reads_resync<<<148 * 16, 128>>>(g, static_cast<const uint2*>(d_rows), static_cast<uint8_t*>(d_bases), n_reads, len, sub_per_1024,
                                         rand_every, lead, read_seed);
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    return 0;
}


}  // extern "C"
