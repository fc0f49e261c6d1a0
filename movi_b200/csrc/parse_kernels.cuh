// parse_kernels.cuh -- FASTA / FASTQ ingest on the device: raw file bytes in HBM -> concatenated bases, read offsets, ids.
//
// Restates BatchLoader::loadBatch + grabNextRead (reference src/batch_loader.cpp:26-143) for a chunk of a read file that starts
// at a record boundary: FASTA = a line starting with '>' followed by any number of sequence lines (trailing white space of
// every line dropped, lines concatenated); FASTQ = records of exactly four lines; the id is the header from its 2nd character
// through AND INCLUDING the first blank / tab / CR (:117-119).  The host parser (movi_host.cpp) does the same with memchr at
// ~1.8 GB/s per thread, which a GPU that walks 7e10 bases/s outruns 40 times over; here the work is line-parallel:
//   1. newline index: positions of all '\n' of the chunk (count per tile, scan of the tile counts, fill);
//   2. per line: header or sequence, trimmed length, id length;
//   3. exclusive scans over the lines: base offset, record index, id offset;
//   4. per line: copy the sequence bytes / the id to their places, write the record's offsets.
// A chunk that is not the last one of its file ends inside some record: that record (from its header line on) is left for the
// next chunk (`consumed`).  Anything the reference treats as an error or as an early end of input (a header that does not
// start with '>' / '@', a missing id, an empty line where a header is due) only raises a flag: the caller hands such a file to
// the host parser, which reproduces the reference's behaviour and messages.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace movi {

constexpr int PARSE_TILE = 8192;          // bytes per block in the newline passes (256 threads x 32 bytes)
constexpr int LINE_TILE = 1024;           // lines per block in the line passes
constexpr uint32_t PARSE_FLAG_BAD_HEADER = 1u, PARSE_FLAG_NO_ID = 2u, PARSE_FLAG_EMPTY_HEADER = 4u, PARSE_FLAG_OVERFLOW = 8u;

struct ParseResult {                      // filled on the device, copied to the host after the passes
    unsigned long long n_reads, n_bases, id_bytes, consumed;
    unsigned int flags, n_lines;
};

__device__ __forceinline__ bool parse_isspace(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }   // std::isspace in the "C" locale

// block-wide exclusive scan of one value per thread (256 or 1024 threads); returns the exclusive prefix, *total = block sum
template <typename T, int THREADS>
__device__ __forceinline__ T block_excl_scan(T v, T* total, T* sh /* THREADS / 32 entries */) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    T inc = v;
    for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (unsigned)o) inc += t; }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < THREADS / 32 ? sh[lane] : (T)0;
        T winc = w;
        for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= (unsigned)o) winc += t; }
        if (lane < THREADS / 32) sh[lane] = winc - w;
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    const T r = sh[warp] + inc - v;
    __syncthreads();
    return r;
}

// ---- 1. newline index
__device__ __forceinline__ uint32_t nl_mask32(const uint8_t* __restrict__ raw, unsigned long long base, unsigned long long n) {
    uint32_t m = 0;                        // bit k = raw[base + k] == '\n'
    if (base + 32 <= n) {
        const uint4 a = *reinterpret_cast<const uint4*>(raw + base), b = *reinterpret_cast<const uint4*>(raw + base + 16);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t x = w[k] ^ 0x0A0A0A0Au;                         // zero byte where '\n'
            const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);   // 0x80 in every zero byte
            m |= (((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u)) << (4 * k);
        }
    } else {
        for (unsigned k = 0; k < 32 && base + k < n; k++) if (raw[base + k] == '\n') m |= 1u << k;
    }
    return m;
}
__global__ void __launch_bounds__(256) parse_count_newlines(const uint8_t* __restrict__ raw, unsigned long long n, uint32_t* __restrict__ tile_cnt) {
    __shared__ uint32_t sh[8];
    const unsigned long long base = (unsigned long long)blockIdx.x * PARSE_TILE + threadIdx.x * 32ull;
    uint32_t c = base < n ? __popc(nl_mask32(raw, base, n)) : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31u) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t s = 0; for (int w = 0; w < 8; w++) s += sh[w]; tile_cnt[blockIdx.x] = s; }
}
// exclusive scan of up to 2^32 32-bit counts by one block (sequential over chunks of 1024); total -> *total_out
__global__ void __launch_bounds__(1024) parse_scan_u32(uint32_t* __restrict__ v, unsigned long long n, uint32_t* __restrict__ total_out) {
    __shared__ uint32_t sh[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (unsigned long long base = 0; base < n; base += 1024) {
        const unsigned long long i = base + threadIdx.x;
        const uint32_t x = i < n ? v[i] : 0u;
        uint32_t tot;
        __shared__ uint32_t tot_s;
        const uint32_t ex = block_excl_scan<uint32_t, 1024>(x, &tot_s, sh);
        const uint32_t carry = carry_s;
        if (i < n) v[i] = carry + ex;
        __syncthreads();
        tot = tot_s;
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}
__global__ void __launch_bounds__(256) parse_fill_newlines(const uint8_t* __restrict__ raw, unsigned long long n, const uint32_t* __restrict__ tile_off,
                                                            uint32_t* __restrict__ nlpos) {
    __shared__ uint32_t sh[8];
    __shared__ uint32_t tot;
    const unsigned long long base = (unsigned long long)blockIdx.x * PARSE_TILE + threadIdx.x * 32ull;
    uint32_t m = base < n ? nl_mask32(raw, base, n) : 0u;
    uint32_t at = tile_off[blockIdx.x] + block_excl_scan<uint32_t, 256>(__popc(m), &tot, sh);
    while (m) { const uint32_t k = __ffs(m) - 1; nlpos[at++] = (uint32_t)(base + k); m &= m - 1; }
}

// ---- 2. per line: kind, trimmed length, id length.  Line j = raw[start_j, end_j): start_0 = 0, start_j = nlpos[j - 1] + 1,
// end_j = nlpos[j]; in the last chunk of a file a final line without '\n' ends at n.
struct LineInfo { uint32_t start, seq_len; uint32_t id_len : 16, is_hdr : 1; };

__device__ __forceinline__ void line_bounds(const uint32_t* __restrict__ nlpos, uint32_t n_nl, unsigned long long n, uint32_t j, uint32_t& b, uint32_t& e) {
    b = j ? nlpos[j - 1] + 1u : 0u;
    e = j < n_nl ? nlpos[j] : (uint32_t)n;
}
__global__ void __launch_bounds__(256) parse_classify_lines(const uint8_t* __restrict__ raw, unsigned long long n, const uint32_t* __restrict__ nlpos,
                                                             uint32_t n_nl, uint32_t n_lines, int fastq, LineInfo* __restrict__ lines,
                                                             ParseResult* __restrict__ res) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_lines) return;
    uint32_t b, e;
    line_bounds(nlpos, n_nl, n, j, b, e);
    LineInfo li; li.start = b; li.seq_len = 0; li.id_len = 0; li.is_hdr = 0;
    const bool hdr = fastq ? (j & 3u) == 0u : (e > b && raw[b] == '>');
    const bool seq = fastq ? (j & 3u) == 1u : !hdr;
    if (hdr) {
        li.is_hdr = 1;
        if (e == b) atomicOr(&res->flags, PARSE_FLAG_EMPTY_HEADER);                       // grabNextRead: an empty line ends the input
        else if (raw[b] != (fastq ? '@' : '>')) atomicOr(&res->flags, PARSE_FLAG_BAD_HEADER);
        else if (e - b <= 2) atomicOr(&res->flags, PARSE_FLAG_NO_ID);
        else {
            uint32_t w = b + 1;
            while (w < e && raw[w] != ' ' && raw[w] != '\t' && raw[w] != '\r') w++;
            const uint32_t idl = w < e ? w - b : e - b;                                   // find_first_of position, or the header size
            const uint32_t take = min(idl, e - b - 1u);
            if (take > 0xFFFFu) atomicOr(&res->flags, PARSE_FLAG_OVERFLOW);
            li.id_len = take;
        }
    } else if (seq) {
        uint32_t t = e;
        while (t > b && parse_isspace(raw[t - 1])) t--;
        li.seq_len = t - b;
    }
    if (!fastq && j == 0 && !hdr) atomicOr(&res->flags, PARSE_FLAG_BAD_HEADER);           // a chunk starts at a record
    lines[j] = li;
}

// ---- 3. exclusive scans over the lines of (seq_len, is_hdr, id_len): tile sums, scan of the tile sums, then applied in pass 4
struct Sum3 { unsigned long long bases; uint32_t recs; uint32_t ids; };
__global__ void __launch_bounds__(LINE_TILE) parse_line_tile_sums(const LineInfo* __restrict__ lines, uint32_t n_lines, Sum3* __restrict__ tile) {
    __shared__ unsigned long long sb[32]; __shared__ uint32_t sr[32], si[32];
    const uint32_t j = blockIdx.x * LINE_TILE + threadIdx.x;
    unsigned long long b = 0; uint32_t r = 0, i = 0;
    if (j < n_lines) { const LineInfo li = lines[j]; b = li.seq_len; r = li.is_hdr; i = li.id_len; }
    for (int o = 16; o > 0; o >>= 1) { b += __shfl_down_sync(0xffffffffu, b, o); r += __shfl_down_sync(0xffffffffu, r, o); i += __shfl_down_sync(0xffffffffu, i, o); }
    if ((threadIdx.x & 31u) == 0) { sb[threadIdx.x >> 5] = b; sr[threadIdx.x >> 5] = r; si[threadIdx.x >> 5] = i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        Sum3 s{0, 0, 0};
        for (int w = 0; w < LINE_TILE / 32; w++) { s.bases += sb[w]; s.recs += sr[w]; s.ids += si[w]; }
        tile[blockIdx.x] = s;
    }
}
__global__ void parse_scan_tiles(Sum3* __restrict__ tile, uint32_t n_tiles, Sum3* __restrict__ total) {   // one thread: n_tiles is small (lines / 1024)
    if (blockIdx.x || threadIdx.x) return;
    Sum3 run{0, 0, 0};
    for (uint32_t t = 0; t < n_tiles; t++) { const Sum3 s = tile[t]; tile[t] = run; run.bases += s.bases; run.recs += s.recs; run.ids += s.ids; }
    *total = run;
}

// ---- 4. apply: one warp per line copies the line's bytes; header lines also write the record's offsets.
// n_keep = records handed out (all of them in the last chunk of a file, all but the last one otherwise; FASTQ: only records whose
// four lines are all present).
__global__ void __launch_bounds__(LINE_TILE) parse_apply(const uint8_t* __restrict__ raw, const LineInfo* __restrict__ lines, uint32_t n_lines,
                                                          const Sum3* __restrict__ tile, uint32_t n_keep, uint8_t* __restrict__ bases,
                                                          unsigned long long* __restrict__ offsets, uint8_t* __restrict__ ids,
                                                          unsigned long long* __restrict__ id_offsets, ParseResult* __restrict__ res, int reverse) {
    __shared__ unsigned long long sb[32]; __shared__ uint32_t sr[32], si[32];
    __shared__ unsigned long long tb; __shared__ uint32_t tr, ti;
    __shared__ unsigned long long s_boff[LINE_TILE]; __shared__ uint32_t s_rec[LINE_TILE], s_ioff[LINE_TILE];
    const uint32_t j = blockIdx.x * LINE_TILE + threadIdx.x;
    LineInfo li; li.start = 0; li.seq_len = 0; li.id_len = 0; li.is_hdr = 0;
    if (j < n_lines) li = lines[j];
    const Sum3 t0 = tile[blockIdx.x];
    const unsigned long long boff = t0.bases + block_excl_scan<unsigned long long, LINE_TILE>((unsigned long long)li.seq_len, &tb, sb);
    const uint32_t rec = t0.recs + block_excl_scan<uint32_t, LINE_TILE>((uint32_t)li.is_hdr, &tr, sr);      // records that start before this line
    const uint32_t ioff = t0.ids + block_excl_scan<uint32_t, LINE_TILE>((uint32_t)li.id_len, &ti, si);
    s_boff[threadIdx.x] = boff; s_rec[threadIdx.x] = rec; s_ioff[threadIdx.x] = ioff;
    if (j < n_lines && li.is_hdr) {
        if (rec < n_keep) { offsets[rec] = boff; id_offsets[rec] = ioff; }
        if (rec == n_keep) {                                   // the first record left out: totals and the bytes consumed end here
            offsets[rec] = boff; id_offsets[rec] = ioff;
            res->n_bases = boff; res->id_bytes = ioff; res->consumed = li.start;
        }
    }
    __syncthreads();
    // copies: warp w of the block takes lines w, w + 32, ... of the tile
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t k = warp; k < LINE_TILE; k += LINE_TILE / 32) {
        const uint32_t jj = blockIdx.x * LINE_TILE + k;
        if (jj >= n_lines) break;
        const LineInfo l = lines[jj];
        if (l.is_hdr) {
            if (s_rec[k] < n_keep) for (uint32_t b = lane; b < l.id_len; b += 32) ids[s_ioff[k] + b] = raw[l.start + 1u + b];
        } else if (l.seq_len && s_rec[k] >= 1 && s_rec[k] - 1 < n_keep) {
            uint8_t* dst = bases + s_boff[k];
            const uint8_t* src = raw + l.start;
            for (uint32_t b = lane; b < l.seq_len; b += 32) dst[b] = src[b];
        }
    }
    (void)reverse;
}
// when every record is kept there is no "first record left out": the totals come from the grand totals
__global__ void parse_finish(const Sum3* __restrict__ total, uint32_t n_keep, unsigned long long n, unsigned long long* __restrict__ offsets,
                             unsigned long long* __restrict__ id_offsets, ParseResult* __restrict__ res) {
    if (blockIdx.x || threadIdx.x) return;
    if (n_keep == total->recs) {
        offsets[n_keep] = total->bases; id_offsets[n_keep] = total->ids;
        res->n_bases = total->bases; res->id_bytes = total->ids; res->consumed = n;
    }
    res->n_reads = n_keep;
}

}  // namespace movi
