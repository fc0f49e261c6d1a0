// movi_host.cpp -- host-side helpers either side of the GPU query path (include/movi_host.h).
// Pure C++ (no CUDA): index.movi reader, FASTA/FASTQ reader, .bpf / text / count writers, classifier.
#include "../../include/movi_host.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <zlib.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <atomic>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return 1; }

constexpr uint32_t MOVI_MAGIC = 0x4D4F5649u;   // reference include/utils.hpp:29
constexpr uint32_t BPF_MAGIC = 0x42504600u;    // reference include/utils.hpp:26

}  // namespace

struct movi_host_index {
    int fd = -1;
    const uint8_t* map = nullptr;
    size_t size = 0;
    MoviGpuIndexDesc d{};
    uint64_t original_r = 0;
    std::vector<uint64_t> alphamap;
    std::vector<uint16_t> sep_thr;
    std::vector<uint64_t> sep_map;
    std::vector<uint32_t> id_blocks;      // aligned copy
    std::string type_name;
};

struct movi_host_reads {
    std::string bases;
    std::vector<uint64_t> offsets{0};
    std::string ids;                      // concatenated ids
    std::vector<uint64_t> id_offsets{0};
};

namespace {
// Batches are big (10 bytes per base and more) and short-lived: a freed one keeps its buffers and serves the next parse, so that
// a streaming run touches fresh pages for its first few batches only (page faults on new allocations were half of the parser's
// time: 1.0 -> 2.1 GB/s on a 322 MB FASTA).  At most three objects wait here, none holding more than 6 GB.
std::mutex g_pool_mu;
std::vector<movi_host_reads*> g_pool;
movi_host_reads* pool_take() {
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_pool.empty()) { movi_host_reads* rd = g_pool.back(); g_pool.pop_back(); return rd; }
    }
    return new movi_host_reads();
}
void pool_give(movi_host_reads* rd) {
    if (!rd) return;
    const size_t held = rd->bases.capacity() + rd->ids.capacity() + 8 * (rd->offsets.capacity() + rd->id_offsets.capacity());
    if (held >= (64u << 10) && held <= (6ull << 30)) {
        rd->bases.clear(); rd->ids.clear(); rd->offsets.assign(1, 0); rd->id_offsets.assign(1, 0);
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_pool.size() < 3) { g_pool.push_back(rd); return; }
    }
    delete rd;
}
struct PoolReturn { void operator()(movi_host_reads* rd) const { pool_give(rd); } };
}  // namespace

namespace {

struct Cursor {
    const uint8_t* p; size_t size; size_t pos = 0; bool ok = true;
    bool need(size_t n) { if (!ok || pos + n > size || pos + n < pos) { ok = false; return false; } return true; }
    template <typename T> T get() { T v{}; if (need(sizeof(T))) { std::memcpy(&v, p + pos, sizeof(T)); pos += sizeof(T); } return v; }
    const uint8_t* take(size_t n) { if (!need(n)) return nullptr; const uint8_t* q = p + pos; pos += n; return q; }
    // a * b [* c] bytes, the product checked against overflow (sizes come from the file)
    const uint8_t* take_n(uint64_t a, uint64_t b, uint64_t c = 1) {
        unsigned __int128 n = (unsigned __int128)a * b * c;
        if ((a && b && c && (n / a / b != c)) || n > size) { ok = false; return nullptr; }
        return take((size_t)n);
    }
};

}  // namespace

extern "C" {

const char* movi_host_last_error(void) { return g_err.c_str(); }

// MoveStructure::deserialize, reference src/move_structure_io.cpp:471-511 (modern header only)
int movi_host_index_open(const char* path, movi_host_index** out) {
    if (!path || !out) return fail("null argument");
    *out = nullptr;
    std::string file(path);
    struct stat sb;
    if (stat(file.c_str(), &sb) == 0 && S_ISDIR(sb.st_mode)) file += "/index.movi";     // io.cpp:16-22
    int fd = open(file.c_str(), O_RDONLY);
    if (fd < 0) return fail("[deserialize index] Failed to open the index file at: " + std::string(path));
    if (fstat(fd, &sb) != 0 || sb.st_size < 64) { close(fd); return fail("index file too small: " + file); }
    void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { close(fd); return fail("mmap failed for " + file); }
    movi_host_index* ix = new movi_host_index();
    ix->fd = fd; ix->map = static_cast<const uint8_t*>(m); ix->size = (size_t)sb.st_size;
    Cursor c{ix->map, ix->size};
    MoviGpuIndexDesc& d = ix->d;
    auto bad = [&](const std::string& why) { movi_host_index_close(ix); return fail("[deserialize index] " + why); };

    // MoviHeader, include/utils.hpp:32-61: u32 magic; u8 version[3]; char type; u8 reserved; pad; 4 x u64
    const uint8_t* h = c.take(48);
    if (!h) return bad("truncated header");
    uint32_t magic; std::memcpy(&magic, h, 4);
    if (magic != MOVI_MAGIC) return bad("Invalid magic number in header. Not a valid Movi 2.0 index file.");
    d.mode = (int32_t)(int8_t)h[7];
    std::memcpy(&d.length, h + 16, 8); std::memcpy(&d.r, h + 24, 8);
    std::memcpy(&ix->original_r, h + 32, 8); std::memcpy(&d.end_bwt_idx, h + 40, 8);
    size_t row_bytes;
    switch (d.mode) {
        case MOVI_MODE_REGULAR: row_bytes = 8; ix->type_name = "regular"; break;                         /* src/utils.cpp:20-24 */
        case MOVI_MODE_REGULAR_THRESHOLDS: row_bytes = 8; ix->type_name = "regular-thresholds"; break;
        case MOVI_MODE_BLOCKED: row_bytes = 6; ix->type_name = "blocked"; break;
        case MOVI_MODE_SAMPLED: row_bytes = 3; ix->type_name = "sampled"; break;
        case MOVI_MODE_BLOCKED_THRESHOLDS: row_bytes = 6; ix->type_name = "blocked-thresholds"; break;
        case MOVI_MODE_SAMPLED_THRESHOLDS: row_bytes = 3; ix->type_name = "sampled-thresholds"; break;
        default: return bad("Index mode " + std::to_string(d.mode) + " is not supported by the GPU query path "
                            "(regular-thresholds, blocked-thresholds, sampled-thresholds only)");
    }
    // read_basic_index_data, io.cpp:145-184
    for (int k = 0; k < 4; k++) d.end_bwt_idx_thresholds[k] = c.get<uint64_t>();
    c.take(64);                                          // end_bwt_idx_next_down / next_up (constant mode only)
    const uint64_t alphamap_size = c.get<uint64_t>();
    if (!c.ok || alphamap_size != 256) return bad("unexpected alphamap size");
    ix->alphamap.resize(256);
    for (auto& a : ix->alphamap) a = c.get<uint64_t>();
    d.alphamap = ix->alphamap.data();
    const uint64_t sigma = c.get<uint64_t>();
    if (!c.ok || sigma == 0 || sigma > 7) return bad("unexpected alphabet size");
    d.sigma = (uint32_t)sigma;
    const uint8_t* al = c.take(sigma);
    if (!al) return bad("truncated alphabet");
    std::memcpy(d.alphabet, al, sigma);
    c.take(3);                                           // u16 nt_splitting; u8 constant
    // read_main_table, io.cpp:361-397
    d.rows = c.take_n(d.r, row_bytes);
    if (!d.rows) return bad("truncated row table");
    // read_tally_table, io.cpp:338-349
    if (d.mode == MOVI_MODE_SAMPLED_THRESHOLDS || d.mode == MOVI_MODE_SAMPLED) {
        d.tally_checkpoints = c.get<uint32_t>();
        d.tally_len = c.get<uint64_t>();
        d.tally_ids = c.take_n(sigma, d.tally_len, 5);
        if (!d.tally_ids) return bad("truncated id checkpoint table");
        if (d.tally_checkpoints == 0 || d.tally_len < d.r / d.tally_checkpoints + 1) return bad("id checkpoint table inconsistent with the number of rows");
    }
    // read_overflow_tables, io.cpp:219-237
    uint64_t sz = c.get<uint64_t>(); c.take_n(sz, 8);
    sz = c.get<uint64_t>(); c.take_n(sz, 8);
    sz = c.get<uint64_t>(); c.take_n(sz, 8, sigma - 1);
    // read_counts_data, io.cpp:269-287
    sz = c.get<uint64_t>();
    if (!c.ok || sz != sigma) return bad("unexpected counts size");
    c.take(sz * 8);
    const uint64_t k1 = c.get<uint64_t>();
    if (!c.ok || k1 != sigma + 1 || k1 > 8) return bad("unexpected first/last runs size");
    for (uint64_t k = 0; k < k1; k++) d.last_runs[k] = c.get<uint64_t>();
    for (uint64_t k = 0; k < k1; k++) d.last_offsets[k] = c.get<uint64_t>();
    for (uint64_t k = 0; k < k1; k++) d.first_runs[k] = c.get<uint64_t>();
    for (uint64_t k = 0; k < k1; k++) d.first_offsets[k] = c.get<uint64_t>();
    // read_id_blocks, io.cpp:305-324
    if (d.mode == MOVI_MODE_BLOCKED_THRESHOLDS || d.mode == MOVI_MODE_BLOCKED) {
        d.id_blocks_size = c.get<uint64_t>();
        const uint8_t* b = c.take_n(sigma, d.id_blocks_size, 4);
        if (!b && d.id_blocks_size) return bad("truncated id_blocks");
        ix->id_blocks.resize((size_t)sigma * d.id_blocks_size);
        if (b) std::memcpy(ix->id_blocks.data(), b, ix->id_blocks.size() * 4);
        d.id_blocks = ix->id_blocks.data();
        d.block_size = c.get<uint64_t>();
        if (c.ok && (d.block_size == 0 || d.id_blocks_size < (d.r + d.block_size - 1) / d.block_size)) return bad("id_blocks inconsistent with the number of rows");
    }
    if (!c.ok) return bad("truncated index file");
    // read_separators_thresholds, io.cpp:415-433 (present iff use_separator(), move_structure.cpp:547-552)
    const bool no_thresholds = d.mode == MOVI_MODE_REGULAR || d.mode == MOVI_MODE_BLOCKED || d.mode == MOVI_MODE_SAMPLED;
    if (sigma == 5 && d.alphabet[0] == '%' && !no_thresholds) {                    // (#if USE_THRESHOLDS, io.cpp:504-508)
        d.sep_count = c.get<uint64_t>();
        const uint8_t* t = c.take_n(d.sep_count, 8);
        if (!c.ok) return bad("truncated separators thresholds");
        ix->sep_thr.resize(d.sep_count * 4);
        if (t) std::memcpy(ix->sep_thr.data(), t, d.sep_count * 8);
        d.sep_map_size = c.get<uint64_t>();
        const uint8_t* mp = c.take_n(d.sep_map_size, 16);
        if (!c.ok) return bad("truncated separators map");
        ix->sep_map.resize(d.sep_map_size * 2);
        if (mp) std::memcpy(ix->sep_map.data(), mp, d.sep_map_size * 16);
        d.separators_thresholds = ix->sep_thr.data();
        d.separators_map = ix->sep_map.data();
    }
    *out = ix;
    return 0;
}

int movi_host_index_desc(const movi_host_index* ix, MoviGpuIndexDesc* out) {
    if (!ix || !out) return fail("null argument");
    *out = ix->d;
    return 0;
}
uint64_t movi_host_index_original_r(const movi_host_index* ix) { return ix ? ix->original_r : 0; }
const char* movi_host_index_type_name(const movi_host_index* ix) { return ix ? ix->type_name.c_str() : ""; }

void movi_host_index_close(movi_host_index* ix) {
    if (!ix) return;
    if (ix->map) munmap(const_cast<uint8_t*>(ix->map), ix->size);
    if (ix->fd >= 0) close(ix->fd);
    delete ix;
}

// BatchLoader::loadBatch + grabNextRead (reference src/batch_loader.cpp:26-143) over the whole input.
// The file is mapped (stdin: slurped) and split with memchr: the reference's getline / stringstream loader parses a few
// hundred MB/s, which is slower than the GPU consumes reads by two orders of magnitude.
namespace {
struct LineCursor {
    const char* p; const char* end;
    bool next(const char*& b, const char*& e) {          // one line without its '\n'; false at end of input
        if (p >= end) return false;
        const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
        b = p; e = nl ? nl : end; p = nl ? nl + 1 : end;
        return true;
    }
    bool peek_is(char c) const { return p < end && *p == c; }
};
inline const char* rtrim_end(const char* b, const char* e) {
    while (e > b && std::isspace(static_cast<unsigned char>(e[-1]))) e--;
    return e;
}
}  // namespace

// The reference opens its read files through zlib (open_kseq, src/utils.cpp:69-80: gzopen), so FASTA / FASTQ may be
// gzip-compressed.  Inflate a gzip stream (all of its members) held in memory.
static bool gunzip_all(const char* in, size_t n, std::string& out) {
    z_stream zs;
    std::memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, 15 + 32) != Z_OK) return false;
    out.clear();
    out.reserve(n * 4);
    std::vector<char> buf(4u << 20);
    zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(in));
    size_t left = n;
    int rc = Z_OK;
    for (;;) {
        if (zs.avail_in == 0) {
            if (left == 0) break;
            const size_t take = std::min<size_t>(left, 1u << 30);
            zs.avail_in = (uInt)take; left -= take;
        }
        zs.next_out = reinterpret_cast<Bytef*>(buf.data()); zs.avail_out = (uInt)buf.size();
        rc = inflate(&zs, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR) { inflateEnd(&zs); return false; }
        out.append(buf.data(), buf.size() - zs.avail_out);
        if (rc == Z_STREAM_END) {                            // next member of a concatenated gzip file, if any
            if (zs.avail_in == 0 && left == 0) break;
            // what follows a complete member is another member only if it starts with the gzip magic; padding or garbage ends the
            // input (gzread stops there too: the reference reads its inputs through gzopen / gzread, src/utils.cpp:69-80)
            if (zs.avail_in < 2 || zs.next_in[0] != 0x1f || zs.next_in[1] != 0x8b) break;
            if (inflateReset(&zs) != Z_OK) { inflateEnd(&zs); return false; }
        } else if (rc == Z_BUF_ERROR && zs.avail_in == 0 && left == 0) break;   // truncated stream: keep what inflated (gzread does)
    }
    inflateEnd(&zs);
    return true;
}

// A read file opened for batch-wise parsing: the file stays mapped (stdin: slurped; gzip: inflated in memory) and every
// movi_host_reader_next() call parses the next run of whole records.
struct movi_host_reader {
    const char* data = nullptr; size_t size = 0;
    std::string slurp, inflated; void* map = nullptr; size_t map_size = 0; int fd = -1;
    bool fastq = false, eof = false; int reverse = 0;
    LineCursor lc{nullptr, nullptr};
    ~movi_host_reader() { if (map) munmap(map, map_size); if (fd >= 0) close(fd); }
};

int movi_host_reader_open(const char* path, int reverse, movi_host_reader** out) {
    if (!path || !out) return fail("null argument");
    *out = nullptr;
    std::unique_ptr<movi_host_reader> r(new movi_host_reader());
    r->reverse = reverse;
    if (std::string(path) == "-") {                      // "-" = stdin, reference src/movi.cpp:106-118
        char buf[1 << 16]; ssize_t n;
        while ((n = ::read(STDIN_FILENO, buf, sizeof buf)) > 0) r->slurp.append(buf, (size_t)n);
        r->data = r->slurp.data(); r->size = r->slurp.size();
    } else {
        r->fd = open(path, O_RDONLY);
        struct stat sb;
        if (r->fd < 0 || fstat(r->fd, &sb) != 0) return fail("The input file " + std::string(path) + " does not exist.");
        r->size = (size_t)sb.st_size;
        if (r->size) {
            void* m = mmap(nullptr, r->size, PROT_READ, MAP_PRIVATE, r->fd, 0);
            if (m == MAP_FAILED) return fail("mmap failed for " + std::string(path));
            r->map = m; r->map_size = r->size;
            madvise(m, r->size, MADV_SEQUENTIAL);
            r->data = static_cast<const char*>(m);
        }
    }
    if (r->size >= 2 && (unsigned char)r->data[0] == 0x1f && (unsigned char)r->data[1] == 0x8b) {
        if (!gunzip_all(r->data, r->size, r->inflated)) return fail("corrupt gzip stream in " + std::string(path));
        r->data = r->inflated.data(); r->size = r->inflated.size();
    }
    if (r->size == 0) r->eof = true;
    else if (r->data[0] == '>') r->fastq = false;
    else if (r->data[0] == '@') r->fastq = true;
    else return fail("unrecognized input query file type - expects FASTA or FASTQ.");
    r->lc = LineCursor{r->data, r->data + r->size};
    *out = r.release();
    return 0;
}

void movi_host_reader_close(movi_host_reader* r) { delete r; }

int movi_host_reader_raw(const movi_host_reader* r, const char** data, uint64_t* size, int* fastq) {
    if (!r || !data || !size) return fail("null argument");
    *data = r->data; *size = r->size;
    if (fastq) *fastq = r->fastq ? 1 : 0;
    return 0;
}

uint64_t movi_host_record_boundary(const char* data, uint64_t size, uint64_t from, int fastq) {
    if (!data || from >= size) return size;
    const char* end = data + size;
    const char* p = data + from;
    if (from > 0 && p[-1] != '\n') {                      // move to the next line start
        p = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
        if (!p) return size;
        p++;
    }
    while (p < end) {
        if (!fastq) { if (*p == '>') return (uint64_t)(p - data); }
        else if (*p == '@') {
            const char* l1 = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
            const char* l2 = l1 ? static_cast<const char*>(memchr(l1 + 1, '\n', (size_t)(end - l1 - 1))) : nullptr;
            if (l2 && l2 + 1 < end && l2[1] == '+') return (uint64_t)(p - data);
            if (!l2 || l2 + 1 >= end) return size;        // not enough lines left for another record
        }
        p = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
        if (!p) return size;
        p++;
    }
    return size;
}

// Whole records until the batch holds max_bases bases (at least one record); *out = NULL once the input is exhausted.
int movi_host_reader_next(movi_host_reader* r, uint64_t max_bases, movi_host_reads** out) {
    return movi_host_reader_next_reads(r, max_bases, 0, 0, out);
}

int movi_host_reader_next_reads(movi_host_reader* r, uint64_t max_bases, uint64_t min_reads, uint64_t cap_bases, movi_host_reads** out) {
    if (!r || !out) return fail("null argument");
    *out = nullptr;
    if (r->eof) return 0;
    const bool fastq = r->fastq;
    LineCursor& lc = r->lc;
    std::unique_ptr<movi_host_reads, PoolReturn> rd(pool_take());
    const size_t left = (size_t)(lc.end - lc.p);
    const size_t expect = (size_t)std::min<uint64_t>(left, max_bases + (max_bases >> 3) + 4096);
    rd->bases.reserve(fastq ? std::min(expect, left / 2 + 64) : expect);
    rd->offsets.reserve(expect / 160 + 16); rd->id_offsets.reserve(expect / 160 + 16);
    const char *b, *e;
    while ((rd->bases.size() < max_bases || rd->offsets.size() - 1 < min_reads) && (cap_bases == 0 || rd->bases.size() < cap_bases)) {
        if (!lc.next(b, e)) { r->eof = true; break; }
        if (b == e) { r->eof = true; break; }            // grabNextRead: an empty line ends the input
        if (fastq ? *b != '@' : *b != '>')
            return fail(std::string("Incorrect ") + (fastq ? "FASTQ entry, it should start with '@'" : "FASTA entry, it should start with '>'"));
        if (e - b <= 2) return fail("header line is missing an id. invalid query cannot be processed.");
        // id: header[1 .. first blank / tab / CR] INCLUDING that character (batch_loader.cpp:117-119)
        const char* w = b + 1;
        while (w < e && *w != ' ' && *w != '\t' && *w != '\r') w++;
        const size_t idl = (w < e) ? (size_t)(w - b) : (size_t)(e - b);    // == find_first_of position, or header size
        rd->ids.append(b + 1, std::min<size_t>(idl, (size_t)(e - b) - 1));
        rd->id_offsets.push_back(rd->ids.size());
        const size_t seq_begin = rd->bases.size();
        auto drop_record = [&]() {                       // a truncated last record is dropped
            rd->bases.resize(seq_begin); rd->ids.resize(rd->id_offsets[rd->id_offsets.size() - 2]); rd->id_offsets.pop_back(); r->eof = true;
        };
        if (fastq) {
            const char *sb_, *se_, *x0, *x1;
            if (!lc.next(sb_, se_)) { drop_record(); break; }
            rd->bases.append(sb_, (size_t)(rtrim_end(sb_, se_) - sb_));
            if (!lc.next(x0, x1) || !lc.next(x0, x1)) { drop_record(); break; }   // '+' line and qualities
        } else {
            while (lc.p < lc.end && !lc.peek_is('>')) {
                lc.next(b, e);
                rd->bases.append(b, (size_t)(rtrim_end(b, e) - b));
            }
        }
        if (r->reverse) std::reverse(rd->bases.begin() + (std::ptrdiff_t)seq_begin, rd->bases.end());   // read_processor.cpp:49-51, movi.cpp:351-352
        rd->offsets.push_back(rd->bases.size());
    }
    if (lc.p >= lc.end) r->eof = true;
    if (rd->offsets.size() == 1) return 0;               // nothing parsed: end of input
    *out = rd.release();
    return 0;
}

int movi_host_reads_load(const char* path, int reverse, movi_host_reads** out) {
    if (!path || !out) return fail("null argument");
    *out = nullptr;
    movi_host_reader* r = nullptr;
    if (movi_host_reader_open(path, reverse, &r)) return 1;
    movi_host_reads* rd = nullptr;
    const int rc = movi_host_reader_next(r, ~0ull, &rd);
    movi_host_reader_close(r);
    if (rc) return rc;
    *out = rd ? rd : new movi_host_reads();              // an empty file is an empty batch
    return 0;
}

// A batch assembled elsewhere (the device-side parser): copies of its four arrays; bases may be NULL when only ids and lengths matter
int movi_host_reads_adopt(const char* bases, const uint64_t* offsets, uint64_t n_reads, const char* ids, const uint64_t* id_offsets,
                          movi_host_reads** out) {
    if (!out || (n_reads && (!offsets || !ids || !id_offsets))) return fail("null argument");
    std::unique_ptr<movi_host_reads> rd(new movi_host_reads());
    if (n_reads) {
        rd->offsets.assign(offsets, offsets + n_reads + 1);
        rd->id_offsets.assign(id_offsets, id_offsets + n_reads + 1);
        if (rd->offsets[0] != 0 || rd->id_offsets[0] != 0) return fail("offsets of an adopted batch must start at 0");
        rd->ids.assign(ids, (size_t)id_offsets[n_reads]);
        if (bases) rd->bases.assign(bases, (size_t)offsets[n_reads]);
    }
    *out = rd.release();
    return 0;
}

// Continue parsing at byte `at` of the input (a record boundary): the device-side ingest hands over here when it meets something
// the reference rejects or stops at
int movi_host_reader_seek(movi_host_reader* r, uint64_t at) {
    if (!r || at > r->size) return fail("seek beyond the input");
    r->lc = LineCursor{r->data + at, r->data + r->size};
    r->eof = (at >= r->size);
    return 0;
}

uint64_t movi_host_reads_count(const movi_host_reads* rd) { return rd ? rd->offsets.size() - 1 : 0; }
const char* movi_host_reads_bases(const movi_host_reads* rd) { return rd ? rd->bases.data() : nullptr; }
const uint64_t* movi_host_reads_offsets(const movi_host_reads* rd) { return rd ? rd->offsets.data() : nullptr; }
const char* movi_host_reads_id(const movi_host_reads* rd, uint64_t i, uint32_t* len) {
    if (!rd || i + 1 >= rd->id_offsets.size()) { if (len) *len = 0; return nullptr; }
    if (len) *len = (uint32_t)(rd->id_offsets[i + 1] - rd->id_offsets[i]);
    return rd->ids.data() + rd->id_offsets[i];
}
const char* movi_host_reads_ids(const movi_host_reads* rd) { return rd ? rd->ids.data() : nullptr; }
const uint64_t* movi_host_reads_id_offsets(const movi_host_reads* rd) { return rd ? rd->id_offsets.data() : nullptr; }
void movi_host_reads_free(movi_host_reads* rd) { pool_give(rd); }

namespace {
struct FdWriter {
    int fd; std::string buf;
    explicit FdWriter(int f) : fd(f) { buf.reserve(1 << 20); }
    bool flush() {
        size_t o = 0;
        while (o < buf.size()) { ssize_t w = ::write(fd, buf.data() + o, buf.size() - o); if (w <= 0) return false; o += (size_t)w; }
        buf.clear(); return true;
    }
    bool put(const char* p, size_t n) { buf.append(p, n); return buf.size() < (1 << 20) || flush(); }
    bool put(const std::string& s) { return put(s.data(), s.size()); }
};
}  // namespace

// Classifier::initialize_report_file :24-32 + EmpNullDatabase::deserialize (emperical_null_database.cpp:106-126)
int movi_host_nulldb_threshold(const char* index_dir, const char* query_type, uint64_t* max_value_thr) {
    if (!index_dir || !query_type || !max_value_thr) return fail("null argument");
    const std::string name = std::string(index_dir) + "/movi." + query_type + ".nulldb";
    std::ifstream f(name, std::ios::binary);
    if (!f) return fail("cannot open null database " + name);
    uint64_t num = 0, percentile = 0; double mean = 0;
    f.read(reinterpret_cast<char*>(&num), 8); f.read(reinterpret_cast<char*>(&mean), 8); f.read(reinterpret_cast<char*>(&percentile), 8);
    if (!f) return fail("truncated null database " + name);
    // max_value_thr is a uint16_t member in the reference (include/classifier.hpp:30)
    *max_value_thr = (uint16_t)(std::max<uint64_t>(percentile, 3) + 1);   // MIN_MATCHING_LENGTH = 3, include/utils.hpp:152
    return 0;
}


// output_base_stats (binary branch) + BPFHeader, reference src/utils.cpp:202-246, include/utils.hpp:64-86
int movi_host_bpf_begin(const char* file) {
    if (!file) { fail("null argument"); return -1; }
    const int fd = open(file, O_RDWR | O_CREAT | O_TRUNC, 0644);      // (read-write: images may be written through a shared mapping)
    if (fd < 0) { fail(std::string("cannot open ") + file); return -1; }
    uint8_t header[12] = {0};
    std::memcpy(header, &BPF_MAGIC, 4);
    header[4] = 1; header[5] = 0; header[6] = 0;         // BPF_VERSION 1.0.0, include/version.hpp:13-15
    header[7] = 16;                                      // entry size in bits
    if (::write(fd, header, 12) != 12) { close(fd); fail(std::string("write failed: ") + file); return -1; }
    return fd;
}

int movi_host_bpf_append(int fd, const movi_host_reads* rd, const uint16_t* ml) {
    if (fd < 0 || !rd || (!ml && !rd->bases.empty())) return fail("null argument");
    FdWriter w(fd);
    bool ok = true;
    const uint64_t n = rd->offsets.size() - 1;
    for (uint64_t i = 0; i < n && ok; i++) {
        const uint16_t idl = (uint16_t)(rd->id_offsets[i + 1] - rd->id_offsets[i]);
        const uint64_t len = rd->offsets[i + 1] - rd->offsets[i];
        ok = w.put(reinterpret_cast<const char*>(&idl), 2) && w.put(rd->ids.data() + rd->id_offsets[i], idl) &&
             w.put(reinterpret_cast<const char*>(&len), 8) && w.put(reinterpret_cast<const char*>(ml + rd->offsets[i]), len * 2);
    }
    ok = ok && w.flush();
    return ok ? 0 : fail("write failed: .bpf");
}

// The records of a batch as one image (assembled on the GPU): positional writes from several threads -- filling the page cache is
// a memory copy, and one thread copies at 2-3 GB/s where the device hands results over at tens of GB/s.
int movi_host_bpf_append_image(int fd, const uint8_t* image, uint64_t bytes) {
    if (fd < 0 || (!image && bytes)) return fail("null argument");
    if (bytes == 0) return 0;
    const off_t at = lseek(fd, 0, SEEK_CUR);
    if (at < 0) return fail("lseek failed: .bpf");
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned n_thr = (unsigned)std::min<uint64_t>(std::min(hw, 8u), (bytes + (8u << 20) - 1) / (8u << 20));
    std::atomic<bool> ok{true};
    auto part = [&](unsigned t) {
        const uint64_t b0 = bytes * t / n_thr, b1 = bytes * (t + 1) / n_thr;
        uint64_t o = b0;
        while (o < b1) {
            const ssize_t w = pwrite(fd, image + o, (size_t)std::min<uint64_t>(b1 - o, 1ull << 30), at + (off_t)o);
            if (w <= 0) { ok = false; return; }
            o += (uint64_t)w;
        }
    };
    if (n_thr <= 1) part(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < n_thr; t++) th.emplace_back(part, t);
        for (auto& t : th) t.join();
    }
    if (!ok || lseek(fd, at + (off_t)bytes, SEEK_SET) < 0) return fail("write failed: .bpf");
    return 0;
}

int movi_host_bpf_end(int fd) { return (fd >= 0 && close(fd) == 0) ? 0 : fail("close failed: .bpf"); }

int movi_host_write_bpf(const char* file, const movi_host_reads* rd, const uint16_t* ml) {
    if (!file || !rd || (!ml && !rd->bases.empty())) return fail("null argument");
    const int fd = movi_host_bpf_begin(file);
    if (fd < 0) return 1;
    const int rc = movi_host_bpf_append(fd, rd, ml);
    return movi_host_bpf_end(fd) | rc;
}

// output_base_stats (--stdout branch), reference src/utils.cpp:204-209 with MoveQuery::add_ml's string
// (include/move_query.hpp:33-37): values in read order, each followed by one blank
int movi_host_write_ml_text(int fd, const movi_host_reads* rd, const uint16_t* ml) {
    if (!rd || (!ml && !rd->bases.empty())) return fail("null argument");
    FdWriter w(fd);
    const uint64_t n = rd->offsets.size() - 1;
    char num[8];
    for (uint64_t i = 0; i < n; i++) {
        w.put(">", 1);
        w.put(rd->ids.data() + rd->id_offsets[i], rd->id_offsets[i + 1] - rd->id_offsets[i]);
        w.put("\n", 1);
        const uint64_t s = rd->offsets[i], e = rd->offsets[i + 1];
        for (uint64_t k = e; k > s; k--) {
            unsigned v = ml[k - 1];
            int l = 0; char tmp[6];
            do { tmp[l++] = (char)('0' + v % 10); v /= 10; } while (v);
            for (int j = 0; j < l; j++) num[j] = tmp[l - 1 - j];
            num[l] = ' ';
            w.put(num, (size_t)l + 1);
        }
        w.put("\n", 1);
    }
    return w.flush() ? 0 : fail("write failed");
}

// output_counts, reference src/utils.cpp:248-256
int movi_host_write_counts(int fd, const movi_host_reads* rd, const uint64_t* counts, const int32_t* pos_on_r) {
    if (!rd || !counts || !pos_on_r) return fail("null argument");
    FdWriter w(fd);
    const uint64_t n = rd->offsets.size() - 1;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t len = rd->offsets[i + 1] - rd->offsets[i];
        w.put(rd->ids.data() + rd->id_offsets[i], rd->id_offsets[i + 1] - rd->id_offsets[i]);
        w.put("\t" + std::to_string(len - (uint64_t)(int64_t)pos_on_r[i]) + "/" + std::to_string(len) + "\t" + std::to_string(counts[i]) + "\n");
    }
    return w.flush() ? 0 : fail("write failed");
}

// Classifier::initialize_report_file (header) + classify (rows), reference src/classifier.cpp:49-60, 99-143;
// filter output = output_read, src/utils.cpp:291-294 via src/movi.cpp:41-47.
// The three report columns come either from the matching lengths (ml != NULL: binned here, classifier.cpp:108-126) or from
// the device (movi_gpu_query_classify: above / below / sum arrays).
static int classify_impl(int fd, int mode, const movi_host_reads* rd, const uint16_t* ml, uint64_t bin_width, uint64_t max_value_thr,
                         const uint32_t* d_above, const uint32_t* d_below, const uint64_t* d_sum, uint8_t* found_out) {
    FdWriter w(fd);
    std::ostringstream out;
    const uint16_t thr = (uint16_t)max_value_thr;
    const bool continuation = (mode & MOVI_HOST_CLASSIFY_CONTINUE) != 0;      // a later batch of the same report: no header
    mode &= ~MOVI_HOST_CLASSIFY_CONTINUE;
    out.precision(4);
    if (mode == 0 && !continuation) {
        out << std::setw(30) << std::left << "read id:" << std::setw(15) << std::left << "status:"
            << std::setw(19) << std::left << "avg max-value (thr=" << std::setw(2) << std::left << thr
            << std::setw(5) << std::left << "):" << std::setw(12) << std::left << "above thr:"
            << std::setw(12) << std::left << "below thr:" << std::endl;
    }
    const uint64_t n = rd->offsets.size() - 1;
    for (uint64_t i = 0; i < n; i++) {
        const size_t len = rd->offsets[i + 1] - rd->offsets[i];
        size_t above = 0, below = 0, sum = 0, nbins = 0;
        if (ml) {
            const uint16_t* v = ml + rd->offsets[i];
            size_t start = 0, end = 0;
            while (start < len) {
                end = (start + bin_width < len) ? start + bin_width : len;
                if (len - end < bin_width) end = len;
                const uint16_t mx = *std::max_element(v + start, v + end);
                if (mx >= thr) above++; else below++;
                sum += mx; nbins++;
                start += (end - start);
            }
        } else {
            above = d_above[i]; below = d_below[i]; sum = d_sum[i]; nbins = above + below;
        }
        const bool found = (above / (above + below + 0.0) > 0.50);
        if (found_out) found_out[i] = found ? 1 : 0;
        const std::string id(rd->ids.data() + rd->id_offsets[i], rd->id_offsets[i + 1] - rd->id_offsets[i]);
        if (mode == 0) {
            out.precision(3);
            out << std::setw(30) << std::left << id << std::setw(15) << std::left << (found ? "FOUND" : "NOT_PRESENT")
                << std::setw(26) << std::left << (sum + 0.0) / nbins << std::setw(12) << std::left << above
                << std::setw(12) << std::left << below << "\n";
        } else if ((mode == 1 && found) || (mode == 2 && !found)) {
            out << ">" << id << "\n";
            out.write(rd->bases.data() + rd->offsets[i], (std::streamsize)len);
            out << "\n";
        }
        if (out.tellp() > (1 << 20)) { w.put(out.str()); out.str(""); }
    }
    w.put(out.str());
    return w.flush() ? 0 : fail("write failed");
}

int movi_host_classify(int fd, int mode, const movi_host_reads* rd, const uint16_t* ml, uint64_t bin_width,
                       uint64_t max_value_thr, uint8_t* found_out) {
    if (!rd || (!ml && !rd->bases.empty()) || bin_width == 0) return fail("bad argument");
    static const uint16_t none = 0;
    return classify_impl(fd, mode, rd, ml ? ml : &none, bin_width, max_value_thr, nullptr, nullptr, nullptr, found_out);
}

int movi_host_classify_bins(int fd, int mode, const movi_host_reads* rd, uint64_t max_value_thr, const uint32_t* bins_above,
                            const uint32_t* bins_below, const uint64_t* sum_max, uint8_t* found_out) {
    if (!rd || ((!bins_above || !bins_below || !sum_max) && rd->offsets.size() > 1)) return fail("bad argument");
    return classify_impl(fd, mode, rd, nullptr, 1, max_value_thr, bins_above, bins_below, sum_max, found_out);
}

// ---------------------------------------------------------------- Movi Color: colour table, sampled suffix array, their outputs
struct movi_host_colors {
    std::vector<uint16_t> flat;
    std::vector<uint8_t> inds;                 // r x 5 bytes
    std::vector<uint32_t> to_taxon_id;
    std::vector<double> log_lens;
    uint32_t num_docs = 0;
};
struct movi_host_ssa { uint64_t rate = 0; std::vector<uint64_t> sampled, all_p; };

// load_color_table -> deserialize_doc_sets_flat + load_document_info (reference src/movi.cpp:121-152,
// src/move_structure_io.cpp:587-607, 643-700)
int movi_host_colors_open(const char* index_dir, uint64_t r, movi_host_colors** out) { return movi_host_colors_open_kind(index_dir, r, 0, out); }

// kind 0: doc_sets_flat.bin (the default); kind 1: doc_sets.bin, the vectors-of-vectors form of `--color-vectors`
// (serialize_doc_sets / deserialize_doc_sets, src/move_structure_io.cpp:571-641), brought into the flat form here the way
// flat_and_serialize_colors_vectors does (:513-548) -- a row whose set index lies beyond the stored sets ("thrown away by
// compression", src/move_structure_query.cpp:390-391) gets an offset beyond the table, which the query skips the same way (:395)
int movi_host_colors_open_kind(const char* index_dir, uint64_t r, int kind, movi_host_colors** out) {
    if (!index_dir || !out) return fail("null argument");
    const std::string dir(index_dir);
    std::unique_ptr<movi_host_colors> c(new movi_host_colors);
    if (kind == 0) {
        const std::string name = dir + "/doc_sets_flat.bin";
        std::ifstream f(name, std::ios::binary);
        if (!f) return fail("[deserialize doc sets flat] Failed to open document sets flat file at " + name);
        f.seekg(0, std::ios::end); const uint64_t size = (uint64_t)f.tellg(); f.seekg(0);
        uint64_t n = 0;
        f.read(reinterpret_cast<char*>(&n), 8);
        if (!f || n > size / 2 || r > size / 5 || 8 + n * 2 + r * 5 > size) return fail("truncated colour table " + name);
        c->flat.resize(n); c->inds.resize(r * 5);
        f.read(reinterpret_cast<char*>(c->flat.data()), (std::streamsize)(n * 2));
        f.read(reinterpret_cast<char*>(c->inds.data()), (std::streamsize)(r * 5));
        if (!f) return fail("truncated colour table " + name);
    } else if (kind == 1) {
        const std::string name = dir + "/doc_sets.bin";
        std::ifstream f(name, std::ios::binary);
        if (!f) return fail("[deserialize doc sets] Failed to open document sets file at " + name);
        f.seekg(0, std::ios::end); const uint64_t size = (uint64_t)f.tellg(); f.seekg(0);
        uint64_t n_sets = 0;
        f.read(reinterpret_cast<char*>(&n_sets), 8);
        if (!f || n_sets > size / 2) return fail("truncated colour table " + name);
        std::vector<uint64_t> flat_index(n_sets);
        std::vector<uint16_t> docs;
        for (uint64_t i = 0; i < n_sets; i++) {
            uint16_t k = 0;
            f.read(reinterpret_cast<char*>(&k), 2);
            docs.resize(k);
            if (k) f.read(reinterpret_cast<char*>(docs.data()), (std::streamsize)k * 2);
            if (!f) return fail("truncated colour table " + name);
            flat_index[i] = c->flat.size();
            c->flat.push_back(k);
            c->flat.insert(c->flat.end(), docs.begin(), docs.end());
        }
        std::vector<uint32_t> set_of(r);
        f.read(reinterpret_cast<char*>(set_of.data()), (std::streamsize)(r * 4));
        if (!f) return fail("truncated colour table " + name);
        if (c->flat.size() >= (1ull << 40) - 1) return fail("colour table too large for 40-bit offsets");
        c->inds.resize(r * 5);
        for (uint64_t i = 0; i < r; i++) {
            const uint64_t off = set_of[i] < n_sets ? flat_index[set_of[i]] : (uint64_t)c->flat.size();
            uint8_t* e = c->inds.data() + 5 * i;
            e[0] = (uint8_t)off; e[1] = (uint8_t)(off >> 8); e[2] = (uint8_t)(off >> 16); e[3] = (uint8_t)(off >> 24); e[4] = (uint8_t)(off >> 32);
        }
    } else return fail("unknown colour table kind");
    std::vector<uint64_t> doc_offsets;
    {
        const std::string name = dir + "/ref.fa.doc_offsets";
        std::ifstream f(name);
        if (!f) return fail("[load document info] doc_offsets file not found at \"" + name + "\"");
        uint64_t v; while (f >> v) doc_offsets.push_back(v);
    }
    c->num_docs = (uint32_t)doc_offsets.size();
    std::vector<uint32_t> doc_ids;
    {
        std::ifstream f(dir + "/ref.fa.doc_ids");
        if (f) { uint32_t v; while (f >> v) doc_ids.push_back(v); }
        else for (uint32_t i = 0; i < c->num_docs; i++) doc_ids.push_back(i + 1);       // "No doc ids information, setting doc ids to 1..."
    }
    // taxon ids compressed to 0 .. num_species - 1 in ascending order (a std::map in the reference)
    c->to_taxon_id = doc_ids;
    std::sort(c->to_taxon_id.begin(), c->to_taxon_id.end());
    c->to_taxon_id.erase(std::unique(c->to_taxon_id.begin(), c->to_taxon_id.end()), c->to_taxon_id.end());
    if (c->to_taxon_id.empty() || c->to_taxon_id.size() >= 0xFFFFu) return fail("the colour table needs between 1 and 65534 species");
    c->log_lens.assign(c->to_taxon_id.size(), 0.0);
    if (doc_ids.size() < doc_offsets.size()) return fail("ref.fa.doc_ids names fewer documents than ref.fa.doc_offsets");
    for (size_t i = 0; i < doc_offsets.size(); i++) {
        const size_t sp = (size_t)(std::lower_bound(c->to_taxon_id.begin(), c->to_taxon_id.end(), doc_ids[i]) - c->to_taxon_id.begin());
        c->log_lens[sp] += (double)(doc_offsets[i] - (i == 0 ? 0 : doc_offsets[i - 1]));
    }
    for (double& v : c->log_lens) v = std::log(v);
    *out = c.release();
    return 0;
}
int movi_host_colors_desc(const movi_host_colors* c, MoviGpuColorDesc* out) {
    if (!c || !out) return fail("null argument");
    std::memset(out, 0, sizeof *out);
    out->flat_colors = c->flat.data(); out->flat_size = c->flat.size(); out->doc_set_flat_inds = c->inds.data();
    out->num_species = (uint32_t)c->to_taxon_id.size(); out->log_lens = c->log_lens.data();
    return 0;
}
uint32_t movi_host_colors_num_species(const movi_host_colors* c) { return c ? (uint32_t)c->to_taxon_id.size() : 0; }
uint32_t movi_host_colors_taxon(const movi_host_colors* c, uint32_t species) { return c && species < c->to_taxon_id.size() ? c->to_taxon_id[species] : 0; }
void movi_host_colors_close(movi_host_colors* c) { delete c; }

// the end of query_pml under is_multi_classify(): "<id>,<taxon>,<taxon>" (reference src/move_structure_query.cpp:439-471)
int movi_host_write_multi_classify(int fd, const movi_host_reads* rd, const movi_host_colors* c, const MoviGpuColorResult* res,
                                   uint32_t min_len, int pvalue) {
    if (!rd || !c || (!res && rd->offsets.size() > 1)) return fail("null argument");
    FdWriter w(fd);
    const uint64_t n = rd->offsets.size() - 1;
    for (uint64_t i = 0; i < n; i++) {
        const MoviGpuColorResult& q = res[i];
        const uint64_t len = rd->offsets[i + 1] - rd->offsets[i];
        w.put(rd->ids.data() + rd->id_offsets[i], rd->id_offsets[i + 1] - rd->id_offsets[i]);
        const float pml_mean = static_cast<float>(q.sum_matching_lengths) / len;                 // :443
        const bool no_best = q.best_doc >= c->to_taxon_id.size();
        if (pml_mean < 0.4 || no_best) { w.put(",0,0\n"); continue; }                            // UNCLASSIFIED_THRESHOLD, include/utils.hpp:169
        std::string line = "," + std::to_string(c->to_taxon_id[q.best_doc]);
        if (q.second_best_doc >= c->to_taxon_id.size()) line += ",0";
        else {
            // the comparison reads classify_cnts when --min-len > 0 and doc_scores otherwise (:452-461), whichever of the two the
            // counting loop filled (:405-433): the other one is all zeros
            float best_cnt = 0, second_cnt = 0;
            if (min_len > 0 && !pvalue) { best_cnt = (float)(uint32_t)q.best_raw; second_cnt = (float)(uint32_t)q.second_raw; }
            else if (min_len == 0 && pvalue) {
                double b, s2; std::memcpy(&b, &q.best_raw, 8); std::memcpy(&s2, &q.second_raw, 8);
                best_cnt = (float)b; second_cnt = (float)s2;
            }
            const float diff = best_cnt - second_cnt;
            line += diff < 0.05 * best_cnt ? "," + std::to_string(c->to_taxon_id[q.second_best_doc]) : std::string(",0");
        }
        w.put(line + "\n");
    }
    return w.flush() ? 0 : fail("write failed");
}

// deserialize_sampled_SA (reference src/move_structure_io.cpp:716-735)
int movi_host_ssa_open(const char* index_dir, movi_host_ssa** out) {
    if (!index_dir || !out) return fail("null argument");
    const std::string name = std::string(index_dir) + "/ssa.movi";
    std::ifstream f(name, std::ios::binary);
    if (!f) return fail("[deserialize sampled SA] Failed to open sampled SA entries file at " + name + "\nBuild the sampled SA by running the build-SA command.");
    f.seekg(0, std::ios::end); const uint64_t size = (uint64_t)f.tellg(); f.seekg(0);
    std::unique_ptr<movi_host_ssa> s(new movi_host_ssa);
    uint64_t k = 0, r = 0;
    f.read(reinterpret_cast<char*>(&s->rate), 8); f.read(reinterpret_cast<char*>(&k), 8);
    if (!f || k > size / 8) return fail("truncated " + name);
    s->sampled.resize(k);
    f.read(reinterpret_cast<char*>(s->sampled.data()), (std::streamsize)(k * 8));
    f.read(reinterpret_cast<char*>(&r), 8);
    if (!f || r > size / 8) return fail("truncated " + name);
    s->all_p.resize(r);
    f.read(reinterpret_cast<char*>(s->all_p.data()), (std::streamsize)(r * 8));
    if (!f) return fail("truncated " + name);
    *out = s.release();
    return 0;
}
uint64_t movi_host_ssa_rate(const movi_host_ssa* s) { return s ? s->rate : 0; }
uint64_t movi_host_ssa_rows(const movi_host_ssa* s) { return s ? s->all_p.size() : 0; }
uint64_t movi_host_ssa_samples(const movi_host_ssa* s) { return s ? s->sampled.size() : 0; }
const uint64_t* movi_host_ssa_sampled(const movi_host_ssa* s) { return s ? s->sampled.data() : nullptr; }
const uint64_t* movi_host_ssa_all_p(const movi_host_ssa* s) { return s ? s->all_p.data() : nullptr; }
void movi_host_ssa_close(movi_host_ssa* s) { delete s; }

// <prefix>.sa_entries.bpf: the records of output_base_stats(DataType::sa_entry) -- {u16 id_len; id; u64 L; u64 sa[L]} -- and NO
// file header (the reference opens that file without writing one, src/utils.cpp:371-375)
int movi_host_sa_append(int fd, const movi_host_reads* rd, const uint64_t* sa) {
    if (fd < 0 || !rd || (!sa && !rd->bases.empty())) return fail("null argument");
    FdWriter w(fd);
    bool ok = true;
    const uint64_t n = rd->offsets.size() - 1;
    for (uint64_t i = 0; i < n && ok; i++) {
        const uint16_t idl = (uint16_t)(rd->id_offsets[i + 1] - rd->id_offsets[i]);
        const uint64_t len = rd->offsets[i + 1] - rd->offsets[i];
        ok = w.put(reinterpret_cast<const char*>(&idl), 2) && w.put(rd->ids.data() + rd->id_offsets[i], idl) &&
             w.put(reinterpret_cast<const char*>(&len), 8) && w.put(reinterpret_cast<const char*>(sa + rd->offsets[i]), len * 8);
    }
    ok = ok && w.flush();
    return ok ? 0 : fail("write failed: .sa_entries.bpf");
}

}  // extern "C"
