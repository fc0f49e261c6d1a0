// movi_cli.cpp -- `movi-b200 query ...`: the reference's `movi query` command line over the GPU query path.
//
// Mirrors the query action of the reference (flags: src/movi_parser.cpp:132-162; driver: query() in
// src/movi.cpp:221-400; output routing: include/movi_options.hpp:70-71, src/utils.cpp:319-384,
// src/classifier.cpp:24-63).  The read file streams through in batches (--chunk-bases; the next batch is parsed while the
// GPUs walk the current one); each batch is sharded over --gpus GPUs (index replicated, one host thread + one context per
// GPU) and its results are appended to the outputs in read order (= reference --no-prefetch -t1).
// `movi-b200 view --bpf F` prints a .bpf file like the reference's view action (src/movi.cpp:402-503).
#include <fcntl.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../../include/movi_gpu.h"
#include "../../include/movi_host.h"

namespace {

struct Options {
    std::string action, index_dir, read_file, out_file, bpf_file;
    std::string query = "pml";
    bool rpml = false;            // --rpml: PML with random repositioning; names its outputs "rpml" (src/utils.cpp:47-52)
    bool classify = false, filter = false, invert = false, to_stdout = false, no_output = false, reverse = false, verbose = false;
    int ignore_illegal = 0, gpus = 1, device = 0, threads = 1, strands = 16;
    uint64_t bin_width = 150;     // include/movi_options.hpp default
    uint64_t chunk_bases = 1ull << 26;   // bases per batch of the streaming loop (host memory: ~10 bytes per base of a batch)
    uint64_t chunk_reads = 0;            // reads per batch at least: 0 = 2.3e5 per GPU (one read per lane in flight, however long the reads)
    uint64_t max_chunk_bases = 1ull << 31;   // ... but never more bases than this per batch (pinned results: 2 x 2 bytes per base)
    uint64_t chunk_bytes = 256ull << 20;     // device-side ingest: file bytes per chunk (grows by itself when reads are long)
    bool host_parse = false;                 // --host-parse: never use the device-side ingest
    bool device_ingest = false;              // --device-ingest: accepted (the device-side ingest is the default wherever it exists)
    int layout = MOVI_GPU_LAYOUT_AUTO;       // --layout auto|expanded|compact: device form of blocked / sampled indexes
    // Movi Color (movi_parser.cpp:155, 164-175)
    bool multi_classify = false, pvalue_scoring = false, sa_entries = false, color_vectors = false;
    uint32_t min_len = 1;                    // --min-len (uint8 in the reference, movi_options.hpp:291)
    bool write_output_allowed() const { return !no_output && !filter; }          // movi_options.hpp:70
    bool write_stdout_enabled() const { return to_stdout && !classify; }         // movi_options.hpp:71
};

[[noreturn]] void usage_error(const std::string& m) {
    fprintf(stderr, "\033[31m\n\n[ERROR] %s\n\033[0m\n", m.c_str());
    fprintf(stderr, "usage: movi-b200 query -i <index dir> -r <reads.fa|fq|-> [--pml [--rpml]|--zml|--count] [--classify|--filter [-v]]\n"
                    "                       [--stdout] [-o prefix] [--no-output] [--reverse] [--ignore-illegal-chars 1|2]\n"
                    "                       [--bin-width N] [--gpus N] [--device D] [--chunk-bases N] [--chunk-reads N] [--max-chunk-bases N]\n"
                    "                       [--chunk-bytes N] [--host-parse|--device-ingest] [--layout auto|expanded|compact]\n"
                    "                       [--multi-classify -o out.csv [--color-vectors] [--min-len N] [--pvalue-scoring]] [--sa-entries]\n"
                    "                       [-t N] [-s N] [-n]\n"
                    "       movi-b200 view --bpf <file.bpf>\n");
    exit(1);
}

// cxxopts semantics used by the reference's tests: "-t1", "-t 1", "--threads 1", "--threads=1"
bool take_value(int argc, char** argv, int& i, const char* shortf, const char* longf, std::string& out) {
    const std::string a = argv[i];
    if (shortf && a.rfind(shortf, 0) == 0 && a.size() > 2 && a[1] != '-') { out = a.substr(2); return true; }
    if ((shortf && a == shortf) || a == longf) {
        if (i + 1 >= argc) usage_error("missing value for " + a);
        out = argv[++i]; return true;
    }
    const std::string eq = std::string(longf) + "=";
    if (a.rfind(eq, 0) == 0) { out = a.substr(eq.size()); return true; }
    return false;
}

Options parse(int argc, char** argv) {
    Options o;
    if (argc < 2) usage_error("no action given");
    o.action = argv[1];
    std::string v;
    for (int i = 2; i < argc; i++) {
        const std::string a = argv[i];
        if (take_value(argc, argv, i, "-i", "--index", v)) o.index_dir = v;
        else if (take_value(argc, argv, i, "-r", "--read", v)) o.read_file = v;
        else if (take_value(argc, argv, i, "-o", "--out-file", v)) o.out_file = v;
        else if (take_value(argc, argv, i, "-t", "--threads", v)) o.threads = atoi(v.c_str());
        else if (take_value(argc, argv, i, "-s", "--strands", v)) o.strands = atoi(v.c_str());
        else if (take_value(argc, argv, i, nullptr, "--bin-width", v)) o.bin_width = strtoull(v.c_str(), nullptr, 10);
        else if (take_value(argc, argv, i, nullptr, "--ignore-illegal-chars", v)) o.ignore_illegal = atoi(v.c_str());
        else if (take_value(argc, argv, i, nullptr, "--gpus", v)) o.gpus = atoi(v.c_str());
        else if (take_value(argc, argv, i, nullptr, "--device", v)) o.device = atoi(v.c_str());
        else if (take_value(argc, argv, i, nullptr, "--bpf", v)) o.bpf_file = v;
        // ftab seeds a backward search with the interval of its first k bases (src/move_structure_search.cpp:203-293): an
        // accelerator with identical results (the stepwise search reaches the same interval and emits the same 0..k-2), so
        // the flags are accepted and the table is not needed
        else if (take_value(argc, argv, i, nullptr, "--ftab-k", v)) { }
        else if (a == "--multi-ftab") { }
        else if (take_value(argc, argv, i, nullptr, "--chunk-bases", v)) o.chunk_bases = std::max<uint64_t>(1, strtoull(v.c_str(), nullptr, 10));
        else if (take_value(argc, argv, i, nullptr, "--chunk-reads", v)) o.chunk_reads = strtoull(v.c_str(), nullptr, 10);
        else if (take_value(argc, argv, i, nullptr, "--chunk-bytes", v)) o.chunk_bytes = std::max<uint64_t>(4096, strtoull(v.c_str(), nullptr, 10));
        else if (a == "--host-parse") o.host_parse = true;
        else if (a == "--device-ingest") o.device_ingest = true;
        else if (take_value(argc, argv, i, nullptr, "--layout", v)) {
            if (v == "auto") o.layout = MOVI_GPU_LAYOUT_AUTO; else if (v == "expanded") o.layout = MOVI_GPU_LAYOUT_EXPANDED8;
            else if (v == "compact") o.layout = MOVI_GPU_LAYOUT_COMPACT; else usage_error("--layout: auto, expanded or compact");
        }
        else if (take_value(argc, argv, i, nullptr, "--max-chunk-bases", v)) o.max_chunk_bases = std::max<uint64_t>(1, strtoull(v.c_str(), nullptr, 10));
        else if (a == "--multi-classify") o.multi_classify = true;
        else if (a == "--pvalue-scoring") o.pvalue_scoring = true;
        else if (a == "--color-vectors") o.color_vectors = true;                // doc_sets.bin instead of doc_sets_flat.bin (movi_parser.cpp:369)
        else if (a == "--sa-entries") o.sa_entries = true;
        else if (take_value(argc, argv, i, nullptr, "--min-len", v)) o.min_len = (uint32_t)(uint8_t)strtoul(v.c_str(), nullptr, 10);
        else if (a == "--early-stop" || a == "--report-all" || a == "--report-colors" || a == "--report-color-ids" || a == "--full" ||
                 a == "--compress" || a == "--freq-compress" || a == "--tree-compress" || a == "--color-move-rows" ||
                 a == "--min-diff-frac" || a == "--min-score-frac")
            usage_error(a + ": only the flat and the vectors-of-vectors colour tables with the output of the --no-prefetch path are built here "
                            "(--multi-classify [--color-vectors] [--min-len N] [--pvalue-scoring]); use the reference for " + a);
        else if (a == "--mem" || a == "--kmer" || a == "--kmer-count" || a == "-l" || a == "--min-mem-length" || a == "-k" || a == "--k-length")
            usage_error(a + ": the MEM and k-mer queries are outside this build (PML / ZML / count, classification, Movi Color): use the reference");
        else if (a == "--pml") o.query = "pml";
        else if (a == "--rpml") o.rpml = true;                                  // movi_parser.cpp:149,357
        else if (a == "--zml") o.query = "zml";
        else if (a == "--count") o.query = "count";
        else if (a == "--classify") o.classify = true;
        else if (a == "--filter") { o.filter = true; o.classify = true; }      // movi_options.hpp:122-125
        else if (a == "-v" || a == "--invert") o.invert = true;
        else if (a == "--stdout") o.to_stdout = true;
        else if (a == "--no-output") o.no_output = true;
        else if (a == "--reverse") o.reverse = true;
        else if (a == "-n" || a == "--no-prefetch") { /* latency hiding is the GPU's job here */ }
        else if (a == "--verbose") o.verbose = true;
        else if (a == "--separators" || a == "--mmap") { /* detected from the index / not applicable */ }
        else usage_error("unknown option " + a);
    }
    return o;
}

int view_bpf(const std::string& file) {   // reference src/movi.cpp:402-503 (16-bit entries)
    FILE* f = fopen(file.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", file.c_str()); return 1; }
    uint8_t h[12];
    uint32_t magic = 0;
    if (fread(h, 1, 12, f) != 12 || (memcpy(&magic, h, 4), magic != 0x42504600u) || h[7] != 16) {
        fprintf(stderr, "not a 16-bit BPF file: %s\n", file.c_str()); fclose(f); return 1;
    }
    for (;;) {
        uint16_t idl;
        if (fread(&idl, 2, 1, f) != 1) break;
        std::string id(idl, '\0');
        uint64_t n = 0;
        if (fread(&id[0], 1, idl, f) != idl || fread(&n, 8, 1, f) != 1) break;
        std::vector<uint16_t> v(n);
        if (n && fread(v.data(), 2, n, f) != n) break;
        printf(">%s\n", id.c_str());
        for (uint64_t k = n; k > 0; k--) printf("%u ", (unsigned)v[k - 1]);
        printf("\n");
    }
    fclose(f);
    return 0;
}

struct ShardResult { int rc = 0; std::string err; };
struct Fetched { int rc = 0; movi_host_reads* rd = nullptr; std::string err; };

// results of one batch in pinned host memory (grown on demand, reused by every second batch)
template <typename T> struct PinnedBuf {
    T* p = nullptr; size_t cap = 0;
    bool need(size_t n) {
        if (n <= cap && p) return true;
        if (p) movi_gpu_host_free(p);
        p = nullptr; cap = 0;
        const size_t want = n + n / 8 + 4096;
        void* q = nullptr;
        if (movi_gpu_host_alloc(&q, want * sizeof(T))) return false;
        p = static_cast<T*>(q); cap = want;
        return true;
    }
    ~PinnedBuf() { if (p) movi_gpu_host_free(p); }
};
struct Results {
    movi_host_reads* rd = nullptr;
    PinnedBuf<uint16_t> ml; PinnedBuf<uint64_t> counts; PinnedBuf<int32_t> pos;
    PinnedBuf<uint32_t> above, below; PinnedBuf<uint64_t> sum;
    PinnedBuf<uint8_t> image; uint64_t image_bytes = 0;     // the batch's .bpf records, assembled on the GPUs
};

// ---- device-side ingest path: file bytes -> GPU (parse, walk, .bpf records / classification) -> file.  The host only finds record
// boundaries near the chunk ends; it copies nothing into buffers of its own and formats nothing but report lines.  Chunk k goes to
// GPU k mod G (one host thread per context): the chunk is parsed there, which tells the size of its records, so the chunks' places
// in the output file are handed out in chunk order while the walks and the writes (positional, straight from the copy-back ring of
// the library: movi_gpu_file_chunk_query) run side by side.  With --classify / --filter the chunk is binned on the device
// (movi_gpu_file_chunk_classify), 16 bytes per read and the ids (--filter: the sequences too) come back, and the report rows are
// written in chunk order.  Returns 0 when the whole input went through; 1 when a chunk holds something the reference rejects or
// stops at: every chunk before it is complete in the outputs, and the caller continues from `end->resume_at` with the host parser,
// which reproduces the reference's behaviour from there (the chunks before it ended exactly where it starts, so the host parser
// sees what a parse of the whole file would see at that point); -1 on errors (message in `err`).
struct IngestStats { uint64_t reads = 0, bases = 0, chunks = 0; };
struct IngestSinks {
    int fd_bpf = -1;                          // <prefix>.bpf (or -1)
    bool classify = false;                    // --classify / --filter
    int report_fd = -1, report_mode = 0;      // movi_host_classify_bins: 0 report rows, 1 --filter, 2 --filter --invert
    uint32_t bin_width = 150, max_value_thr = 0;
};
struct IngestEnd { uint64_t resume_at = 0, file_off = 12; bool report_started = false; };

int run_device_ingest(const Options& o, const std::vector<movi_gpu_ctx*>& ctx, const char* data, uint64_t size, int fastq, bool zml,
                      uint32_t qflags, const IngestSinks& sinks, IngestStats* stats, IngestEnd* fin, std::string* err) {
    const int G = (int)ctx.size();
    const int fd_bpf = sinks.fd_bpf;
    constexpr uint64_t NONE = ~0ull;
    std::mutex mu;
    std::condition_variable cv;
    uint64_t next_begin = 0, next_seq = 0, place_seq = 0, report_seq = 0, file_off = 12, chunk_bytes = o.chunk_bytes;     // 12: the BPFHeader
    uint64_t fail_seq = NONE, fail_begin = 0;      // first chunk (in input order) the device parser does not take (under mu)
    bool hard = false, report_started = false;     // hard: an error (message in *err)
    // long reads: a chunk should hold a read per lane from the start (a chunk of few long reads costs a full walk however small
    // it is, and runs the plain-fetch kernel): judge by the size of the first record
    if (size > 1) {
        const uint64_t rec = movi_host_record_boundary(data, size, 1, fastq);
        if (rec > 4096) chunk_bytes = std::max<uint64_t>(chunk_bytes, std::min<uint64_t>(rec * 230000ull, 3ull << 30));
    }

    // inputs of several chunks: a second context per GPU (sharing the uploaded index) so that the copy-in and parse of one chunk run
    // under the walk and write-back of another
    std::vector<movi_gpu_ctx*> lanes2(G, nullptr);
    const int per_gpu = size > 2 * chunk_bytes ? 2 : 1;
    if (per_gpu == 2) {
        std::vector<std::thread> mk;
        for (int g = 0; g < G; g++) mk.emplace_back([&, g]() {
            movi_gpu_ctx* c = nullptr;
            if (movi_gpu_create(o.device + g, &c) == MOVI_GPU_OK && movi_gpu_share_index(c, ctx[g]) == MOVI_GPU_OK) lanes2[g] = c;
            else if (c) movi_gpu_destroy(c);
        });
        for (auto& t : mk) t.join();
    }
    auto worker = [&](int g, movi_gpu_ctx* my) {
        (void)g;
        std::vector<uint32_t> above, below; std::vector<uint64_t> sum, offs, id_offs; std::vector<char> ids;
        for (;;) {
            uint64_t begin, end, seq;
            {
                std::lock_guard<std::mutex> lk(mu);
                if (hard || fail_seq != NONE || next_begin >= size) return;
                begin = next_begin;
                end = begin + chunk_bytes >= size ? size : movi_host_record_boundary(data, size, begin + chunk_bytes, fastq);
                next_begin = end; seq = next_seq++;
            }
            auto soft_fail = [&]() {               // this chunk goes to the host parser, and with it everything behind it
                std::lock_guard<std::mutex> lk(mu);
                if (seq < fail_seq) { fail_seq = seq; fail_begin = begin; }
                cv.notify_all();
            };
            auto hard_fail = [&](const std::string& m) {
                std::lock_guard<std::mutex> lk(mu);
                if (!hard) { *err = m; hard = true; }
                cv.notify_all();
            };
            if (end - begin >= 0xFFFFFF00ull) { soft_fail(); return; }     // one record of 4 GiB or more: not a chunk the device parser takes
            MoviGpuChunkResult res{};
            int rc = movi_gpu_file_chunk_parse(my, data + begin, end - begin, fastq, 1, &res);
            if (rc) { hard_fail(movi_gpu_last_error(my)); return; }
            if (res.parse_flags || res.consumed_bytes != end - begin) { soft_fail(); return; }
            uint64_t my_off = 0;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return place_seq == seq || hard || seq > fail_seq; });      // places in the file go out in chunk order
                if (hard || seq > fail_seq) return;
                my_off = file_off; file_off += res.image_bytes; place_seq += 1;
                if (fd_bpf >= 0 && res.image_bytes && ftruncate(fd_bpf, (off_t)file_off) != 0) { /* the library falls back to pwrite */ }
                stats->reads += res.n_reads; stats->bases += res.n_bases; stats->chunks += 1;
                // a chunk should hold a read per lane of its GPU: long reads get bigger chunks (the table rows are random gathers,
                // only the number of walks in flight hides their latency)
                if (res.n_reads < 230000 && chunk_bytes < (3ull << 30)) chunk_bytes = std::min<uint64_t>(chunk_bytes * 2, 3ull << 30);
                cv.notify_all();
            }
            movi_host_reads* rd = nullptr;
            const bool filter = sinks.classify && sinks.report_mode != 0;
            if (!sinks.classify) rc = movi_gpu_file_chunk_query(my, zml ? 1 : 0, qflags, nullptr, 0, fd_bpf, my_off);
            else if (filter) {
                // --filter: the echo of the kept reads is assembled on the device; it is fetched when this chunk's turn comes
                uint64_t text_bytes = 0;
                rc = movi_gpu_file_chunk_filter(my, zml ? 1 : 0, qflags, sinks.bin_width, sinks.max_value_thr, sinks.report_mode == 2 ? 1 : 0, &text_bytes);
            } else {
                above.resize(res.n_reads + 1); below.resize(res.n_reads + 1); sum.resize(res.n_reads + 1);
                offs.resize(res.n_reads + 1); id_offs.resize(res.n_reads + 1); ids.resize(res.id_bytes + 1);
                rc = movi_gpu_file_chunk_classify(my, zml ? 1 : 0, qflags, sinks.bin_width, sinks.max_value_thr, above.data(), below.data(), sum.data(),
                                                  nullptr, 0, fd_bpf, my_off);
                if (!rc && res.n_reads) rc = movi_gpu_file_chunk_fetch(my, nullptr, offs.data(), ids.data(), id_offs.data());
                if (!rc && res.n_reads && movi_host_reads_adopt(nullptr, offs.data(), res.n_reads, ids.data(), id_offs.data(), &rd)) {
                    hard_fail(movi_host_last_error()); return;
                }
            }
            if (rc) { hard_fail(movi_gpu_last_error(my)); if (rd) movi_host_reads_free(rd); return; }
            if (sinks.classify) {
                {   // report rows / kept reads in chunk order: wait for the turn, write outside the lock (only the turn's holder writes)
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return report_seq == seq || hard; });
                    if (hard) { if (rd) movi_host_reads_free(rd); return; }
                }
                int wrc = 0;
                std::string werr;
                if (filter) {
                    if (sinks.report_fd >= 0 && (wrc = movi_gpu_file_chunk_filter_deliver(my, nullptr, 0, sinks.report_fd))) werr = movi_gpu_last_error(my);
                } else if (rd && sinks.report_fd >= 0) {
                    wrc = movi_host_classify_bins(sinks.report_fd, sinks.report_mode | (report_started ? MOVI_HOST_CLASSIFY_CONTINUE : 0), rd, sinks.max_value_thr,
                                                  above.data(), below.data(), sum.data(), nullptr);
                    if (wrc) werr = movi_host_last_error();
                }
                if (rd) movi_host_reads_free(rd);
                if (wrc) { hard_fail(werr); return; }
                std::lock_guard<std::mutex> lk(mu);
                if (res.n_reads && sinks.report_fd >= 0) report_started = true;
                report_seq += 1;
                cv.notify_all();
            }
        }
    };
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++) {
        th.emplace_back(worker, g, ctx[g]);
        if (lanes2[g]) th.emplace_back(worker, g, lanes2[g]);
    }
    for (auto& t : th) t.join();
    for (movi_gpu_ctx* c : lanes2) if (c) movi_gpu_destroy(c);
    fin->file_off = file_off; fin->report_started = report_started;
    fin->resume_at = fail_seq != NONE ? fail_begin : size;
    return hard ? -1 : (fail_seq != NONE ? 1 : 0);
}

// ---- Movi Color flows: `--multi-classify` writes "<id>,<taxon>,<taxon>" per read to the -o file
// (src/move_structure_query.cpp:439-471) next to the usual <prefix>.pml.bpf; `--sa-entries` writes <prefix>.sa_entries.bpf
// (src/utils.cpp:371-375, no header) next to it.  Batches are sharded over the GPUs like every other query.
int run_color(const Options& o, const std::vector<movi_gpu_ctx*>& ctx, const MoviGpuIndexDesc& desc, movi_host_reader* reader,
              const std::string& prefix, uint64_t min_reads, std::string* err) {
    const int gpus = (int)ctx.size();
    movi_host_colors* colors = nullptr; movi_host_ssa* ssa = nullptr;
    if (o.multi_classify) {
        if (movi_host_colors_open_kind(o.index_dir.c_str(), desc.r, o.color_vectors ? 1 : 0, &colors)) { *err = movi_host_last_error(); return -1; }
        MoviGpuColorDesc cd;
        movi_host_colors_desc(colors, &cd);
        for (int g = 0; g < gpus; g++) if (movi_gpu_upload_colors(ctx[g], &cd)) { *err = movi_gpu_last_error(ctx[g]); return -1; }
    }
    if (o.sa_entries) {
        if (movi_host_ssa_open(o.index_dir.c_str(), &ssa)) { *err = movi_host_last_error(); return -1; }
        if (movi_host_ssa_rows(ssa) != desc.r) { *err = "ssa.movi belongs to another index: run build-SA again"; return -1; }
        for (int g = 0; g < gpus; g++)
            if (movi_gpu_upload_sampled_sa(ctx[g], movi_host_ssa_rate(ssa), movi_host_ssa_sampled(ssa), movi_host_ssa_samples(ssa), movi_host_ssa_all_p(ssa))) {
                *err = movi_gpu_last_error(ctx[g]); return -1;
            }
    }
    int fd_csv = -1, fd_bpf = -1, fd_sa = -1;
    if (o.multi_classify) {
        fd_csv = open(o.out_file.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);                // open_output_files, src/utils.cpp:323-326
        if (fd_csv < 0) { *err = "cannot open " + o.out_file; return -1; }
    }
    const bool want_files = o.write_output_allowed() && !o.write_stdout_enabled();
    if (want_files) {
        fd_bpf = movi_host_bpf_begin((prefix + ".bpf").c_str());
        if (fd_bpf < 0) { *err = movi_host_last_error(); return -1; }
        if (o.sa_entries) {
            fd_sa = open((prefix + ".sa_entries.bpf").c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
            if (fd_sa < 0) { *err = "cannot open " + prefix + ".sa_entries.bpf"; return -1; }
        }
    }
    const bool want_ml = o.write_output_allowed();
    std::vector<uint16_t> ml; std::vector<uint64_t> sa; std::vector<MoviGpuColorResult> res;
    for (;;) {
        movi_host_reads* rd = nullptr;
        if (movi_host_reader_next_reads(reader, o.chunk_bases, min_reads, std::max(o.max_chunk_bases, o.chunk_bases), &rd)) { *err = movi_host_last_error(); return -1; }
        if (!rd) break;
        const uint64_t n_reads = movi_host_reads_count(rd);
        const uint64_t* offsets = movi_host_reads_offsets(rd);
        const char* bases = movi_host_reads_bases(rd);
        const uint64_t n_bases = n_reads ? offsets[n_reads] : 0;
        if (want_ml) ml.resize(n_bases + 1);
        if (o.sa_entries) sa.resize(n_bases + 1);
        if (o.multi_classify) res.resize(n_reads + 1);
        std::vector<uint64_t> cut(gpus + 1, n_reads);
        cut[0] = 0;
        for (int g = 1; g < gpus; g++) {
            const uint64_t target = n_bases / gpus * g;
            uint64_t i = cut[g - 1];
            while (i < n_reads && offsets[i] < target) i++;
            cut[g] = i;
        }
        std::vector<ShardResult> sr(gpus);
        auto worker = [&](int g) {
            if (cut[g + 1] <= cut[g]) return;
            const uint64_t i0 = cut[g], nr = cut[g + 1] - cut[g];
            int r = 0;
            if (o.multi_classify)
                r = movi_gpu_query_multi_classify(ctx[g], bases, offsets + i0, nr, 0, o.min_len, o.pvalue_scoring ? 1 : 0, res.data() + i0,
                                                  (want_ml && !o.sa_entries) ? ml.data() : nullptr);
            if (!r && o.sa_entries) r = movi_gpu_query_pml_sa(ctx[g], bases, offsets + i0, nr, 0, want_ml ? ml.data() : nullptr, sa.data());
            if (r) sr[g] = {r, movi_gpu_last_error(ctx[g])};
        };
        if (gpus == 1) worker(0);
        else {
            std::vector<std::thread> th;
            for (int g = 0; g < gpus; g++) th.emplace_back(worker, g);
            for (auto& t : th) t.join();
        }
        for (int g = 0; g < gpus; g++) if (sr[g].rc) { *err = "GPU " + std::to_string(o.device + g) + ": " + sr[g].err; return -1; }
        int rc = 0;
        if (o.multi_classify) rc = movi_host_write_multi_classify(fd_csv, rd, colors, res.data(), o.min_len, o.pvalue_scoring ? 1 : 0);
        if (!rc && o.write_output_allowed()) {
            if (o.write_stdout_enabled()) rc = movi_host_write_ml_text(STDOUT_FILENO, rd, ml.data());
            else {
                rc = movi_host_bpf_append(fd_bpf, rd, ml.data());
                if (!rc && o.sa_entries) rc = movi_host_sa_append(fd_sa, rd, sa.data());
            }
        }
        movi_host_reads_free(rd);
        if (rc) { *err = movi_host_last_error(); return -1; }
    }
    if (fd_csv >= 0) close(fd_csv);
    if (fd_sa >= 0) close(fd_sa);
    if (fd_bpf >= 0 && movi_host_bpf_end(fd_bpf)) { *err = movi_host_last_error(); return -1; }
    return 0;
}

void die(const std::string& m) { fprintf(stderr, "\033[31m\n\n[ERROR] %s\n\033[0m", m.c_str()); exit(1); }

}  // namespace

int main(int argc, char** argv) {
    Options o = parse(argc, argv);
    if (o.action == "view") { if (o.bpf_file.empty()) usage_error("view needs --bpf"); return view_bpf(o.bpf_file); }
    if (o.action != "query") usage_error("only the query (and view) actions exist in the GPU build; use the reference for " + o.action);
    if (o.index_dir.empty() || o.read_file.empty()) usage_error("query needs --index and --read");

    auto t_load0 = std::chrono::steady_clock::now();
    movi_host_index* hix = nullptr;
    if (movi_host_index_open(o.index_dir.c_str(), &hix)) die(movi_host_last_error());
    MoviGpuIndexDesc desc;
    movi_host_index_desc(hix, &desc);
    const std::string type = movi_host_index_type_name(hix);
    if ((desc.mode == MOVI_MODE_SAMPLED_THRESHOLDS || desc.mode == MOVI_MODE_SAMPLED) && o.query != "pml")     // src/movi.cpp:228-233
        die("This query type is not suppported by the sampled modes yet.");
    movi_host_reader* reader = nullptr;
    if (movi_host_reader_open(o.read_file.c_str(), o.reverse ? 1 : 0, &reader)) die(movi_host_last_error());
    auto t_files = std::chrono::steady_clock::now();

    int n_dev = movi_gpu_device_count();
    auto t_cuda = std::chrono::steady_clock::now();
    if (n_dev <= 0) die("no CUDA device: movi-b200 has no CPU path (use the reference binaries on CPU-only hosts)");
    const int gpus = std::max(1, std::min(o.gpus, n_dev));

    uint64_t max_value_thr = 0;
    // query_type() of the reference (src/utils.cpp:47-66): names the outputs and the null database
    const std::string qname = (o.query == "pml" && o.rpml) ? "rpml" : o.query;
    const uint32_t qflags = (o.query == "pml" && o.rpml) ? MOVI_GPU_FLAG_RPML : MOVI_GPU_FLAG_NONE;
    if (o.classify && movi_host_nulldb_threshold(o.index_dir.c_str(), qname.c_str(), &max_value_thr)) die(movi_host_last_error());

    // the input streams through in batches of --chunk-bases bases: batch k + 1 is parsed while the GPUs walk batch k
    // (the reference's batches: loadBatch(input, num_bases, min_number_of_reads), src/movi.cpp:281-284.)  A batch holds at least
    // one read per GPU lane -- 227 k per B200 -- so that long reads fill the machine too: a batch of 2^26 bases is 447 k reads of
    // 150 bp but only 6.7 k reads of 10 kbp, and a read's walk is one dependent chain whatever the GPU count.
    const uint64_t min_reads = o.chunk_reads ? o.chunk_reads : 230000ull * (uint64_t)gpus;
    auto fetch = [&]() {
        Fetched f;
        f.rc = movi_host_reader_next_reads(reader, o.chunk_bases, min_reads, std::max(o.max_chunk_bases, o.chunk_bases), &f.rd);
        if (f.rc) f.err = movi_host_last_error();
        return f;
    };
    if (o.multi_classify || o.sa_entries) {
        if (o.query != "pml" || o.rpml) usage_error("--multi-classify / --sa-entries go with --pml");
        if (o.classify) usage_error("--multi-classify / --sa-entries do not combine with --classify / --filter here");
        if (o.multi_classify && o.out_file.empty()) usage_error("Please include the out file for multi-classify.");     // movi_parser.cpp:396-399
    }
    // Device-side ingest: when all that is wanted of a PML / ZML query is its .bpf file (or nothing: --no-output), the file
    // bytes go to the GPUs as they are and come back as .bpf records (run_device_ingest) ...
    const bool color_flow = o.multi_classify || o.sa_entries;
    const char* raw_data = nullptr; uint64_t raw_size = 0; int raw_fastq = 0;
    movi_host_reader_raw(reader, &raw_data, &raw_size, &raw_fastq);
    // ... or its classification: --classify (report + .bpf) and --filter (the kept reads on stdout) bin the lengths on the device too
    // (with --no-output the reference prints a report header and nothing else: left to the host path, like an empty input)
    const bool plain_ingest = !o.classify && !o.write_stdout_enabled() && !(o.to_stdout && o.write_output_allowed());
    // (--filter: the echo of the kept reads is assembled on the device, filter_kernels.cuh)
    const bool classify_ingest = o.classify && !o.no_output && raw_size > 0;
    const bool ingest_on_device = !color_flow && !o.host_parse && (o.query == "pml" || o.query == "zml") && !o.reverse && o.ignore_illegal == 0 &&
                                  (plain_ingest || classify_ingest);
    std::future<Fetched> ahead;
    if (!ingest_on_device && !color_flow) ahead = std::async(std::launch::async, fetch);

    // one context per GPU, index replicated (uploaded while the first batch is being parsed)
    std::vector<movi_gpu_ctx*> ctx(gpus, nullptr);
    std::vector<ShardResult> res(gpus);
    {
        std::vector<std::thread> th;
        for (int g = 0; g < gpus; g++) th.emplace_back([&, g]() {
            int rc = movi_gpu_create(o.device + g, &ctx[g]);
            if (rc) { res[g] = {rc, movi_gpu_last_error(nullptr)}; return; }
            rc = movi_gpu_upload_index(ctx[g], &desc, o.layout);
            if (rc) res[g] = {rc, movi_gpu_last_error(ctx[g])};
        });
        for (auto& t : th) t.join();
        for (int g = 0; g < gpus; g++) if (res[g].rc) die("GPU " + std::to_string(o.device + g) + ": " + res[g].err);
    }
    auto t_q0 = std::chrono::steady_clock::now();
    if (o.verbose)
        fprintf(stderr, "[INFO] start-up: files opened in %.3f s, CUDA driver up in %.3f s, %d context(s) + index upload in %.3f s\n",
                std::chrono::duration<double>(t_files - t_load0).count(), std::chrono::duration<double>(t_cuda - t_files).count(), gpus,
                std::chrono::duration<double>(t_q0 - t_cuda).count());

    if (o.multi_classify || o.sa_entries) {
        std::string err;
        const std::string cprefix = (o.out_file.empty() ? o.read_file + "." + type : o.out_file) + "." + qname;
        if (run_color(o, ctx, desc, reader, cprefix, min_reads, &err)) die(err);
        fflush(stdout); fflush(stderr);
        _exit(0);
    }

    // With --classify / --filter the binning runs on the device (movi_gpu_query_classify); the matching lengths only come
    // back when something needs them (the .bpf file next to a --classify report; never for --filter or --no-output).
    const bool fused_classify = o.classify && o.query != "count";
    // A .bpf file as the only consumer of the lengths: its records are assembled on the device (movi_gpu_query_bpf) and the
    // writer appends one image per batch (output_base_stats / output_binary, src/utils.cpp:202-246).
    const bool device_bpf = o.query != "count" && !fused_classify && o.write_output_allowed() && !o.write_stdout_enabled();
    const bool need_ml = o.query != "count" && !device_bpf && (!fused_classify || o.write_output_allowed());

    // ---- output sinks, routed like handle_pml_zml / handle_count (src/movi.cpp:23-76)
    const std::string prefix = (o.out_file.empty() ? o.read_file + "." + type : o.out_file) + "." + qname;   // utils.cpp:348-351
    FILE* f_counts = nullptr; FILE* f_report = nullptr; int fd_bpf = -1;
    if (o.query == "count") {
        if (o.write_output_allowed() && !o.write_stdout_enabled()) {
            f_counts = fopen((prefix + ".matches").c_str(), "w");
            if (!f_counts) die("cannot open " + prefix + ".matches");
        }
    } else {
        if (o.classify && !o.filter && !o.to_stdout) {          // <reads>.<type>.<q>.report (classifier.cpp:39-60)
            f_report = fopen((o.read_file + "." + type + "." + qname + ".report").c_str(), "w");
            if (!f_report) die("cannot open the report file");
        }
        if (o.write_output_allowed() && !o.write_stdout_enabled()) {
            fd_bpf = movi_host_bpf_begin((prefix + ".bpf").c_str());
            if (fd_bpf < 0) die(movi_host_last_error());
        }
    }

    uint64_t total_reads = 0, total_bases = 0, batches = 0;
    bool report_started = false;
    if (ingest_on_device) {
        IngestStats st; IngestEnd fin; std::string err;
        IngestSinks sinks;
        sinks.fd_bpf = fd_bpf; sinks.classify = o.classify;
        sinks.report_mode = o.filter ? (o.invert ? 2 : 1) : 0;
        sinks.report_fd = (!o.filter && f_report) ? fileno(f_report) : STDOUT_FILENO;
        sinks.bin_width = (uint32_t)o.bin_width; sinks.max_value_thr = (uint32_t)max_value_thr;
        const int rc = run_device_ingest(o, ctx, raw_data, raw_size, raw_fastq, o.query == "zml", qflags, sinks, &st, &fin, &err);
        if (rc < 0) die(err);
        if (rc == 0) {
            auto t_q1 = std::chrono::steady_clock::now();
            if (o.verbose)
                fprintf(stderr, "[INFO] index loaded + uploaded in %.3f s; %llu chunk(s) ingested, queried on %d GPU(s) and written in %.3f s: %llu reads / %llu bases (device-side ingest)\n",
                        std::chrono::duration<double>(t_q0 - t_load0).count(), (unsigned long long)st.chunks, gpus,
                        std::chrono::duration<double>(t_q1 - t_q0).count(), (unsigned long long)st.reads, (unsigned long long)st.bases);
            if (f_report) fclose(f_report);
            if (fd_bpf >= 0 && movi_host_bpf_end(fd_bpf)) die(movi_host_last_error());
            fflush(stdout); fflush(stderr);
            _exit(0);
        }
        // a chunk holds something the reference rejects or stops at: everything before it is in the outputs; the host parser, which
        // reproduces the reference's behaviour, carries on from that chunk's first byte
        if (movi_host_reader_seek(reader, fin.resume_at)) die(movi_host_last_error());
        if (fd_bpf >= 0 && (ftruncate(fd_bpf, (off_t)fin.file_off) != 0 || lseek(fd_bpf, (off_t)fin.file_off, SEEK_SET) < 0)) die("cannot reposition the .bpf file");
        report_started = fin.report_started; batches = st.chunks; total_reads = st.reads; total_bases = st.bases;
        ahead = std::async(std::launch::async, fetch);
    }

    // Three stages run side by side: the parser thread reads batch k + 1, the GPUs walk batch k, the writer thread appends
    // batch k - 1 to the outputs.  Results land in pinned host buffers (DMA speed, no page faults), two sets alternating.
    Results slot[2];
    std::mt19937 gen{std::random_device{}()};
    std::uniform_int_distribution<int> dist(1, 100);

    auto write_batch = [&](Results& R) -> std::string {       // returns an error message, empty on success
        movi_host_reads* rd = R.rd;
        int rc = 0;
        if (o.query == "count") {
            if (o.write_output_allowed())
                rc = movi_host_write_counts(f_counts ? (fflush(f_counts), fileno(f_counts)) : STDOUT_FILENO, rd, R.counts.p, R.pos.p);
        } else {
            if (o.classify) {
                const int cont = report_started ? MOVI_HOST_CLASSIFY_CONTINUE : 0;
                if (!o.filter) {
                    const int fd = f_report ? (fflush(f_report), fileno(f_report)) : STDOUT_FILENO;
                    // with --no-output the reference still prints the header but no rows (write_output_allowed)
                    if (o.no_output) {
                        if (!report_started) {
                            movi_host_reads* none = nullptr; movi_host_reads_load("/dev/null", 0, &none);
                            rc = movi_host_classify(fd, 0, none, nullptr, o.bin_width, max_value_thr, nullptr); movi_host_reads_free(none);
                        }
                    } else rc = movi_host_classify_bins(fd, 0 | cont, rd, max_value_thr, R.above.p, R.below.p, R.sum.p, nullptr);
                } else if (!o.no_output) {
                    rc = movi_host_classify_bins(STDOUT_FILENO, (o.invert ? 2 : 1) | cont, rd, max_value_thr, R.above.p, R.below.p, R.sum.p, nullptr);
                }
                report_started = true;
            }
            if (!rc && o.write_output_allowed()) {
                if (o.write_stdout_enabled()) rc = movi_host_write_ml_text(STDOUT_FILENO, rd, R.ml.p);
                else if (device_bpf) rc = movi_host_bpf_append_image(fd_bpf, R.image.p, R.image_bytes);
                else rc = movi_host_bpf_append(fd_bpf, rd, R.ml.p);
            }
        }
        std::string err = rc ? movi_host_last_error() : "";
        movi_host_reads_free(rd);
        R.rd = nullptr;
        return err;
    };
    std::future<std::string> writing;

    for (;;) {
        Fetched cur = ahead.get();
        if (cur.rc) die(cur.err);
        movi_host_reads* rd = cur.rd;
        const bool last = (rd == nullptr);
        if (!last) ahead = std::async(std::launch::async, fetch);
        // an empty input still produces its (empty) outputs: run the writers once on an empty batch
        if (last) { if (batches) break; movi_host_reads_load("/dev/null", 0, &rd); }
        Results& R = slot[batches & 1];                  // (its previous batch has been written: see the wait below)
        R.rd = rd;
        const uint64_t n_reads = movi_host_reads_count(rd);
        const uint64_t* offsets = movi_host_reads_offsets(rd);
        char* bases = const_cast<char*>(movi_host_reads_bases(rd));
        const uint64_t n_bases = n_reads ? offsets[n_reads] : 0;

        if (o.ignore_illegal > 0) {   // check_alphabet's in-place substitution, src/move_structure.cpp:389-395
            const bool sep = (desc.sigma == 5 && desc.alphabet[0] == '%');
            for (uint64_t i = 0; i < n_bases; i++) {
                const unsigned char c = (unsigned char)bases[i];
                if (sep && c == '%') continue;
                if (desc.alphamap[c] == 256) bases[i] = o.ignore_illegal == 1 ? 'A' : (char)desc.alphabet[dist(gen) % desc.sigma];
            }
        }
        bool ok = true;
        if (o.query == "count") ok = R.counts.need(n_reads) && R.pos.need(n_reads);
        else if (need_ml) ok = R.ml.need(n_bases);
        const char* ids = movi_host_reads_ids(rd);
        const uint64_t* id_offsets = movi_host_reads_id_offsets(rd);
        if (device_bpf) {
            R.image_bytes = n_reads ? movi_gpu_bpf_image_bytes(offsets, id_offsets, n_reads) : 0;
            ok = ok && R.image.need(R.image_bytes);
        }
        if (ok && fused_classify) ok = R.above.need(n_reads) && R.below.need(n_reads) && R.sum.need(n_reads);
        if (!ok) die("cannot allocate pinned host memory for the results of a batch (try a smaller --chunk-bases)");

        // contiguous shards balanced by bases (same decomposition as the reference's OpenMP batches, movi.cpp:274-301)
        std::vector<uint64_t> cut(gpus + 1, n_reads);
        cut[0] = 0;
        for (int g = 1; g < gpus; g++) {
            const uint64_t target = n_bases / gpus * g;
            uint64_t i = cut[g - 1];
            while (i < n_reads && offsets[i] < target) i++;
            cut[g] = i;
        }
        auto worker = [&](int g) {
            if (cut[g + 1] <= cut[g]) return;
            const uint64_t i0 = cut[g], nr = cut[g + 1] - cut[g];
            int r;
            // offsets stay absolute: the entry points index bases / outputs by them
            if (fused_classify)
                r = movi_gpu_query_classify(ctx[g], o.query == "zml", bases, offsets + i0, nr, qflags, (uint32_t)o.bin_width, (uint32_t)max_value_thr,
                                            R.above.p + i0, R.below.p + i0, R.sum.p + i0, need_ml ? R.ml.p : nullptr);
            else if (device_bpf) {
                const uint64_t at = movi_gpu_bpf_image_bytes(offsets, id_offsets, i0);       // where read i0's record starts in the image
                r = movi_gpu_query_bpf(ctx[g], o.query == "zml", bases, offsets + i0, nr, qflags, ids, id_offsets + i0, R.image.p + at, R.image_bytes - at);
            }
            else if (o.query == "pml") r = movi_gpu_query_pml(ctx[g], bases, offsets + i0, nr, qflags, R.ml.p);
            else if (o.query == "zml") r = movi_gpu_query_zml(ctx[g], bases, offsets + i0, nr, 0, R.ml.p);
            else r = movi_gpu_query_count(ctx[g], bases, offsets + i0, nr, 0, R.counts.p + i0, R.pos.p + i0);
            if (r) res[g] = {r, movi_gpu_last_error(ctx[g])};
        };
        if (gpus == 1) worker(0);
        else {
            std::vector<std::thread> th;
            for (int g = 0; g < gpus; g++) th.emplace_back(worker, g);
            for (auto& t : th) t.join();
        }
        for (int g = 0; g < gpus; g++) if (res[g].rc) die("GPU " + std::to_string(o.device + g) + ": " + res[g].err);

        // hand the batch to the writer once the previous one is on disk (outputs are appended in read order)
        if (writing.valid()) { const std::string err = writing.get(); if (!err.empty()) die(err); }
        writing = std::async(std::launch::async, write_batch, std::ref(R));
        total_reads += n_reads; total_bases += n_bases; batches += 1;
        if (last) break;
    }
    if (writing.valid()) { const std::string err = writing.get(); if (!err.empty()) die(err); }
    auto t_q1 = std::chrono::steady_clock::now();
    if (o.verbose)
        fprintf(stderr, "[INFO] index loaded + uploaded in %.3f s; %llu batch(es) parsed, queried on %d GPU(s) and written in %.3f s: %llu reads / %llu bases\n",
                std::chrono::duration<double>(t_q0 - t_load0).count(), (unsigned long long)batches, gpus,
                std::chrono::duration<double>(t_q1 - t_q0).count(), (unsigned long long)total_reads, (unsigned long long)total_bases);
    if (f_counts) fclose(f_counts);
    if (f_report) fclose(f_report);
    if (fd_bpf >= 0 && movi_host_bpf_end(fd_bpf)) die(movi_host_last_error());
    // Everything is on disk.  Tearing down the CUDA contexts, the pinned buffers and the mappings takes a few tenths of a second
    // that a one-shot command has no use for: leave that to the operating system.
    fflush(stdout); fflush(stderr);
    _exit(0);
}
