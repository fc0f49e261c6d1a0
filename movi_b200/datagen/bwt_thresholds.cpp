// bwt_thresholds -- suffix-array based stand-in for the un-vendored `pfp_thresholds -P -r -f`
// step of Movi's build pipeline (reference src/movi_launcher.cpp:204-212; its outputs are
// consumed at src/move_structure_build.cpp:19-20,145-199 and src/utils.cpp:150-200).
//
// BENCH / TEST DATA GENERATOR, not part of the query path.  It manufactures the inputs an
// index builder needs from a (cleaned) FASTA:
//     <out>.bwt            raw BWT bytes, '\0' standing for '$'
//  or <out>.bwt.heads / <out>.bwt.len (1 B + 5 B LE per run; `--rle`)
//     <out>.thr_pos        one 5-byte LE integer per BWT run: 0 for the first run of each
//                          character, else the FIRST position p in (end of the previous run
//                          of that character, start of this run] that minimises LCP[p].
// Conventions (text = all sequence lines concatenated + one '\0'; plain byte order; the
// first-argmin tie-break) are pinned by the reference's golden index sizes and PMLs: see
// tests/test_datagen.py and SURVEY.md Appendix B.
//
// Algorithms: SA-IS (Nong, Zhang, Chan 2009) for the suffix array, the Phi/PLCP method
// (Karkkainen, Manzini, Puglisi 2009) for LCP, parallelised over text blocks with OpenMP.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include <chrono>

namespace {

template <typename Idx>
struct Sais {
    static constexpr Idx EMPTY = static_cast<Idx>(-1);

    template <typename Ch>
    static void get_buckets(const Ch* s, Idx* bkt, Idx n, Idx K, bool end) {
        for (Idx i = 0; i <= K; i++) bkt[i] = 0;
        for (Idx i = 0; i < n; i++) bkt[static_cast<Idx>(s[i])]++;
        Idx sum = 0;
        for (Idx i = 0; i <= K; i++) {
            sum += bkt[i];
            bkt[i] = end ? sum : sum - bkt[i];
        }
    }

    static inline bool tget(const std::vector<uint64_t>& t, Idx i) { return (t[i >> 6] >> (i & 63)) & 1ULL; }
    static inline void tset(std::vector<uint64_t>& t, Idx i, bool b) {
        if (b) t[i >> 6] |= (1ULL << (i & 63)); else t[i >> 6] &= ~(1ULL << (i & 63));
    }
    static inline bool is_lms(const std::vector<uint64_t>& t, Idx i) { return i > 0 && tget(t, i) && !tget(t, i - 1); }

    template <typename Ch>
    static void induce_l(const std::vector<uint64_t>& t, Idx* SA, const Ch* s, Idx* bkt, Idx n, Idx K) {
        get_buckets(s, bkt, n, K, false);
        for (Idx i = 0; i < n; i++) {
            Idx j = SA[i];
            if (j != EMPTY && j > 0) {
                j -= 1;
                if (!tget(t, j)) SA[bkt[static_cast<Idx>(s[j])]++] = j;
            }
        }
    }
    template <typename Ch>
    static void induce_s(const std::vector<uint64_t>& t, Idx* SA, const Ch* s, Idx* bkt, Idx n, Idx K) {
        get_buckets(s, bkt, n, K, true);
        for (Idx i = n - 1; i >= 0; i--) {
            Idx j = SA[i];
            if (j != EMPTY && j > 0) {
                j -= 1;
                if (tget(t, j)) SA[--bkt[static_cast<Idx>(s[j])]] = j;
            }
        }
    }

    // s[n-1] must be the unique smallest symbol. Symbols in [0, K].
    template <typename Ch>
    static void run(const Ch* s, Idx* SA, Idx n, Idx K) {
        if (n == 1) { SA[0] = 0; return; }
        std::vector<uint64_t> t((static_cast<size_t>(n) >> 6) + 1, 0);
        tset(t, n - 2, false); tset(t, n - 1, true);
        for (Idx i = n - 3; i >= 0; i--)
            tset(t, i, (s[i] < s[i + 1] || (s[i] == s[i + 1] && tget(t, i + 1))));
        std::vector<Idx> bkt(static_cast<size_t>(K) + 1);

        get_buckets(s, bkt.data(), n, K, true);
        for (Idx i = 0; i < n; i++) SA[i] = EMPTY;
        for (Idx i = 1; i < n; i++) if (is_lms(t, i)) SA[--bkt[static_cast<Idx>(s[i])]] = i;
        induce_l(t, SA, s, bkt.data(), n, K);
        induce_s(t, SA, s, bkt.data(), n, K);

        Idx n1 = 0;
        for (Idx i = 0; i < n; i++) if (is_lms(t, SA[i])) SA[n1++] = SA[i];
        for (Idx i = n1; i < n; i++) SA[i] = EMPTY;
        Idx name = 0, prev = EMPTY;
        for (Idx i = 0; i < n1; i++) {
            Idx pos = SA[i]; bool diff = false;
            for (Idx d = 0; d < n; d++) {
                if (prev == EMPTY || s[pos + d] != s[prev + d] || tget(t, pos + d) != tget(t, prev + d)) { diff = true; break; }
                else if (d > 0 && (is_lms(t, pos + d) || is_lms(t, prev + d))) break;
            }
            if (diff) { name++; prev = pos; }
            SA[n1 + (pos >> 1)] = name - 1;
        }
        for (Idx i = n - 1, j = n - 1; i >= n1; i--) if (SA[i] != EMPTY) SA[j--] = SA[i];

        Idx* SA1 = SA; Idx* s1 = SA + n - n1;
        if (name < n1) run<Idx>(s1, SA1, n1, name - 1);
        else for (Idx i = 0; i < n1; i++) SA1[s1[i]] = i;

        get_buckets(s, bkt.data(), n, K, true);
        for (Idx i = 1, j = 0; i < n; i++) if (is_lms(t, i)) s1[j++] = i;
        for (Idx i = 0; i < n1; i++) SA1[i] = s1[SA1[i]];
        for (Idx i = n1; i < n; i++) SA[i] = EMPTY;
        for (Idx i = n1 - 1; i >= 0; i--) {
            Idx j = SA[i]; SA[i] = EMPTY;
            SA[--bkt[static_cast<Idx>(s[j])]] = j;
        }
        induce_l(t, SA, s, bkt.data(), n, K);
        induce_s(t, SA, s, bkt.data(), n, K);
    }
};

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void write5(std::ofstream& f, uint64_t v) { f.write(reinterpret_cast<const char*>(&v), 5); }

template <typename Idx>
int generate(const std::vector<uint8_t>& T, const std::string& out, bool rle, bool verbose) {
    const Idx n = static_cast<Idx>(T.size());
    double t0 = now();
    std::vector<Idx> SA(static_cast<size_t>(n));
    Sais<Idx>::template run<uint8_t>(T.data(), SA.data(), n, 255);
    if (verbose) fprintf(stderr, "[bwt_thresholds] suffix array: %.1f s\n", now() - t0);

    // Phi / PLCP, blockwise parallel: PLCP[i] = lcp(suffix i, suffix Phi[i]) with Phi[SA[p]] = SA[p-1].
    t0 = now();
    std::vector<Idx> plcp(static_cast<size_t>(n));
    #pragma omp parallel for schedule(static)
    for (Idx p = 0; p < n; p++) plcp[SA[p]] = (p == 0) ? static_cast<Idx>(-1) : SA[p - 1];
    const Idx nblocks = 256;
    #pragma omp parallel for schedule(dynamic, 1)
    for (Idx b = 0; b < nblocks; b++) {
        Idx lo = static_cast<Idx>((static_cast<__int128>(n) * b) / nblocks);
        Idx hi = static_cast<Idx>((static_cast<__int128>(n) * (b + 1)) / nblocks);
        Idx l = 0;
        for (Idx i = lo; i < hi; i++) {
            Idx j = plcp[i];
            if (j == static_cast<Idx>(-1)) { plcp[i] = 0; l = 0; continue; }
            while (i + l < n && j + l < n && T[i + l] == T[j + l]) l++;
            plcp[i] = l;
            if (l > 0) l--;
        }
    }
    if (verbose) fprintf(stderr, "[bwt_thresholds] PLCP: %.1f s\n", now() - t0);

    t0 = now();
    std::ofstream bwt_f, heads_f, len_f, thr_f(out + ".thr_pos", std::ios::binary);
    if (rle) {
        heads_f.open(out + ".bwt.heads", std::ios::binary);
        len_f.open(out + ".bwt.len", std::ios::binary);
    } else {
        bwt_f.open(out + ".bwt", std::ios::binary);
    }
    const uint64_t INF = ~0ULL;
    uint64_t minv[256], minp[256]; bool seen[256];
    for (int x = 0; x < 256; x++) { minv[x] = INF; minp[x] = 0; seen[x] = false; }
    // Only symbols that occur can ever start a run; restrict the per-position update to them.
    std::vector<int> syms;
    { bool occ[256] = {false}; for (Idx i = 0; i < n; i++) occ[T[i]] = true;
      for (int x = 0; x < 256; x++) if (occ[x]) syms.push_back(x); }
    std::vector<char> bwt_buf; bwt_buf.reserve(1 << 20);
    uint64_t runs = 0, run_len = 0; int prev_c = -1;
    // gather BWT[p] and LCP[p] in SA order block by block (the two random gathers are the expensive part: parallel),
    // then scan the block sequentially
    const Idx BLK = 1 << 24;
    std::vector<uint8_t> bw(static_cast<size_t>(std::min<Idx>(BLK, n)));
    std::vector<Idx> lc(bw.size());
    for (Idx base = 0; base < n; base += BLK) {
        const Idx cnt = std::min<Idx>(BLK, n - base);
        #pragma omp parallel for schedule(static)
        for (Idx k = 0; k < cnt; k++) {
            const Idx sa = SA[base + k];
            bw[k] = T[sa == 0 ? n - 1 : sa - 1];
            lc[k] = plcp[sa];
        }
        for (Idx k = 0; k < cnt; k++) {
            const Idx p = base + k;
            const int c = bw[k];
            const uint64_t lcp = static_cast<uint64_t>(lc[k]);
            for (int x : syms) if (lcp < minv[x]) { minv[x] = lcp; minp[x] = static_cast<uint64_t>(p); }
            if (c != prev_c) {
                if (prev_c >= 0 && rle) { char h = static_cast<char>(prev_c); heads_f.write(&h, 1); write5(len_f, run_len); }
                write5(thr_f, seen[c] ? minp[c] : 0);
                seen[c] = true; runs++; run_len = 0; prev_c = c;
            }
            run_len++;
            minv[c] = INF; minp[c] = static_cast<uint64_t>(p) + 1;
            if (!rle) {
                bwt_buf.push_back(static_cast<char>(c));
                if (bwt_buf.size() == (1 << 20)) { bwt_f.write(bwt_buf.data(), bwt_buf.size()); bwt_buf.clear(); }
            }
        }
    }
    if (rle) { char h = static_cast<char>(prev_c); heads_f.write(&h, 1); write5(len_f, run_len); }
    else if (!bwt_buf.empty()) bwt_f.write(bwt_buf.data(), bwt_buf.size());
    if (verbose) fprintf(stderr, "[bwt_thresholds] BWT + thresholds: %.1f s; n = %llu, runs = %llu\n",
                         now() - t0, static_cast<unsigned long long>(n), static_cast<unsigned long long>(runs));
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: bwt_thresholds <clean.fasta | --raw text.bin> <out-prefix> [--rle] [-q]\n");
        return 2;
    }
    int a = 1; bool raw = false;
    if (std::string(argv[a]) == "--raw") { raw = true; a++; }
    std::string in = argv[a++], out = argv[a++];
    bool rle = false, verbose = true;
    for (; a < argc; a++) {
        if (std::string(argv[a]) == "--rle") rle = true;
        else if (std::string(argv[a]) == "-q") verbose = false;
    }
    std::vector<uint8_t> T;
    {
        std::ifstream f(in, std::ios::binary);
        if (!f) { fprintf(stderr, "cannot open %s\n", in.c_str()); return 1; }
        if (raw) {
            T.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
        } else {
            std::string line;
            while (std::getline(f, line)) {
                if (!line.empty() && line[0] == '>') continue;
                while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
                T.insert(T.end(), line.begin(), line.end());
            }
        }
        for (uint8_t c : T) if (c == 0) { fprintf(stderr, "input contains a 0 byte\n"); return 1; }
        T.push_back(0);
    }
    if (T.size() < (1ULL << 31) - 2) return generate<int32_t>(T, out, rle, verbose);
    return generate<int64_t>(T, out, rle, verbose);
}
