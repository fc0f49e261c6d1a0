// build_index -- writes a Movi `index.movi` (regular- / blocked- / sampled-thresholds) from a BWT and its
// thresholds.  BENCH / TEST DATA GENERATOR: index construction is outside the GPU query path,
// but the benchmarks need real indexes on a box that has neither the reference sources nor pfp-thresholds.
//
// An independent restatement of the reference builder for MODE 6 / 8 / 7 without separators
// (src/move_structure_build.cpp: run splitting :223-426, run-head ranks :74-122, rows :449-692, base
// intervals :694-731, thresholds :807-935, blocked ids :939-1074; serialisation src/move_structure_io.cpp:436-469).
// tests/test_datagen.py checks its output byte for byte against index files written by the unmodified reference.
//
//   build_index <prefix> <out-dir> <regular-thresholds|blocked-thresholds|sampled-thresholds> [--rle] [--checkpoint N]
//     reads <prefix>.bwt (raw, '\0' = '$') or, with --rle, <prefix>.bwt.heads + <prefix>.bwt.len (5-byte lengths),
//     and <prefix>.thr_pos (5 bytes per BWT run); writes <out-dir>/index.movi
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace {

[[noreturn]] void die(const std::string& m) { fprintf(stderr, "build_index: %s\n", m.c_str()); exit(1); }

const uint32_t alphamap_3[4][4] = {{3, 0, 1, 2}, {0, 3, 1, 2}, {0, 1, 3, 2}, {0, 1, 2, 3}};   // reference src/utils.cpp:5-8

std::vector<uint8_t> slurp(const std::string& f) {
    std::ifstream in(f, std::ios::binary);
    if (!in) die("cannot open " + f);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
}
std::vector<uint64_t> read5(const std::string& f) {
    std::vector<uint8_t> b = slurp(f);
    std::vector<uint64_t> v(b.size() / 5);
    for (size_t i = 0; i < v.size(); i++) { uint64_t x = 0; memcpy(&x, &b[5 * i], 5); v[i] = x; }
    return v;
}
template <typename T> void put(std::ofstream& o, const T& v) { o.write(reinterpret_cast<const char*>(&v), sizeof(T)); }

}  // namespace

int main(int argc, char** argv) {
    if (argc < 4) die("usage: build_index <prefix> <out-dir> <type> [--rle] [--checkpoint N]");
    const std::string prefix = argv[1], outdir = argv[2], type = argv[3];
    bool rle = false; uint32_t checkpoint = 20;                 // tally_checkpoints default, include/movi_options.hpp:257
    for (int a = 4; a < argc; a++) {
        if (!strcmp(argv[a], "--rle")) rle = true;
        else if (!strcmp(argv[a], "--checkpoint") && a + 1 < argc) checkpoint = (uint32_t)atoi(argv[++a]);
    }
    int mode; uint64_t max_run;
    if (type == "regular-thresholds") { mode = 6; max_run = 2047; }          // include/move_row_configs.hpp:50
    else if (type == "blocked-thresholds") { mode = 8; max_run = 1023; }     // :101
    else if (type == "sampled-thresholds") { mode = 7; max_run = 511; }      // :135
    else die("unknown index type " + type);

    // ---- original BWT runs
    std::vector<uint8_t> oheads; std::vector<uint64_t> olens;
    if (rle) {
        oheads = slurp(prefix + ".bwt.heads");
        olens = read5(prefix + ".bwt.len");
        if (oheads.size() != olens.size()) die("heads / len size mismatch");
    } else {
        std::vector<uint8_t> bwt = slurp(prefix + ".bwt");
        for (size_t i = 0; i < bwt.size();) {
            size_t j = i; while (j < bwt.size() && bwt[j] == bwt[i]) j++;
            oheads.push_back(bwt[i]); olens.push_back(j - i); i = j;
        }
    }
    const uint64_t original_r = oheads.size();
    std::vector<uint64_t> thresholds = read5(prefix + ".thr_pos");
    if (thresholds.size() != original_r) die("thr_pos has " + std::to_string(thresholds.size()) + " entries for " + std::to_string(original_r) + " runs");
    uint64_t n = 0; for (uint64_t l : olens) n += l;

    // ---- split points: bit 0 and every threshold position (fill_bits_by_thresholds, build.cpp:733-746)
    std::vector<uint64_t> bits((n + 1 + 63) / 64 + 1, 0);
    auto setbit = [&](uint64_t p) { bits[p >> 6] |= 1ULL << (p & 63); };
    auto getbit = [&](uint64_t p) { return (bits[p >> 6] >> (p & 63)) & 1ULL; };
    setbit(0);
    for (uint64_t t : thresholds) { if (t > n) die("threshold beyond the BWT"); setbit(t); }
    auto next_set = [&](uint64_t from, uint64_t end) -> uint64_t {   // first set bit in [from, end), or end
        while (from < end) {
            uint64_t w = bits[from >> 6] >> (from & 63);
            if (w) { uint64_t p = from + __builtin_ctzll(w); return p < end ? p : end; }
            from = (from | 63) + 1;
        }
        return end;
    };

    // ---- move rows after threshold and max-run splitting (detect_move_row_boundaries, build.cpp:223-426)
    std::vector<uint8_t> heads; std::vector<uint16_t> lens; std::vector<uint64_t> all_p;
    std::vector<uint64_t> occ(256, 0);
    {
        uint64_t pos = 0;
        auto emit = [&](uint8_t c, uint64_t start, uint64_t len) {
            while (len > max_run) { heads.push_back(c); lens.push_back((uint16_t)max_run); all_p.push_back(start); start += max_run; len -= max_run; }
            if (len > 0) { heads.push_back(c); lens.push_back((uint16_t)len); all_p.push_back(start); }
        };
        for (uint64_t i = 0; i < original_r; i++) {
            const uint64_t len = olens[i], end = pos + len;
            if (oheads[i] != 0) occ[oheads[i]] += len;
            uint64_t run_head = pos;
            for (uint64_t j = next_set(pos + 1, end); j < end; j = next_set(j + 1, end)) { emit(oheads[i], run_head, j - run_head); run_head = j; }
            emit(oheads[i], run_head, end - run_head);
            pos = end;
        }
    }
    const uint64_t r = heads.size();
    all_p.push_back(n);

    // ---- alphabet (build_alphabet, build.cpp:428-447)
    std::vector<uint64_t> alphamap(256, 256), counts; std::vector<uint8_t> alphabet;
    for (int c = 1; c < 256; c++) if (occ[c]) { alphamap[c] = alphabet.size(); alphabet.push_back((uint8_t)c); counts.push_back(occ[c]); }
    const uint64_t sigma = alphabet.size();
    if (sigma > 4 || sigma < 1) die("this generator supports 1..4 characters without separators");

    // ---- run-head ranks and the '$' row (find_run_heads_information, build.cpp:74-122)
    std::vector<uint64_t> heads_rank(r, 0), seen(sigma, 0);
    uint64_t end_bwt_idx = 0; bool have_end = false;
    for (uint64_t i = 0; i < r; i++) {
        if (heads[i] == 0) { end_bwt_idx = i; have_end = true; heads_rank[i] = 0; continue; }
        const uint64_t a = alphamap[heads[i]];
        heads_rank[i] = seen[a]; seen[a] += lens[i];
    }
    if (!have_end) die("the BWT has no terminator run");
    auto row_c = [&](uint64_t i) -> uint32_t { return heads[i] == 0 ? 0u : (uint32_t)alphamap[heads[i]]; };   // set_c: 256 truncates to 0

    // ---- LF of every row head -> destination row and offset (build_move_rows, build.cpp:449-692)
    std::vector<uint64_t> ids(r, 0); std::vector<uint16_t> offs(r, 0);
    {
        std::vector<uint64_t> cstart(sigma + 1, 1);
        for (uint64_t a = 0; a < sigma; a++) cstart[a + 1] = cstart[a] + counts[a];
        uint64_t p = 0;
        for (uint64_t a = 0; a < sigma; a++) {
            for (uint64_t i = 0; i < r; i++) {
                if (heads[i] == 0 || alphamap[heads[i]] != a) continue;
                const uint64_t lf = cstart[a] + heads_rank[i];          // LF_heads, move_structure.cpp:515-522
                while (p + 1 < r && all_p[p + 1] <= lf) p++;
                ids[i] = p; offs[i] = (uint16_t)(lf - all_p[p]);
                if (lf - all_p[p] > max_run) die("offset does not fit");
            }
        }
        ids[end_bwt_idx] = 0; offs[end_bwt_idx] = 0;                    // lf = 0 for '$' (build.cpp:520-522)
    }

    // ---- backward-search seed intervals (find_base_interval_data, build.cpp:694-731)
    std::vector<uint64_t> first_runs{0}, first_offsets{0}, last_runs{0}, last_offsets{0};
    {
        uint64_t char_count = 1;
        for (uint64_t a = 0; a < sigma; a++) {
            const uint64_t lr = last_runs.back(), lo = last_offsets.back();
            if (lo + 1 >= lens[lr]) { first_runs.push_back(lr + 1); first_offsets.push_back(0); }
            else { first_runs.push_back(lr); first_offsets.push_back(lo + 1); }
            char_count += counts[a];
            // number of row starts in [0, char_count) - 1
            const uint64_t row = (uint64_t)(std::upper_bound(all_p.begin(), all_p.begin() + r, char_count - 1) - all_p.begin()) - 1;
            last_runs.push_back(row); last_offsets.push_back(char_count - all_p[row] - 1);
        }
    }

    // ---- threshold bits (compute_thresholds, build.cpp:807-935): after splitting, every threshold is 0 or n
    std::vector<uint8_t> thr_bits(r, 0);
    uint64_t end_thr[4] = {0, 0, 0, 0};
    {
        std::vector<uint64_t> at(sigma, n);
        uint64_t thr_i = original_r - 1;
        for (uint64_t i = r - 1; i > 0; --i) {
            const uint32_t ci = row_c(i);                 // the '$' row is read as alphabet[0], as in the reference
            for (uint64_t j = 0; j < sigma; j++) {
                if (j == ci) { at[j] = thresholds[thr_i]; continue; }
                const uint32_t slot = (i == end_bwt_idx) ? (uint32_t)j : alphamap_3[ci][j];
                const uint64_t cur = at[j];
                uint64_t value; int split;
                if (cur >= all_p[i] + lens[i]) { value = lens[i]; split = 1; }
                else if (cur <= all_p[i]) { value = 0; split = 0; }
                else die("a threshold falls inside a row: the runs were not split at it");
                if (i == end_bwt_idx) end_thr[slot] = value;
                else if (split) thr_bits[i] |= (uint8_t)(1u << slot);
            }
            if (row_c(i) != row_c(i - 1) || i == end_bwt_idx || i - 1 == end_bwt_idx) thr_i--;
        }
    }

    // ---- blocked ids (compute_blocked_ids, build.cpp:939-1074)
    std::vector<std::vector<uint32_t>> id_blocks; uint64_t block_size = 1048576; std::vector<uint64_t> stored_id = ids;
    if (mode == 8) {
        uint64_t max_allowed = 4194303;
        for (;;) {
            id_blocks.assign(sigma, {});
            std::vector<uint32_t> block_start(sigma, 0);
            uint64_t block_count = 0; bool small = false;
            for (uint64_t i = 0; i < r; i++) {
                if (i >= block_size * block_count) { block_count++; for (uint64_t j = 0; j < sigma; j++) id_blocks[j].push_back(block_start[j]); }
                if (i == end_bwt_idx) { stored_id[i] = 0; continue; }
                const uint32_t c = row_c(i);
                const uint64_t adjusted = ids[i] - first_runs[c + 1];
                const uint64_t blocked = adjusted - id_blocks[c][i / block_size];
                if (blocked > max_allowed) { small = true; break; }
                stored_id[i] = blocked;
                block_start[c] = (uint32_t)adjusted;
            }
            if (!small) break;
            block_size /= 2; max_allowed = ((max_allowed + 1) / 2) - 1;
        }
    }

    // ---- id checkpoints (build_move_rows TALLY branch, build.cpp:487-497, 561-583, 679-684)
    std::vector<std::vector<uint64_t>> tally;
    uint64_t tally_len = 0;
    if (mode == 7) {
        tally_len = r / checkpoint + 2;
        tally.assign(sigma, std::vector<uint64_t>(tally_len, 0));
        std::vector<uint64_t> current(sigma, r);
        for (uint64_t i = 0; i < r; i++) {
            if (i != end_bwt_idx) {
                const uint32_t c = row_c(i);
                if (current[c] == r) for (uint64_t t = 0; t <= i / checkpoint; t++) tally[c][t] = ids[i];
                current[c] = ids[i];
            }
            if (i % checkpoint == 0) for (uint64_t a = 0; a < sigma; a++) tally[a][i / checkpoint] = current[a];
        }
        for (uint64_t a = 0; a < sigma; a++) tally[a][tally_len - 1] = current[a];
    }

    // ---- serialize (reference src/move_structure_io.cpp:436-469; layout SURVEY A.2)
    mkdir(outdir.c_str(), 0777);
    std::ofstream o(outdir + "/index.movi", std::ios::binary);
    if (!o) die("cannot write " + outdir + "/index.movi");
    {
        uint8_t h[48] = {0};
        const uint32_t magic = 0x4D4F5649u; memcpy(h, &magic, 4);
        h[4] = 2; h[5] = 0; h[6] = 0; h[7] = (uint8_t)mode; h[8] = 0;
        memcpy(h + 16, &n, 8); memcpy(h + 24, &r, 8); memcpy(h + 32, &original_r, 8); memcpy(h + 40, &end_bwt_idx, 8);
        o.write(reinterpret_cast<char*>(h), 48);
    }
    for (int k = 0; k < 4; k++) put<uint64_t>(o, end_thr[k]);
    for (int k = 0; k < 8; k++) put<uint64_t>(o, 0);                      // next_down / next_up (constant mode only)
    put<uint64_t>(o, 256); for (uint64_t a : alphamap) put<uint64_t>(o, a);
    put<uint64_t>(o, sigma); o.write(reinterpret_cast<const char*>(alphabet.data()), sigma);
    put<uint16_t>(o, 0); put<uint8_t>(o, 0);                              // nt_splitting, constant
    for (uint64_t i = 0; i < r; i++) {
        const uint32_t nn = lens[i], off = offs[i], c = row_c(i), t = thr_bits[i];
        if (mode == 6) {        // include/move_row.hpp:131-142 + move_row_configs.hpp:34-51
            const uint32_t id_lo = (uint32_t)stored_id[i];
            const uint16_t fn = (uint16_t)(nn | (((t >> 1) & 1) << 11) | (((t >> 2) & 1) << 12) | (c << 13));
            const uint16_t fo = (uint16_t)(off | ((t & 1) << 11) | (uint32_t)((stored_id[i] >> 32) & 0xF) << 12);
            put(o, id_lo); put(o, fn); put(o, fo);
        } else if (mode == 8) { // move_row_configs.hpp:76-104
            const uint16_t fid = (uint16_t)(stored_id[i] & 0xFFFF);
            const uint16_t fn = (uint16_t)(nn | (uint32_t)((stored_id[i] >> 16) & 0x3F) << 10);
            const uint16_t fo = (uint16_t)(off | (c << 10) | ((t & 7) << 13));
            put(o, fid); put(o, fn); put(o, fo);
        } else {                // move_row_configs.hpp:120-136
            const uint8_t b0 = (uint8_t)(nn & 0xFF), b1 = (uint8_t)(off & 0xFF);
            const uint8_t b2 = (uint8_t)(((off >> 8) & 1) | (((nn >> 8) & 1) << 1) | (c << 2) | ((t & 7) << 5));
            put(o, b0); put(o, b1); put(o, b2);
        }
    }
    if (mode == 7) {
        put<uint32_t>(o, checkpoint); put<uint64_t>(o, tally_len);
        for (uint64_t a = 0; a < sigma; a++)
            for (uint64_t t = 0; t < tally_len; t++) { const uint64_t v = tally[a][t]; o.write(reinterpret_cast<const char*>(&v), 5); }
    }
    put<uint64_t>(o, 0); put<uint64_t>(o, 0); put<uint64_t>(o, 0);        // overflow tables (unused by these layouts)
    put<uint64_t>(o, sigma); for (uint64_t c : counts) put<uint64_t>(o, c);
    put<uint64_t>(o, sigma + 1);
    for (uint64_t v : last_runs) put<uint64_t>(o, v);
    for (uint64_t v : last_offsets) put<uint64_t>(o, v);
    for (uint64_t v : first_runs) put<uint64_t>(o, v);
    for (uint64_t v : first_offsets) put<uint64_t>(o, v);
    if (mode == 8) {
        put<uint64_t>(o, (uint64_t)id_blocks[0].size());
        for (uint64_t a = 0; a < sigma; a++) o.write(reinterpret_cast<const char*>(id_blocks[a].data()), id_blocks[a].size() * 4);
        put<uint64_t>(o, block_size);
    }
    o.close();
    fprintf(stderr, "build_index: %s n=%llu r=%llu original_r=%llu end_bwt_idx=%llu\n", type.c_str(),
            (unsigned long long)n, (unsigned long long)r, (unsigned long long)original_r, (unsigned long long)end_bwt_idx);
    return 0;
}
