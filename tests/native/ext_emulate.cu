// ext_emulate.cu -- TEST INFRASTRUCTURE: runs movi_b200/csrc/ext_kernel.cuh's per-lane state machine (the very source the GPU
// kernel movi_ext_kernel is made of: ExtLane::start / step are __host__ __device__) on the CPU, a few lanes taking turns, so that
// the Movi Color walks are checked against the oracle and the reference's outputs in the CPU test run.  Nothing here is product
// code and nothing in the product links it; it is not a fallback (the library's entry points need a CUDA device).
//
// Only what the emulation needs of the upload is restated: regular-thresholds rows -> the device row format (r8_from_file, the
// function relayout_regular_kernel applies) and the side fields of DevIndex (movi_gpu.cu, upload_impl).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../movi_b200/csrc/ext_kernel.cuh"

using namespace movi;

namespace {

template <typename IdxT, bool SEP, int EXT>
void run(const DevIndex& ix, const ExtArgs& a, int lanes) {
    std::vector<ExtLane<IdxT, SEP, EXT>> L(lanes);
    unsigned long long next = 0;
    std::vector<char> done(lanes, 0);
    int live = lanes;
    while (live) {
        for (int l = 0; l < lanes; l++) {
            if (done[l]) continue;
            if (L[l].st == ST_IDLE) {
                if (next >= a.n_reads) { done[l] = 1; live--; continue; }
                L[l].start(ix, a, (uint32_t)next++, (unsigned long long)l);
                continue;
            }
            L[l].step(ix, a, (unsigned long long)l);
        }
    }
}

}  // namespace

extern "C" int movi_ext_emulate(const MoviGpuIndexDesc* d, int ext, const MoviGpuColorDesc* col, uint64_t rate, const uint64_t* sampled,
                                const uint64_t* all_p, const char* bases, const uint64_t* offsets, uint64_t n_reads, uint32_t min_len,
                                int lanes, int force_wide, MoviGpuColorResult* out_color, uint16_t* out_ml, uint64_t* out_sa) {
    if (!d || d->mode != MOVI_MODE_REGULAR_THRESHOLDS || lanes < 1) return -1;
    const bool sep = d->sigma == 5 && d->alphabet[0] == '%';
    std::vector<uint2> rows(d->r + 4, make_uint2(0, 0));
    const uint2* file_rows = static_cast<const uint2*>(d->rows);
    for (uint64_t i = 0; i < d->r; i++) rows[i] = r8_from_file(file_rows[i], sep ? 1u : 0u);
    DevIndex ix;
    std::memset(&ix, 0, sizeof ix);
    ix.rows = rows.data(); ix.layout = LAYOUT_R8; ix.r = d->r; ix.end_idx = d->end_bwt_idx; ix.sep = sep; ix.sigma = d->sigma;
    for (int k = 0; k < 4; k++) ix.end_thr[k] = (uint32_t)std::min<uint64_t>(d->end_bwt_idx_thresholds[k], 0xFFFFFFFFull);
    for (int b = 0; b < 256; b++) {
        const uint64_t am = d->alphamap[b];
        ix.code_of[b] = (am < d->sigma && !(sep && b == '%')) ? (uint8_t)am : (uint8_t)CODE_ILLEGAL;
    }
    ix.n_last = rows[d->r - 1].y & 0x7FFu;
    std::vector<unsigned long long> keys; std::vector<uint16_t> vals;
    if (sep) {
        std::vector<std::pair<uint64_t, uint64_t>> kv(d->sep_map_size);
        for (uint64_t i = 0; i < d->sep_map_size; i++) kv[i] = {d->separators_map[2 * i], d->separators_map[2 * i + 1]};
        std::sort(kv.begin(), kv.end());
        keys.assign(kv.size() + 1, 0); vals.assign((kv.size() + 1) * 4, 0);
        for (size_t i = 0; i < kv.size(); i++) { keys[i] = kv[i].first; for (int k = 0; k < 4; k++) vals[i * 4 + k] = d->separators_thresholds[kv[i].second * 4 + k]; }
        if (d->sep_count) for (int k = 0; k < 4; k++) vals[kv.size() * 4 + k] = d->separators_thresholds[k];
        ix.sep_keys = keys.data(); ix.sep_vals = vals.data(); ix.sep_cnt = (uint32_t)kv.size();
    }

    unsigned long long counter = 0; unsigned int status = 0;
    std::vector<double> lk;
    std::vector<uint8_t> cnt;
    ExtArgs a{};
    a.bases = reinterpret_cast<const uint8_t*>(bases) + offsets[0]; a.offsets = reinterpret_cast<const unsigned long long*>(offsets);
    a.origin = offsets[0]; a.n_reads = (uint32_t)n_reads;
    a.out_ml = out_ml ? out_ml + offsets[0] : nullptr; a.counter = &counter; a.status = &status;
    if (ext & EXT_COLOR) {
        if (!col) return -1;
        a.color_inds = static_cast<const uint8_t*>(col->doc_set_flat_inds); a.flat = col->flat_colors; a.flat_size = col->flat_size;
        a.num_species = col->num_species; a.min_len = min_len;
        if (ext & EXT_PVALUE) { const double log4 = std::log(4.0); lk.resize(col->num_species); for (uint32_t i = 0; i < col->num_species; i++) lk[i] = col->log_lens[i] / log4; a.lk = lk.data(); }
        cnt.assign((size_t)lanes * col->num_species * 8, 0xAB);          // (dirty on purpose: a read must clear its own counters)
        a.cnt = cnt.data(); a.cnt_stride = lanes; a.out_color = out_color;
    }
    if (ext & EXT_SA) { a.all_p = reinterpret_cast<const unsigned long long*>(all_p); a.sampled = reinterpret_cast<const unsigned long long*>(sampled); a.rate = rate; a.out_sa = reinterpret_cast<unsigned long long*>(out_sa) + offsets[0]; }
    const bool wide = force_wide != 0;
#define PICK(EXTV) do { if (wide) { if (sep) run<unsigned long long, true, EXTV>(ix, a, lanes); else run<unsigned long long, false, EXTV>(ix, a, lanes); } \
                        else      { if (sep) run<uint32_t, true, EXTV>(ix, a, lanes); else run<uint32_t, false, EXTV>(ix, a, lanes); } } while (0)
    if (ext == EXT_COLOR) PICK(EXT_COLOR);
    else if (ext == (EXT_COLOR | EXT_PVALUE)) PICK(EXT_COLOR | EXT_PVALUE);
    else if (ext == EXT_SA) PICK(EXT_SA);
    else return -1;
#undef PICK
    return (int)status;
}

// ---- the --filter echo (movi_b200/csrc/filter_kernels.cuh): sizes, positions and the per-read copy, a read at a time
#include "../../movi_b200/csrc/filter_kernels.cuh"
extern "C" uint64_t movi_filter_emulate(const uint32_t* above, const uint32_t* below, int invert, const char* bases, const uint64_t* offsets,
                                        const char* ids, const uint64_t* id_offsets, uint64_t n_reads, char* text, uint64_t capacity) {
    std::vector<unsigned long long> pos(n_reads + 1, 0);
    for (uint64_t i = 0; i < n_reads; i++)
        pos[i + 1] = pos[i] + (filter_keeps(above[i], below[i], invert) ? filter_echo_bytes(id_offsets[i + 1] - id_offsets[i], offsets[i + 1] - offsets[i]) : 0ull);
    if (pos[n_reads] > capacity) return ~0ull;
    for (uint64_t i = 0; i < n_reads; i++) {
        if (pos[i + 1] == pos[i]) continue;
        for (unsigned lane = 0; lane < 3; lane++)                       // three "lanes" taking turns: the strided copy loops
            filter_emit_read(reinterpret_cast<uint8_t*>(text) + pos[i], reinterpret_cast<const uint8_t*>(ids) + id_offsets[i], id_offsets[i + 1] - id_offsets[i],
                             reinterpret_cast<const uint8_t*>(bases) + offsets[i], offsets[i + 1] - offsets[i], lane, 3u);
    }
    return pos[n_reads];
}
