// reference_bridge.cpp -- the REFERENCE-SIDE binding of include/movi_gpu.h, as a maintainer of Movi would add it
// (src/movi_gpu_bridge.cpp in the reference tree, linked with -lmovi_gpu): everything INTEGRATION.md section B shows, in one
// translation unit that the CPU test run compiles against the reference's own headers (tests/test_integration_bridge.py; where
// /root/reference is mounted) for MODE 6, 8 and 7 -- so the field names, types and MODE macros the binding relies on are checked,
// not just described.  It is not part of this repository's product and is never linked here.
//
// MoveStructure's tables are private (include/move_structure.hpp:267); ReadProcessor reaches them as a friend (:266).  The one
// line the reference's header gains is the same for the bridge:      friend struct MoviGpuBridge;
// (this file gets that access with the usual test-only define below, so that the reference's header stays untouched).
#include <bits/stdc++.h>          // (every standard header first: the define below must not reach them)
#include <omp.h>
#include <zlib.h>
#include "sdsl/bit_vectors.hpp"
#define private public
#include "move_structure.hpp"
#undef private
#include "batch_loader.hpp"
#include "movi_gpu.h"

struct MoviGpuBridge {
    movi_gpu_ctx* ctx = nullptr;
    std::vector<uint32_t> id_blocks_flat;          // id_blocks / tally_ids are vectors of vectors in the reference: flattened character-major
    std::vector<MoveTally> tally_flat;
    std::vector<uint64_t> sep_pairs;

    // once per process and GPU, after mv_.deserialize()  (src/movi.cpp:661-667)
    void attach(MoveStructure& mv, int device) {
        MoviGpuIndexDesc d{};
        d.mode = MODE;                                                             // include/utils.hpp:133-147
        d.length = mv.length; d.r = mv.r; d.end_bwt_idx = mv.end_bwt_idx;
        d.rows = mv.rlbwt.data();                                                  // packed MoveRow, 8 / 6 / 3 bytes
        d.sigma = (uint32_t)mv.alphabet.size();
        std::copy(mv.alphabet.begin(), mv.alphabet.end(), d.alphabet);
        d.alphamap = mv.alphamap.data();
        std::copy(mv.end_bwt_idx_thresholds, mv.end_bwt_idx_thresholds + 4, d.end_bwt_idx_thresholds);
        for (size_t k = 0; k < mv.first_runs.size() && k < 8; k++) {
            d.first_runs[k] = mv.first_runs[k]; d.first_offsets[k] = mv.first_offsets[k];
            d.last_runs[k] = mv.last_runs[k];   d.last_offsets[k] = mv.last_offsets[k];
        }
#if BLOCKED_MODES
        for (auto& b : mv.id_blocks) id_blocks_flat.insert(id_blocks_flat.end(), b.begin(), b.end());
        d.id_blocks = id_blocks_flat.data(); d.id_blocks_size = mv.id_blocks.empty() ? 0 : mv.id_blocks[0].size(); d.block_size = mv.block_size;
#endif
#if TALLY_MODES
        for (auto& t : mv.tally_ids) tally_flat.insert(tally_flat.end(), t.begin(), t.end());
        d.tally_ids = tally_flat.data(); d.tally_len = mv.tally_ids.empty() ? 0 : mv.tally_ids[0].size(); d.tally_checkpoints = mv.tally_checkpoints;
#endif
        if (mv.use_separator()) {                                                  // src/move_structure_io.cpp:415-433
            static_assert(sizeof(ThresholdsRow) == 4 * sizeof(uint16_t), "ThresholdsRow = 4 x u16");
            d.separators_thresholds = reinterpret_cast<const uint16_t*>(mv.separators_thresholds.data());
            d.sep_count = mv.separators_thresholds.size();
            for (auto& kv : mv.separators_thresholds_map) { sep_pairs.push_back(kv.first); sep_pairs.push_back(kv.second); }
            d.separators_map = sep_pairs.data(); d.sep_map_size = sep_pairs.size() / 2;
        }
        if (movi_gpu_create(device, &ctx) || movi_gpu_upload_index(ctx, &d, MOVI_GPU_LAYOUT_AUTO))
            throw std::runtime_error(ERROR_MSG(movi_gpu_last_error(ctx)));
    }

    // after load_color_table(mv_, movi_options)  (src/movi.cpp:121-152, flat colour table)
    void attach_colors(MoveStructure& mv) {
        static_assert(sizeof(MoveTally) == 5, "doc_set_flat_inds entries are 5 bytes");
        MoviGpuColorDesc c{};
        c.flat_colors = mv.flat_colors.data(); c.flat_size = mv.flat_colors.size();
        c.doc_set_flat_inds = mv.doc_set_flat_inds.data();
        c.num_species = mv.num_species; c.log_lens = mv.log_lens.data();
        if (movi_gpu_upload_colors(ctx, &c)) throw std::runtime_error(ERROR_MSG(movi_gpu_last_error(ctx)));
    }

    // after mv_.deserialize_sampled_SA()  (src/movi.cpp:669-675)
    void attach_sampled_sa(MoveStructure& mv) {
        if (movi_gpu_upload_sampled_sa(ctx, mv.movi_options->get_SA_sample_rate(), mv.sampled_SA_entries.data(), mv.sampled_SA_entries.size(), mv.all_p.data()))
            throw std::runtime_error(ERROR_MSG(movi_gpu_last_error(ctx)));
    }

    // one batch of reads instead of the per-read loop of query() (src/movi.cpp:317-383): the MoveQuery objects come back filled
    // exactly as query_pml / query_zml / query_backward_search fill them, for the unchanged writers and classifier
    void query_batch(MoveStructure& mv, BatchLoader& reader, std::vector<MoveQuery>& out, std::vector<uint64_t>& counts, std::vector<int32_t>& pos_on_r) {
        MoviOptions& opt = *mv.movi_options;
        std::string bases;                      // (movi_gpu_host_alloc'ed buffers copy at DMA speed; pageable ones are staged by the library)
        std::vector<uint64_t> off{0};
        std::vector<std::string> ids;
        Read read_struct;
        while (reader.grabNextRead(read_struct)) {
            std::string s = read_struct.seq;
            if (opt.is_reverse()) std::reverse(s.begin(), s.end());
            bases += s; off.push_back(bases.size()); ids.push_back(read_struct.id);
        }
        const uint64_t n = ids.size();
        std::vector<uint16_t> ml(bases.size() + 1);
        const uint32_t fl = opt.is_random_repositioning() ? MOVI_GPU_FLAG_RPML : MOVI_GPU_FLAG_NONE;
        int rc;
        if (opt.is_pml()) rc = movi_gpu_query_pml(ctx, bases.data(), off.data(), n, fl, ml.data());
        else if (opt.is_zml()) rc = movi_gpu_query_zml(ctx, bases.data(), off.data(), n, 0, ml.data());
        else { counts.resize(n); pos_on_r.resize(n); rc = movi_gpu_query_count(ctx, bases.data(), off.data(), n, 0, counts.data(), pos_on_r.data()); }
        if (rc) throw std::runtime_error(ERROR_MSG(movi_gpu_last_error(ctx)));      // same exit path as src/movi.cpp:744-747
        out.clear();
        for (uint64_t i = 0; i < n; i++) {
            std::string seq = bases.substr(off[i], off[i + 1] - off[i]);
            out.emplace_back(seq);
            out.back().set_query_id(ids[i]);
            if (opt.is_pml() || opt.is_zml())
                for (uint64_t k = off[i]; k < off[i + 1]; k++) out.back().add_ml(ml[k], opt.is_stdout());      // storage order is identical
        }
    }

    // --multi-classify: the line query_pml writes to output_files->out_file (src/move_structure_query.cpp:439-471)
    void multi_classify_batch(MoveStructure& mv, const std::string& bases, const std::vector<uint64_t>& off, const std::vector<std::string>& ids) {
        MoviOptions& opt = *mv.movi_options;
        std::vector<MoviGpuColorResult> res(ids.size());
        if (movi_gpu_query_multi_classify(ctx, bases.data(), off.data(), ids.size(), 0, (uint32_t)opt.get_min_match_len(), opt.is_pvalue_scoring(), res.data(), nullptr))
            throw std::runtime_error(ERROR_MSG(movi_gpu_last_error(ctx)));
        for (size_t i = 0; i < ids.size(); i++) {
            const MoviGpuColorResult& q = res[i];
            mv.output_files->out_file << ids[i] << ",";
            const float PML_mean = static_cast<float>(q.sum_matching_lengths) / (off[i + 1] - off[i]);
            if (PML_mean < UNCLASSIFIED_THRESHOLD || q.best_doc == std::numeric_limits<uint16_t>::max()) { mv.output_files->out_file << "0,0\n"; continue; }
            if (q.second_best_doc == std::numeric_limits<uint16_t>::max()) { mv.output_files->out_file << mv.to_taxon_id[q.best_doc] << ",0\n"; continue; }
            float best_cnt = 0, second_cnt = 0;                                   // classify_cnts or doc_scores, whichever the flags filled (:452-461)
            if (opt.get_min_match_len() > 0 && !opt.is_pvalue_scoring()) { best_cnt = (float)(uint32_t)q.best_raw; second_cnt = (float)(uint32_t)q.second_raw; }
            else if (opt.get_min_match_len() == 0 && opt.is_pvalue_scoring()) {
                double b, s; std::memcpy(&b, &q.best_raw, 8); std::memcpy(&s, &q.second_raw, 8);
                best_cnt = (float)b; second_cnt = (float)s;
            }
            const float diff = best_cnt - second_cnt;
            mv.output_files->out_file << mv.to_taxon_id[q.best_doc] << "," << (diff < 0.05 * best_cnt ? mv.to_taxon_id[q.second_best_doc] : 0u) << "\n";
        }
    }

    // --sa-entries: mq.add_sa_entries per base (src/move_structure_query.cpp:353-357)
    void sa_entries_batch(const std::string& bases, const std::vector<uint64_t>& off, std::vector<MoveQuery>& mqs) {
        std::vector<uint16_t> ml(bases.size() + 1); std::vector<uint64_t> sa(bases.size() + 1);
        if (movi_gpu_query_pml_sa(ctx, bases.data(), off.data(), mqs.size(), 0, ml.data(), sa.data()))
            throw std::runtime_error(ERROR_MSG(movi_gpu_last_error(ctx)));
        for (size_t i = 0; i < mqs.size(); i++)
            for (uint64_t k = off[i]; k < off[i + 1]; k++) { mqs[i].add_ml(ml[k], false); mqs[i].add_sa_entries(sa[k]); }
    }

    ~MoviGpuBridge() { if (ctx) movi_gpu_destroy(ctx); }
};
