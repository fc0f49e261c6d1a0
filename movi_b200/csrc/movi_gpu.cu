// movi_gpu.cu -- C ABI (include/movi_gpu.h) over the sm_100a query kernels.
//
// Owns: the device context, the upload / re-layout of an index into HBM, the pipelined host-buffer
// entry points (H2D copy -> kernel -> D2H copy over three streams) and the device-buffer entry points.
// There is deliberately no CPU fallback anywhere in this file.
#include "../../include/movi_gpu.h"

#include <cuda.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "bs_kernel.cuh"
#include "convert.cuh"
#include "dev_index.cuh"
#include "ext_kernel.cuh"
#include "filter_kernels.cuh"
#include "kernel_common.cuh"
#include "misc_kernels.cuh"
#include "parse_kernels.cuh"
#include "pml_kernel.cuh"

using namespace movi;

namespace {

constexpr int NBUF = 8;                       // slices in flight in the host-buffer entry points
constexpr int NKS = 4;                        // kernel streams the slices rotate over
std::string g_create_error;

struct Staging {                              // one pipeline slot
    cudaEvent_t in_done = nullptr, k_done = nullptr, out_done = nullptr;
    uint8_t* d_bases = nullptr; size_t bases_cap = 0;
    unsigned long long* d_offsets = nullptr; size_t offsets_cap = 0;   // in elements
    uint16_t* d_ml = nullptr; size_t ml_cap = 0;                        // in elements
    uint8_t* d_ml8 = nullptr; size_t ml8_cap = 0;                       // movi_gpu_query_ml8: the slice's lengths as bytes
    unsigned long long* d_count = nullptr; int32_t* d_pos = nullptr; size_t count_cap = 0;
    uint32_t* d_above = nullptr; uint32_t* d_below = nullptr; unsigned long long* d_sum = nullptr; size_t cls_cap = 0;
    uint8_t* d_ids = nullptr; size_t ids_cap = 0;                       // movi_gpu_query_bpf: the slice's read ids,
    unsigned long long* d_id_offsets = nullptr; size_t id_offsets_cap = 0;   // their offsets,
    uint8_t* d_image = nullptr; size_t image_cap = 0;                   // and its .bpf records
    unsigned long long* d_counter = nullptr;
    // pageable caller buffers: the slice's bases / results pass through these pinned arenas (a copy to or from pageable memory
    // is staged by the driver and makes the call wait for the stream: the pipeline would run slice by slice)
    uint8_t* h_in = nullptr; size_t h_in_cap = 0;
    uint8_t* h_out = nullptr; size_t h_out_cap = 0;
    struct Pending { void* dst; const void* src; size_t n; } pend[8];
    int n_pend = 0;
    bool busy = false;
};

}  // namespace

struct movi_gpu_ctx {
    int device = 0;
    int num_sms = 0;
    std::string err;
    // index
    bool has_index = false;
    bool borrowed_index = false;      // movi_gpu_share_index: the device tables belong to another context of the same GPU
    DevIndex ix{};
    uint2* d_rows = nullptr;
    uint32_t* d_rows6 = nullptr;      // LAYOUT_B6: packed blocked-thresholds rows
    uint32_t* d_id_blocks = nullptr;
    uint32_t* d_rows4 = nullptr;      // LAYOUT_T4: sampled-thresholds rows, 4 bytes each (row + walk-back distance)
    uint8_t* d_tally = nullptr;
    unsigned long long* d_sep_keys = nullptr; uint16_t* d_sep_vals = nullptr;
    unsigned long long* d_prefix = nullptr;
    uint64_t index_bytes = 0;
    int mode = 0;
    // query plumbing
    unsigned int* d_status = nullptr;
    unsigned long long* d_counters = nullptr;        // ring of work-queue heads for the *_device calls
    uint32_t counter_slot = 0;
    Staging slot[NBUF];
    // host-buffer pipeline: every H2D copy goes through st_in and every D2H copy through st_out, in slice order, so that a
    // copy-back waiting for its kernel never sits in front of the next slice's upload in a copy queue (per-slot streams did
    // exactly that: gpurun_out/trace_e2e.log of round 1); kernels rotate over st_k so that consecutive slices overlap
    cudaStream_t st_in = nullptr, st_out = nullptr, st_k[NKS] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t launches = 0;
    bool always_rpml = false;         // non-threshold index type: every PML query repositions with the --rpml rule
    uint32_t call_flags = 0;          // MOVI_GPU_FLAG_* of the entry point that is running
    uint64_t batch_reads = 0;         // host-buffer entry points: reads of the whole batch the slice being launched belongs to
    // options
    int64_t pml_sector = -1;          // sector-cached row fetch: -1 by table size (> 512 MB) and batch size, 0 off, 1 on
    int64_t smem_carveout = -1;       // >= 0: preferred shared-memory carve-out of the query kernels, percent of the maximum (exploration)
    int64_t out_stage = 32;           // sector-cached PML kernel: matching lengths staged per lane before a flush (16 or 32)
    int64_t force_wide = 0;           // tests: run the 64-bit-index instantiations on a small table
    int64_t blocks_per_sm = 0;        // 0: occupancy query
    int64_t threads_per_block = 256;
    int64_t batch_bases = 12ll << 20; // bases per pipeline slice in the host-buffer entry points (D2H of the results is the
                                      // bottleneck of the host path: small slices keep the pipeline ramp short)
    int64_t kernel_streams = 4;       // 1..NKS: slices whose kernels may run concurrently (sweep: profiles/r01_pipeline_sweep.jsonl)
    int64_t pipeline_slots = NBUF;    // 2..NBUF
    int64_t size_grid_to_batch = 1;   // 0: always launch the full persistent grid; k >= 1: a grid that gives every lane k reads of the batch
    // device-side ingest (movi_gpu_query_file_chunk): raw chunk, parse scratch, parsed reads, lengths, .bpf image
    uint8_t* p_raw = nullptr; size_t p_raw_cap = 0;
    uint32_t* p_tile = nullptr; size_t p_tile_cap = 0;
    uint32_t* p_nlpos = nullptr; size_t p_nlpos_cap = 0;
    LineInfo* p_lines = nullptr; size_t p_lines_cap = 0;
    Sum3* p_ltile = nullptr; size_t p_ltile_cap = 0;
    ParseResult* p_res = nullptr;                      // + one Sum3 + one u32 behind it
    uint8_t* p_bases = nullptr; size_t p_bases_cap = 0;
    unsigned long long* p_offsets = nullptr; size_t p_offsets_cap = 0;
    uint8_t* p_ids = nullptr; size_t p_ids_cap = 0;
    unsigned long long* p_id_offsets = nullptr; size_t p_id_offsets_cap = 0;
    uint16_t* p_ml = nullptr; size_t p_ml_cap = 0;
    uint8_t* p_image = nullptr; size_t p_image_cap = 0;
    ParseResult p_last{};                              // result of the last movi_gpu_file_chunk_parse (what _query works on)
    unsigned long long* p_pos = nullptr; size_t p_pos_cap = 0; uint64_t p_text_bytes = 0;      // movi_gpu_file_chunk_filter: echo positions; the text sits in p_image
    uint32_t* p_above = nullptr; uint32_t* p_below = nullptr; unsigned long long* p_sum = nullptr; size_t p_cls_cap = 0;   // movi_gpu_file_chunk_classify
    // pinned staging ring for copies from / to pageable memory and files (a file mapping is pageable by nature)
    uint8_t* pin[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pin_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // Movi Color: the flat colour table, the sampled suffix array, scratch of the ext queries (ext_kernel.cuh)
    uint8_t* c_inds = nullptr; uint16_t* c_flat = nullptr; double* c_lk = nullptr; uint64_t c_flat_size = 0; uint32_t c_species = 0; bool c_pvalue_ok = false;
    unsigned long long* s_all_p = nullptr; unsigned long long* s_sampled = nullptr; uint64_t s_rate = 0, s_n = 0;
    uint8_t* e_bases = nullptr; size_t e_bases_cap = 0;
    unsigned long long* e_offsets = nullptr; size_t e_offsets_cap = 0;
    uint16_t* e_ml = nullptr; size_t e_ml_cap = 0;
    unsigned long long* e_sa = nullptr; size_t e_sa_cap = 0;
    MoviGpuColorResult* e_res = nullptr; size_t e_res_cap = 0;
    uint8_t* e_cnt = nullptr; size_t e_cnt_cap = 0;
    // gather probe scratch
    uint2* d_probe = nullptr; uint64_t probe_bytes = 0;
    int64_t probe_vmm_align_mb = 0;   // > 0: allocate the probe table through the VMM API at this alignment
    unsigned long long probe_vmm[3] = {0, 0, 0};   // VmmBlock {ptr, size, handle} of the probe table
};

namespace {

int fail(movi_gpu_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, MOVI_GPU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// Device memory through the virtual-memory-management driver API (entry points resolved at run time, so the
// library does not link against libcuda): lets the row table be mapped at a large alignment, which the driver may back
// with larger page-table entries than cudaMalloc's 2 MB pages -- the reach of the GPU TLBs is what bounds random row
// fetches on tables beyond ~1 GB (profiles/r01_summary.md).
struct VmmBlock { CUdeviceptr ptr = 0; size_t size = 0; CUmemGenericAllocationHandle handle = 0; };

template <typename F> bool drv(const char* name, F* fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) { cudaGetLastError(); return false; }
    *fn = reinterpret_cast<F>(p);
    return true;
}

bool vmm_alloc(int device, size_t bytes, size_t align, VmmBlock* out) {
    CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*addrReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*getGran)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    if (!drv("cuMemCreate", &memCreate) || !drv("cuMemAddressReserve", &addrReserve) || !drv("cuMemMap", &memMap) ||
        !drv("cuMemSetAccess", &setAccess) || !drv("cuMemGetAllocationGranularity", &getGran)) return false;
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    size_t gran = 0;
    if (getGran(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) return false;
    const size_t unit = std::max(gran, align);
    const size_t size = (bytes + unit - 1) / unit * unit;
    CUmemGenericAllocationHandle h;
    if (memCreate(&h, size, &prop, 0) != CUDA_SUCCESS) return false;
    CUdeviceptr ptr = 0;
    if (addrReserve(&ptr, size, align, 0, 0) != CUDA_SUCCESS) return false;
    if (memMap(ptr, size, 0, h, 0) != CUDA_SUCCESS) return false;
    CUmemAccessDesc acc = {};
    acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (setAccess(ptr, size, &acc, 1) != CUDA_SUCCESS) return false;
    out->ptr = ptr; out->size = size; out->handle = h;
    return true;
}
void vmm_free(VmmBlock* b) {
    if (!b->ptr) return;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr; CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*addrFree)(CUdeviceptr, size_t) = nullptr;
    if (drv("cuMemUnmap", &unmap) && drv("cuMemRelease", &release) && drv("cuMemAddressFree", &addrFree)) {
        unmap(b->ptr, b->size); release(b->handle); addrFree(b->ptr, b->size);
    }
    *b = VmmBlock{};
}

template <typename K>
int blocks_for(movi_gpu_ctx* ctx, K kernel, int threads, int cap_per_sm = 0) {
    int per_sm = 0;
    if (ctx->blocks_per_sm > 0) per_sm = (int)ctx->blocks_per_sm;
    else {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        if (cap_per_sm > 0) per_sm = std::min(per_sm, cap_per_sm);
    }
    return per_sm * ctx->num_sms;
}

void free_index(movi_gpu_ctx* ctx) {
    if (ctx->borrowed_index) {           // nothing here is ours, except a prefix table built for the count query
        ctx->d_rows = nullptr; ctx->d_rows6 = nullptr; ctx->d_id_blocks = nullptr; ctx->d_rows4 = nullptr; ctx->d_tally = nullptr;
        ctx->d_sep_keys = nullptr; ctx->d_sep_vals = nullptr;
        cudaFree(ctx->d_prefix); ctx->d_prefix = nullptr;
        ctx->borrowed_index = false; ctx->has_index = false; ctx->index_bytes = 0;
        return;
    }
    cudaFree(ctx->d_rows); ctx->d_rows = nullptr;
    cudaFree(ctx->d_rows6); ctx->d_rows6 = nullptr;
    cudaFree(ctx->d_id_blocks); ctx->d_id_blocks = nullptr;
    cudaFree(ctx->d_rows4); ctx->d_rows4 = nullptr;
    cudaFree(ctx->d_tally); ctx->d_tally = nullptr;
    cudaFree(ctx->d_sep_keys); ctx->d_sep_keys = nullptr;
    cudaFree(ctx->d_sep_vals); ctx->d_sep_vals = nullptr;
    cudaFree(ctx->d_prefix); ctx->d_prefix = nullptr;
    ctx->has_index = false; ctx->index_bytes = 0;
}

int ensure_prefix(movi_gpu_ctx* ctx, cudaStream_t st) {
    if (ctx->d_prefix) return MOVI_GPU_OK;
    const unsigned long long r = ctx->ix.r;
    const unsigned long long n_groups = (r >> PREFIX_SHIFT) + 2;
    const unsigned long long n_blocks = (n_groups + SCAN_BLOCK - 1) / SCAN_BLOCK;
    unsigned long long* d_block = nullptr;
    CU(cudaMalloc(&ctx->d_prefix, n_groups * sizeof(unsigned long long)));
    CU(cudaMalloc(&d_block, n_blocks * sizeof(unsigned long long)));
    prefix_group_sums_kernel<<<(unsigned)n_blocks, SCAN_BLOCK, 0, st>>>(ctx->ix, r, n_groups, ctx->d_prefix, d_block);
    prefix_scan_blocks_kernel<<<1, SCAN_BLOCK, 0, st>>>(d_block, n_blocks);
    prefix_finish_kernel<<<(unsigned)n_blocks, SCAN_BLOCK, 0, st>>>(ctx->d_prefix, n_groups, d_block);
    ctx->launches += 3;
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    cudaFree(d_block);
    ctx->ix.prefix = ctx->d_prefix;
    ctx->index_bytes += n_groups * sizeof(unsigned long long);
    return MOVI_GPU_OK;
}

enum Kind { Q_PML = 0, Q_ZML = 1, Q_COUNT = 2 };

// enqueue one query kernel on `st` (device pointers)
int launch_query(movi_gpu_ctx* ctx, Kind kind, const QueryArgs& qa, cudaStream_t st) {
    const int threads = (int)ctx->threads_per_block;
    // a small batch (a pipeline slice) gets a grid sized to one read per lane instead of the whole machine: full warps
    // issue a quarter of the instructions of warps holding a handful of reads each, and leave SM room for the kernels of
    // the neighbouring slices
    const uint64_t per_block = (uint64_t)threads * (uint64_t)std::max<int64_t>(1, ctx->size_grid_to_batch);
    const uint64_t grid_cap = ctx->size_grid_to_batch ? std::max<uint64_t>(1, ((uint64_t)qa.n_reads + per_block - 1) / per_block) : ~0ull;
    const bool wide = ctx->ix.r > 0xFFFFFFF0ull || ctx->force_wide;   // 64-bit row indices (tables beyond 2^32 rows = 34 GB)
#define LAUNCH(KERNEL)                                                            \
    do {                                                                          \
        auto k_ = KERNEL;                                                         \
        if (ctx->smem_carveout >= 0) cudaFuncSetAttribute(k_, cudaFuncAttributePreferredSharedMemoryCarveout, (int)ctx->smem_carveout); \
        cudaLaunchConfig_t cfg_ = {};                                             \
        cfg_.gridDim = dim3((unsigned)std::min<uint64_t>((uint64_t)blocks_for(ctx, k_, threads, sm_cap), grid_cap)); \
        cfg_.blockDim = dim3((unsigned)threads);                                  \
        cfg_.stream = st;                                                         \
        CU(cudaLaunchKernelEx(&cfg_, k_, ctx->ix, qa));                           \
    } while (0)
    // tables far beyond the L2 are bound by the request rate of the memory system, not by issue: there the kernels keep the
    // last fetched 32-byte sector in registers ("pml_sector" -1 = by table size, 0 / 1 = off / on)
    // (measured, profiles/r01_big_tables.jsonl: 8 and 32 GB tables 1.2 -> 2.2-2.5e10 bases/s; the 124 MB c2 table -27 %;
    //  a batch of few long reads, which leaves most lanes empty and is latency bound, -10 %: those keep the plain fetch)
    const bool b6 = ctx->ix.layout == LAYOUT_B6, t3 = ctx->ix.layout == LAYOUT_T4;
    int sm_cap = 0;                     // resident blocks per SM at most (0 = what the registers allow)
    // (a slice of a bigger batch walks next to its neighbours: what counts is the reads of the batch, not of the slice -- slices of
    //  5.7e4 long reads ran the plain fetch until round 2's pipeline trace showed it: profiles/r02_e2e_pipeline_trace_before.txt)
    const uint64_t in_flight = std::max<uint64_t>(qa.n_reads, ctx->batch_reads);
    // crossover between the two fetch forms (profiles/r02_sweep_g1.jsonl, r02_sweep_1.jsonl): PML 124 MB table plain 2.11 / sector
    // 2.57 ms, 423 MB table 3.82 / 3.50 ms; ZML and count still prefer the plain fetch at 423 MB (5.63 / 6.18, 2.81 / 3.32 ms)
    const uint64_t sector_from = kind == Q_PML ? (256ull << 20) : (512ull << 20);
    const bool sector = !b6 && !t3 && !ctx->ix.nt && (ctx->pml_sector == 1 || (ctx->pml_sector < 0 &&
                        ctx->ix.r * sizeof(uint2) > sector_from && in_flight >= (1u << 17)));
    if (kind == Q_PML) {
        if (sector) sm_cap = 4;
        // 6 blocks/SM (40 registers) measured fastest on L2-sized tables (c2); the sector-cached kernel of HBM-resident tables runs
        // 4 blocks/SM: those tables are bound by the request rate of the memory system, and fewer walks in flight queue less
        // (36.8 GB table, 4.1e9 bases: 180 ms at 4 blocks/SM, 195 ms at 5, 193 ms at 3 -- profiles/r02_sweep_c4.jsonl)
        const bool rpml = (qa.flags & MOVI_GPU_FLAG_RPML) != 0;
#define PML(IDX, SEPV) do { \
            if (ctx->ix.nt) LAUNCH((movi_pml_kernel<IDX, SEPV, 6, LAYOUT_R8, false, true, true>));   /* non-threshold table: 12-bit accessors, always the --rpml rule (move_structure_query.cpp:308-311) */ \
            else if (rpml) { \
                if (t3) LAUNCH((movi_pml_kernel<IDX, SEPV, 5, LAYOUT_T4, false, true>)); \
                else if (b6) LAUNCH((movi_pml_kernel<IDX, SEPV, 6, LAYOUT_B6, false, true>)); \
                else LAUNCH((movi_pml_kernel<IDX, SEPV, 6, LAYOUT_R8, false, true>)); } \
            else if (t3) LAUNCH((movi_pml_kernel<IDX, SEPV, 5, LAYOUT_T4>)); \
            else if (b6) LAUNCH((movi_pml_kernel<IDX, SEPV, 6, LAYOUT_B6>)); \
            else if (sector && ctx->out_stage == 16) LAUNCH((movi_pml_kernel<IDX, SEPV, 4, LAYOUT_R8, true>)); \
            else if (sector) LAUNCH((movi_pml_kernel<IDX, SEPV, 4, LAYOUT_R8, true, false, false, 32>)); \
            else LAUNCH((movi_pml_kernel<IDX, SEPV, 6, LAYOUT_R8>)); } while (0)
        if (ctx->ix.sep) { if (wide) PML(unsigned long long, true); else PML(uint32_t, true); }
        else             { if (wide) PML(unsigned long long, false); else PML(uint32_t, false); }
#undef PML
    } else {
#define BS(IDX, KINDV) do { if (ctx->ix.nt) LAUNCH((movi_bs_kernel<IDX, KINDV, LAYOUT_R8, false, true>)); \
                            else if (b6) LAUNCH((movi_bs_kernel<IDX, KINDV, LAYOUT_B6>)); \
                            else if (sector) LAUNCH((movi_bs_kernel<IDX, KINDV, LAYOUT_R8, true>)); \
                            else LAUNCH((movi_bs_kernel<IDX, KINDV, LAYOUT_R8>)); } while (0)
        if (kind == Q_ZML) { if (wide) BS(unsigned long long, KIND_ZML); else BS(uint32_t, KIND_ZML); }
        else               { if (wide) BS(unsigned long long, KIND_COUNT); else BS(uint32_t, KIND_COUNT); }
#undef BS
    }
#undef LAUNCH
    ctx->launches += 1;
    CU(cudaGetLastError());
    return MOVI_GPU_OK;
}

int check_query_ready(movi_gpu_ctx* ctx, Kind kind, uint64_t n_reads) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    if (!ctx->has_index) return fail(ctx, MOVI_GPU_ERR_INDEX, "no index uploaded");
    if (n_reads >= 0xFFFFFFF0ull) return fail(ctx, MOVI_GPU_ERR_ARG, "too many reads in one device call (limit 2^32 - 16)");
    if (kind != Q_PML && (ctx->mode == MOVI_MODE_SAMPLED_THRESHOLDS || ctx->mode == MOVI_MODE_SAMPLED))
        // reference src/movi.cpp:228-233
        return fail(ctx, MOVI_GPU_ERR_INDEX, "This query type is not suppported by the sampled modes yet.");
    return MOVI_GPU_OK;
}

int device_query(movi_gpu_ctx* ctx, Kind kind, const char* d_bases, const uint64_t* d_offsets, uint64_t origin,
                 uint64_t n_reads, uint16_t* d_ml, uint64_t* d_count, int32_t* d_pos, unsigned long long* d_counter,
                 cudaStream_t st) {
    if (n_reads == 0) return MOVI_GPU_OK;
    if (reinterpret_cast<uintptr_t>(d_bases) & 31u) return fail(ctx, MOVI_GPU_ERR_ARG, "d_bases must be 32-byte aligned");
    if (d_ml && (reinterpret_cast<uintptr_t>(d_ml) & 31u)) return fail(ctx, MOVI_GPU_ERR_ARG, "d_out_ml must be 32-byte aligned");
    if (kind == Q_COUNT) { int rc = ensure_prefix(ctx, st); if (rc) return rc; }
    CU(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st));
    QueryArgs qa{};
    qa.bases = reinterpret_cast<const uint8_t*>(d_bases);
    qa.offsets = reinterpret_cast<const unsigned long long*>(d_offsets);
    qa.origin = origin;
    qa.n_reads = (uint32_t)n_reads;
    qa.out_ml = d_ml;
    qa.out_count = reinterpret_cast<unsigned long long*>(d_count);
    qa.out_pos = d_pos;
    qa.counter = d_counter;
    qa.status = ctx->d_status;
    qa.flags = kind == Q_PML ? (ctx->call_flags | (ctx->always_rpml ? MOVI_GPU_FLAG_RPML : 0u)) : 0u;
    return launch_query(ctx, kind, qa, st);
}

template <typename T>
int grow(movi_gpu_ctx* ctx, T*& p, size_t& cap, size_t need) {
    if (need <= cap) return MOVI_GPU_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = need + need / 8 + 64;
    CU(cudaMalloc(&p, want * sizeof(T)));
    cap = want;
    return MOVI_GPU_OK;
}

// ---- staged copies: pageable host memory (a memory-mapped file, a std::vector) and file descriptors on one side, HBM on the other.
// cudaMemcpyAsync on pageable memory is staged by the driver through one thread at ~6 GB/s and blocks; here the pieces go through
// a ring of four pinned buffers, filled / drained by a few host threads while the previous pieces are on the wire.
constexpr size_t PIN_PIECE = 64u << 20;
inline int pin_threads() {                           // host threads that fill / drain a piece (page-cache writes run at ~1 GB/s per thread)
    static const int n = (int)std::min(12u, std::max(2u, std::thread::hardware_concurrency() > 3 ? std::thread::hardware_concurrency() - 2 : 2u));
    return n;
}

int ensure_pin_ring(movi_gpu_ctx* ctx) {
    for (int k = 0; k < 4; k++) {
        if (!ctx->pin[k]) CU(cudaHostAlloc(reinterpret_cast<void**>(&ctx->pin[k]), PIN_PIECE, cudaHostAllocDefault));
        if (!ctx->pin_ev[k]) CU(cudaEventCreateWithFlags(&ctx->pin_ev[k], cudaEventDisableTiming));
    }
    return MOVI_GPU_OK;
}
bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
template <typename F> void in_parts(size_t n, F f) {          // f(begin, end) on pin_threads() threads
    const int t = (int)std::min<size_t>((size_t)pin_threads(), (n + (4u << 20) - 1) / (4u << 20));
    if (t <= 1) { f((size_t)0, n); return; }
    std::vector<std::thread> th;
    for (int k = 0; k < t; k++) th.emplace_back(f, n * k / t, n * (k + 1) / t);
    for (auto& x : th) x.join();
}
int staged_h2d(movi_gpu_ctx* ctx, void* d_dst, const char* src, size_t n, cudaStream_t st) {
    if (n == 0) return MOVI_GPU_OK;
    if (is_pinned(src)) { CU(cudaMemcpyAsync(d_dst, src, n, cudaMemcpyHostToDevice, st)); return MOVI_GPU_OK; }
    int rc = ensure_pin_ring(ctx);
    if (rc) return rc;
    size_t k = 0;
    for (size_t off = 0; off < n; off += PIN_PIECE, k++) {
        const size_t len = std::min(PIN_PIECE, n - off);
        const int slot = (int)(k & 3);
        // the copy that last read this piece must be over before it is filled again -- also a copy of an EARLIER call (two staged
        // uploads in a row: the second one's bytes would otherwise land in the first one's device buffer); an event never recorded
        // is complete
        CU(cudaEventSynchronize(ctx->pin_ev[slot]));
        uint8_t* buf = ctx->pin[slot];
        in_parts(len, [&](size_t b, size_t e) { std::memcpy(buf + b, src + off + b, e - b); });
        CU(cudaMemcpyAsync(static_cast<uint8_t*>(d_dst) + off, buf, len, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(ctx->pin_ev[slot], st));
    }
    return MOVI_GPU_OK;
}
// device -> host memory (dst != NULL) or -> file (fd >= 0, positional writes at file_off, or write() calls in order when file_off is
// SEQUENTIAL_WRITES); returns when everything has landed
constexpr uint64_t SEQUENTIAL_WRITES = ~0ull;
int staged_d2h(movi_gpu_ctx* ctx, uint8_t* dst, int fd, uint64_t file_off, const uint8_t* d_src, size_t n, cudaStream_t st) {
    if (n == 0) return MOVI_GPU_OK;
    if (dst && fd < 0 && is_pinned(dst)) { CU(cudaMemcpyAsync(dst, d_src, n, cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st)); return MOVI_GPU_OK; }
    int rc = ensure_pin_ring(ctx);
    if (rc) return rc;
    const size_t pieces = (n + PIN_PIECE - 1) / PIN_PIECE;
    size_t issued = 0;
    bool io_ok = true;
    uint64_t file_len = 0;
    if (fd >= 0 && file_off != SEQUENTIAL_WRITES) { const off_t e = lseek(fd, 0, SEEK_END); if (e > 0) file_len = (uint64_t)e; }
    for (size_t done = 0; done < pieces; done++) {
        for (; issued < pieces && issued - done < 4; issued++) {
            const size_t off = issued * PIN_PIECE, len = std::min(PIN_PIECE, n - off);
            CU(cudaMemcpyAsync(ctx->pin[issued & 3], d_src + off, len, cudaMemcpyDeviceToHost, st));
            CU(cudaEventRecord(ctx->pin_ev[issued & 3], st));
        }
        CU(cudaEventSynchronize(ctx->pin_ev[done & 3]));
        const size_t off = done * PIN_PIECE, len = std::min(PIN_PIECE, n - off);
        const uint8_t* buf = ctx->pin[done & 3];
        if (fd >= 0 && file_off == SEQUENTIAL_WRITES) {
            // a stream (stdout may be a pipe): plain write() calls, piece after piece
            size_t b = 0;
            while (b < len && io_ok) { const ssize_t w = ::write(fd, buf + b, len - b); if (w <= 0) io_ok = false; else b += (size_t)w; }
        } else if (fd >= 0) {
            // Into the page cache through a shared mapping when the file already reaches that far (the caller extends it): positional
            // writes of several threads to ONE file take turns on its inode lock (2.9 GB/s on tmpfs whatever the thread count),
            // page faults on a mapping do not.  Otherwise pwrite.
            const uint64_t at = file_off + off, page = 4096, m0 = at & ~(page - 1);
            const size_t mlen = (size_t)(at + len - m0);
            void* m = (file_len >= at + len) ? mmap(nullptr, mlen, PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)m0) : MAP_FAILED;
            if (m != MAP_FAILED) {
                uint8_t* to = static_cast<uint8_t*>(m) + (at - m0);
                in_parts(len, [&](size_t b, size_t e) { std::memcpy(to + b, buf + b, e - b); });
                munmap(m, mlen);
            } else
            in_parts(len, [&](size_t b, size_t e) {
                while (b < e) { const ssize_t w = pwrite(fd, buf + b, e - b, (off_t)(file_off + off + b)); if (w <= 0) { io_ok = false; return; } b += (size_t)w; }
            });
        }
        if (dst) in_parts(len, [&](size_t b, size_t e) { std::memcpy(dst + off + b, buf + b, e - b); });
    }
    if (!io_ok) return fail(ctx, MOVI_GPU_ERR_ARG, "write failed (output file / stream)");
    return MOVI_GPU_OK;
}

int grow_pinned(movi_gpu_ctx* ctx, uint8_t*& p, size_t& cap, size_t need) {
    if (need <= cap) return MOVI_GPU_OK;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    const size_t want = need + need / 8 + 4096;
    CU(cudaHostAlloc(reinterpret_cast<void**>(&p), want, cudaHostAllocDefault));
    cap = want;
    return MOVI_GPU_OK;
}
void flush_slot(Staging& S) {            // results of a finished slice: pinned arena -> the caller's (pageable) buffers
    for (int k = 0; k < S.n_pend; k++) {
        const Staging::Pending& q = S.pend[k];
        in_parts(q.n, [&](size_t b, size_t e) { std::memcpy(static_cast<uint8_t*>(q.dst) + b, static_cast<const uint8_t*>(q.src) + b, e - b); });
    }
    S.n_pend = 0;
}

// Host-buffer entry point: split the batch into slices of ~batch_bases, run them through a 3-deep
// pipeline (H2D on slot stream, kernel, D2H), results land directly in the caller's buffers.
struct ClassifyOut { uint32_t bin_width = 0, thr = 0; uint32_t* above = nullptr; uint32_t* below = nullptr; uint64_t* sum = nullptr; };
struct BpfOut { const char* ids = nullptr; const uint64_t* id_offsets = nullptr; uint8_t* image = nullptr; };
// byte position of read i's record in the .bpf image of a batch (both offset arrays are prefix sums)
inline uint64_t bpf_record_at(const uint64_t* offsets, const uint64_t* id_offsets, uint64_t i) {
    return 10ull * i + (id_offsets[i] - id_offsets[0]) + 2ull * (offsets[i] - offsets[0]);
}

int host_query(movi_gpu_ctx* ctx, Kind kind, const char* bases, const uint64_t* offsets, uint64_t n_reads,
               uint16_t* out_ml, uint64_t* out_count, int32_t* out_pos, const ClassifyOut* cls = nullptr, uint8_t* out_ml8 = nullptr,
               const BpfOut* bpf = nullptr) {
    int rc = check_query_ready(ctx, kind, 0);
    if (rc) return rc;
    if (n_reads == 0) return MOVI_GPU_OK;
    if (!bases || !offsets) return fail(ctx, MOVI_GPU_ERR_ARG, "null input buffer");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemset(ctx->d_status, 0, sizeof(unsigned int)));
    if (kind == Q_COUNT) { rc = ensure_prefix(ctx, ctx->st_k[0]); if (rc) return rc; }

    struct BatchReads { movi_gpu_ctx* c; BatchReads(movi_gpu_ctx* x, uint64_t n) : c(x) { c->batch_reads = n; } ~BatchReads() { c->batch_reads = 0; } } batch_reads(ctx, n_reads);
    const int n_slots = (int)std::min<int64_t>(std::max<int64_t>(ctx->pipeline_slots, 2), NBUF);
    const int n_ks = (int)std::min<int64_t>(std::max<int64_t>(ctx->kernel_streams, 1), NKS);
    // caller buffers in pageable memory (std::vector, numpy): stage through the slots' pinned arenas
    const void* main_out = kind == Q_COUNT ? static_cast<const void*>(out_count) : out_ml ? static_cast<const void*>(out_ml)
                         : out_ml8 ? static_cast<const void*>(out_ml8) : bpf ? static_cast<const void*>(bpf->image)
                         : cls ? static_cast<const void*>(cls->above) : nullptr;
    const bool stage_in = !is_pinned(bases), stage_out = main_out && !is_pinned(main_out);
    for (int k = 0; k < NBUF; k++) ctx->slot[k].n_pend = 0;         // (a call that failed half way may have left some)
    // a slice holds batch_bases bases, or -- reads longer than ~200 bases -- as many reads as one of the n_ks concurrently
    // walking slices has lanes: a read is one dependent chain, so a slice of few long reads would take the full walk latency
    // of a read however small it is (12 M bases = 1200 reads of 10 kbp: 4.7 s per 4.1e9-base batch instead of 0.3 s)
    // (lanes of the kernel the slices will run: the sector-cached forms keep 4 / 3 blocks per SM resident, the plain ones 6 / 5;
    //  slices sized for more lanes than there are spill into an extra round of walks: profiles/r02_e2e_pipeline_trace_after.txt)
    const bool big_table = ctx->ix.layout == LAYOUT_R8 && !ctx->ix.nt && (ctx->pml_sector == 1 || (ctx->pml_sector < 0 &&
                           ctx->ix.r * sizeof(uint2) > (kind == Q_PML ? (256ull << 20) : (512ull << 20)) && n_reads >= (1u << 17)));
    const unsigned blocks_per_sm = ctx->blocks_per_sm > 0 ? (unsigned)ctx->blocks_per_sm
                                 : kind == Q_PML ? (big_table ? 4u : 6u) : (big_table ? 3u : 5u);
    const uint64_t lanes = (uint64_t)ctx->num_sms * blocks_per_sm * (uint64_t)ctx->threads_per_block;
    const uint64_t avg_len = std::max<uint64_t>(1, (offsets[n_reads] - offsets[0]) / n_reads);
    // n_ks slices walk at a time; with twice the lanes' worth of reads among them, blocks of the next slice move in as blocks of an
    // earlier one drain (c4, 4.1e9 bases per call, profiles/r02_e2e_sweep.jsonl: 274 ms with slices of lanes / 2 reads, 329 ms with
    // lanes / 4, 289 ms with lanes / 1.4; two or one slices at a time: 322 - 452 ms)
    const uint64_t max_bases = std::min<uint64_t>(std::max<uint64_t>((uint64_t)std::max<int64_t>(ctx->batch_bases, 1 << 16),
                                                                     avg_len * (2 * lanes / (uint64_t)n_ks)), 1ull << 30);
    // MOVI_GPU_TRACE=1: per-slice device timeline (ms since the first copy was enqueued) on stderr -- a debugging aid for
    // the copy / kernel overlap of this pipeline, off the product path
    static const bool trace = getenv("MOVI_GPU_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };

    uint64_t i0 = 0, slice = 0;
    while (i0 < n_reads) {
        // slice [i0, i1): at least one read, at most max_bases bases (a single longer read gets its own slice); the first
        // slices are smaller so that the first results start travelling back early
        const uint64_t want = slice < 2 ? std::max<uint64_t>(max_bases >> (2 - slice), 1 << 16) : max_bases;
        uint64_t i1 = i0 + 1;
        {
            const uint64_t limit = offsets[i0] + want;
            // offsets is non-decreasing: binary search the last read that still fits
            uint64_t lo = i0 + 1, hi = std::min<uint64_t>(n_reads, i0 + 0x7FFFFFF0ull);
            while (lo < hi) { uint64_t mid = (lo + hi + 1) / 2; if (offsets[mid] <= limit) lo = mid; else hi = mid - 1; }
            i1 = std::max(i1, lo);
        }
        const uint64_t o0 = offsets[i0], o1 = offsets[i1];
        const uint64_t nb = o1 - o0, nr = i1 - i0;
        if (nb >= 0xFFFFFFF0ull) return fail(ctx, MOVI_GPU_ERR_ARG, "a single read longer than 2^32 bases is not supported");
        Staging& S = ctx->slot[slice % n_slots];
        cudaStream_t kst = ctx->st_k[slice % n_ks];
        if (S.busy) { CU(cudaEventSynchronize(S.out_done)); flush_slot(S); S.busy = false; }   // the slot's previous results have left
        // keep the device copy congruent to the host offsets modulo 32: the kernels stage whole aligned 32-byte sectors
        const uint64_t shift = o0 & 31u;
        if ((rc = grow(ctx, S.d_bases, S.bases_cap, nb + shift + 64))) return rc;
        if ((rc = grow(ctx, S.d_offsets, S.offsets_cap, nr + 1))) return rc;
        if (kind == Q_COUNT) {
            size_t cap2 = S.count_cap;
            if ((rc = grow(ctx, S.d_count, S.count_cap, nr))) return rc;
            if ((rc = grow(ctx, S.d_pos, cap2, nr))) return rc;
        } else {
            if ((rc = grow(ctx, S.d_ml, S.ml_cap, nb + shift + 64))) return rc;
        }
        if (out_ml8 && (rc = grow(ctx, S.d_ml8, S.ml8_cap, nb + shift + 64))) return rc;
        uint64_t img0 = 0, img_bytes = 0, id_bytes = 0;
        if (bpf) {
            img0 = bpf_record_at(offsets, bpf->id_offsets, i0);
            img_bytes = bpf_record_at(offsets, bpf->id_offsets, i1) - img0;
            id_bytes = bpf->id_offsets[i1] - bpf->id_offsets[i0];
            if ((rc = grow(ctx, S.d_ids, S.ids_cap, id_bytes + 16))) return rc;
            if ((rc = grow(ctx, S.d_id_offsets, S.id_offsets_cap, nr + 1))) return rc;
            if ((rc = grow(ctx, S.d_image, S.image_cap, img_bytes + 16))) return rc;
        }
        if (cls) {
            size_t c1 = S.cls_cap, c2 = S.cls_cap;
            if ((rc = grow(ctx, S.d_above, S.cls_cap, nr))) return rc;
            if ((rc = grow(ctx, S.d_below, c1, nr))) return rc;
            if ((rc = grow(ctx, S.d_sum, c2, nr))) return rc;
        }
        // in
        mark(ctx->st_in);
        const bool staged_in = stage_in && nb <= (64ull << 20), staged = stage_out && nb <= (64ull << 20);
        size_t out_at = 0;
        if (staged_in) {
            if ((rc = grow_pinned(ctx, S.h_in, S.h_in_cap, nb + 64))) return rc;
            in_parts(nb, [&](size_t b, size_t e) { std::memcpy(S.h_in + b, bases + o0 + b, e - b); });
        }
        if (staged) {
            const size_t need_out = (kind == Q_COUNT ? nr * 12 : 0) + (out_ml ? nb * 2 : 0) + (out_ml8 ? nb : 0) + (cls ? nr * 16 : 0) + (bpf ? img_bytes : 0) + 1024;
            if ((rc = grow_pinned(ctx, S.h_out, S.h_out_cap, need_out))) return rc;
        }
        // a result: straight to the caller's pinned buffer, or into the slot's pinned arena with the last hop deferred
        auto copy_back = [&](void* dst, const void* d_src, size_t n) -> cudaError_t {
            if (!staged) return cudaMemcpyAsync(dst, d_src, n, cudaMemcpyDeviceToHost, ctx->st_out);
            uint8_t* via = S.h_out + out_at;
            out_at += (n + 63) & ~(size_t)63;
            S.pend[S.n_pend++] = Staging::Pending{dst, via, n};
            return cudaMemcpyAsync(via, d_src, n, cudaMemcpyDeviceToHost, ctx->st_out);
        };
        CU(cudaMemcpyAsync(S.d_bases + shift, staged_in ? reinterpret_cast<const char*>(S.h_in) : bases + o0, nb, cudaMemcpyHostToDevice, ctx->st_in));
        CU(cudaMemcpyAsync(S.d_offsets, offsets + i0, (nr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->st_in));
        if (bpf) {
            CU(cudaMemcpyAsync(S.d_ids, bpf->ids + bpf->id_offsets[i0], id_bytes, cudaMemcpyHostToDevice, ctx->st_in));
            CU(cudaMemcpyAsync(S.d_id_offsets, bpf->id_offsets + i0, (nr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->st_in));
        }
        CU(cudaEventRecord(S.in_done, ctx->st_in));
        mark(ctx->st_in);
        // walk
        CU(cudaStreamWaitEvent(kst, S.in_done, 0));
        rc = device_query(ctx, kind, reinterpret_cast<const char*>(S.d_bases), reinterpret_cast<const uint64_t*>(S.d_offsets),
                          o0 - shift, nr, S.d_ml, reinterpret_cast<uint64_t*>(S.d_count), S.d_pos, S.d_counter, kst);
        if (rc) return rc;
        if (cls) {
            // classify the slice where it sits: only 16 bytes per read travel back instead of 2 bytes per base
            const unsigned blocks = (unsigned)std::min<uint64_t>((nr * 32 + 255) / 256, (uint64_t)ctx->num_sms * 16);
            classify_kernel<<<blocks, 256, 0, kst>>>(S.d_ml, S.d_offsets, o0 - shift, nr, cls->bin_width, cls->thr,
                                                    S.d_above, S.d_below, S.d_sum);
            ctx->launches += 1;
            CU(cudaGetLastError());
        }
        if (out_ml8) {
            const unsigned long long groups = (nb + shift + 15) / 16;
            const unsigned blocks = (unsigned)std::min<uint64_t>((groups + 255) / 256, (uint64_t)ctx->num_sms * 8);
            narrow_ml_kernel<<<blocks, 256, 0, kst>>>(reinterpret_cast<const uint4*>(S.d_ml), reinterpret_cast<uint4*>(S.d_ml8), groups);
            ctx->launches += 1;
            CU(cudaGetLastError());
        }
        if (bpf) {
            // the slice's .bpf records assembled where the lengths sit: one copy back, no per-read work on the host
            const unsigned blocks = (unsigned)std::min<uint64_t>((nr * 32 + 255) / 256, (uint64_t)ctx->num_sms * 16);
            bpf_assemble_kernel<<<blocks, 256, 0, kst>>>(S.d_ml, S.d_offsets, o0 - shift, nr, S.d_ids, S.d_id_offsets, S.d_image);
            ctx->launches += 1;
            CU(cudaGetLastError());
        }
        CU(cudaEventRecord(S.k_done, kst));
        mark(kst);
        // out
        CU(cudaStreamWaitEvent(ctx->st_out, S.k_done, 0));
        if (kind == Q_COUNT) {
            CU(copy_back(out_count + i0, S.d_count, nr * sizeof(uint64_t)));
            CU(copy_back(out_pos + i0, S.d_pos, nr * sizeof(int32_t)));
        } else {
            if (cls) {
                CU(copy_back(cls->above + i0, S.d_above, nr * sizeof(uint32_t)));
                CU(copy_back(cls->below + i0, S.d_below, nr * sizeof(uint32_t)));
                CU(copy_back(cls->sum + i0, S.d_sum, nr * sizeof(uint64_t)));
            }
            if (out_ml) CU(copy_back(out_ml + o0, S.d_ml + shift, nb * sizeof(uint16_t)));
            if (out_ml8) CU(copy_back(out_ml8 + o0, S.d_ml8 + shift, nb));
            if (bpf) CU(copy_back(bpf->image + (img0 - bpf_record_at(offsets, bpf->id_offsets, 0)), S.d_image, img_bytes));
        }
        CU(cudaEventRecord(S.out_done, ctx->st_out));
        mark(ctx->st_out);
        S.busy = true;
        slice += 1;
        i0 = i1;
    }
    CU(cudaStreamSynchronize(ctx->st_out));            // st_out runs in slice order: its end is the end of the batch
    for (int k = 0; k < NBUF; k++) { flush_slot(ctx->slot[k]); ctx->slot[k].busy = false; }
    if (trace) {
        for (size_t k = 0; k + 3 < tev.size(); k += 4) {
            float t[4];
            for (int j = 0; j < 4; j++) cudaEventElapsedTime(&t[j], tev[0], tev[k + j]);
            fprintf(stderr, "[movi_gpu trace] slice %2zu  h2d %.3f..%.3f  kernel ..%.3f  d2h ..%.3f ms\n", k / 4, t[0], t[1], t[2], t[3]);
        }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    return movi_gpu_check_device_status(ctx);
}

}  // namespace

extern "C" {

int movi_gpu_abi_version(void) { return MOVI_GPU_ABI_VERSION; }

int movi_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* movi_gpu_last_error(const movi_gpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int movi_gpu_create(int device, movi_gpu_ctx** out) {
    movi_gpu_ctx* ctx = nullptr;   // for the CU macro: errors go to g_create_error
    if (!out) return MOVI_GPU_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, MOVI_GPU_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                    "); this library has no CPU fallback");
    if (device < 0 || device >= n) return fail(nullptr, MOVI_GPU_ERR_ARG, "device ordinal out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    movi_gpu_ctx* c = new movi_gpu_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    ctx = c;
    cudaError_t e2 = cudaMalloc(&c->d_status, sizeof(unsigned int));
    if (e2 == cudaSuccess) e2 = cudaMemset(c->d_status, 0, sizeof(unsigned int));
    if (e2 == cudaSuccess) e2 = cudaMalloc(&c->d_counters, 64 * sizeof(unsigned long long));
    if (e2 == cudaSuccess) e2 = cudaStreamCreateWithFlags(&c->st_in, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaStreamCreateWithFlags(&c->st_out, cudaStreamNonBlocking);
    for (int k = 0; k < NKS && e2 == cudaSuccess; k++) e2 = cudaStreamCreateWithFlags(&c->st_k[k], cudaStreamNonBlocking);
    for (int k = 0; k < NBUF && e2 == cudaSuccess; k++) {
        e2 = cudaEventCreateWithFlags(&c->slot[k].in_done, cudaEventDisableTiming);
        if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&c->slot[k].k_done, cudaEventDisableTiming);
        if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&c->slot[k].out_done, cudaEventDisableTiming);
        if (e2 == cudaSuccess) e2 = cudaMalloc(&c->slot[k].d_counter, sizeof(unsigned long long));
    }
    if (e2 != cudaSuccess) {
        g_create_error = std::string("context setup failed: ") + cudaGetErrorString(e2);
        movi_gpu_destroy(c);
        return MOVI_GPU_ERR_CUDA;
    }
    *out = c;
    return MOVI_GPU_OK;
}

int movi_gpu_destroy(movi_gpu_ctx* ctx) {
    if (!ctx) return MOVI_GPU_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    free_index(ctx);
    for (int k = 0; k < NBUF; k++) {
        Staging& S = ctx->slot[k];
        cudaFree(S.d_bases); cudaFree(S.d_offsets); cudaFree(S.d_ml); cudaFree(S.d_ml8); cudaFree(S.d_count); cudaFree(S.d_pos); cudaFree(S.d_counter);
        cudaFree(S.d_above); cudaFree(S.d_below); cudaFree(S.d_sum);
        cudaFree(S.d_ids); cudaFree(S.d_id_offsets); cudaFree(S.d_image);
        if (S.h_in) cudaFreeHost(S.h_in);
        if (S.h_out) cudaFreeHost(S.h_out);
        if (S.in_done) cudaEventDestroy(S.in_done);
        if (S.k_done) cudaEventDestroy(S.k_done);
        if (S.out_done) cudaEventDestroy(S.out_done);
    }
    if (ctx->st_in) cudaStreamDestroy(ctx->st_in);
    if (ctx->st_out) cudaStreamDestroy(ctx->st_out);
    for (int k = 0; k < NKS; k++) if (ctx->st_k[k]) cudaStreamDestroy(ctx->st_k[k]);
    for (int k = 0; k < 4; k++) { if (ctx->pin[k]) cudaFreeHost(ctx->pin[k]); if (ctx->pin_ev[k]) cudaEventDestroy(ctx->pin_ev[k]); }
    cudaFree(ctx->p_raw); cudaFree(ctx->p_tile); cudaFree(ctx->p_nlpos); cudaFree(ctx->p_lines); cudaFree(ctx->p_ltile); cudaFree(ctx->p_res);
    cudaFree(ctx->p_bases); cudaFree(ctx->p_offsets); cudaFree(ctx->p_ids); cudaFree(ctx->p_id_offsets); cudaFree(ctx->p_ml); cudaFree(ctx->p_image);
    cudaFree(ctx->p_above); cudaFree(ctx->p_below); cudaFree(ctx->p_sum); cudaFree(ctx->p_pos);
    cudaFree(ctx->d_status); cudaFree(ctx->d_counters);
    cudaFree(ctx->c_inds); cudaFree(ctx->c_flat); cudaFree(ctx->c_lk); cudaFree(ctx->s_all_p); cudaFree(ctx->s_sampled);
    cudaFree(ctx->e_bases); cudaFree(ctx->e_offsets); cudaFree(ctx->e_ml); cudaFree(ctx->e_sa); cudaFree(ctx->e_res); cudaFree(ctx->e_cnt);
    if (ctx->probe_vmm[0]) { VmmBlock b; b.ptr = ctx->probe_vmm[0]; b.size = ctx->probe_vmm[1]; b.handle = ctx->probe_vmm[2]; vmm_free(&b); }
    else cudaFree(ctx->d_probe);
    delete ctx;
    return MOVI_GPU_OK;
}

uint64_t movi_gpu_index_device_bytes(const movi_gpu_ctx* ctx) { return ctx ? ctx->index_bytes : 0; }
const void* movi_gpu_index_device_rows(const movi_gpu_ctx* ctx) { return ctx && ctx->has_index ? ctx->d_rows : nullptr; }
int movi_gpu_index_device_layout(const movi_gpu_ctx* ctx) { return ctx && ctx->has_index ? (ctx->ix.layout != LAYOUT_R8 ? MOVI_GPU_LAYOUT_COMPACT : MOVI_GPU_LAYOUT_EXPANDED8) : -1; }
uint64_t movi_gpu_launch_count(const movi_gpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

int movi_gpu_set_option(movi_gpu_ctx* ctx, const char* key, int64_t value) {
    if (!ctx || !key) return MOVI_GPU_ERR_ARG;
    const std::string k(key);
    if (k == "force_wide") { ctx->force_wide = value ? 1 : 0; }
    else if (k == "out_stage") { if (value != 16 && value != 32) return fail(ctx, MOVI_GPU_ERR_ARG, "out_stage: 16 or 32"); ctx->out_stage = value; }
    else if (k == "probe_vmm_align_mb") { if (value < 0 || value > 4096) return fail(ctx, MOVI_GPU_ERR_ARG, "probe_vmm_align_mb: 0..4096"); ctx->probe_vmm_align_mb = value; }
    else if (k == "blocks_per_sm") { if (value < 0 || value > 32) return fail(ctx, MOVI_GPU_ERR_ARG, "blocks_per_sm: 0..32"); ctx->blocks_per_sm = value; }
    else if (k == "threads_per_block") {
        if (value < 32 || value > 256 || value % 32) return fail(ctx, MOVI_GPU_ERR_ARG, "threads_per_block: multiple of 32 in [32, 256]");
        ctx->threads_per_block = value;
    }
    else if (k == "batch_bases") { if (value < (1 << 16)) return fail(ctx, MOVI_GPU_ERR_ARG, "batch_bases: >= 65536"); ctx->batch_bases = value; }
    else if (k == "pml_sector") { if (value < -1 || value > 1) return fail(ctx, MOVI_GPU_ERR_ARG, "pml_sector: -1 (by table size), 0, 1"); ctx->pml_sector = value; }
    else if (k == "smem_carveout") { if (value < -1 || value > 100) return fail(ctx, MOVI_GPU_ERR_ARG, "smem_carveout: -1, 0..100"); ctx->smem_carveout = value; }
    else if (k == "kernel_streams") { if (value < 1 || value > NKS) return fail(ctx, MOVI_GPU_ERR_ARG, "kernel_streams: 1..4"); ctx->kernel_streams = value; }
    else if (k == "pipeline_slots") { if (value < 2 || value > NBUF) return fail(ctx, MOVI_GPU_ERR_ARG, "pipeline_slots: 2..8"); ctx->pipeline_slots = value; }
    else if (k == "size_grid_to_batch") { if (value < 0 || value > 64) return fail(ctx, MOVI_GPU_ERR_ARG, "size_grid_to_batch: 0 (off) or reads per lane 1..64"); ctx->size_grid_to_batch = value; }
    else return fail(ctx, MOVI_GPU_ERR_ARG, "unknown option " + k);
    return MOVI_GPU_OK;
}
int64_t movi_gpu_get_option(const movi_gpu_ctx* ctx, const char* key) {
    if (!ctx || !key) return -1;
    const std::string k(key);
    if (k == "blocks_per_sm") return ctx->blocks_per_sm;
    if (k == "threads_per_block") return ctx->threads_per_block;
    if (k == "batch_bases") return ctx->batch_bases;
    if (k == "kernel_streams") return ctx->kernel_streams;
    if (k == "pml_sector") return ctx->pml_sector;
    if (k == "pipeline_slots") return ctx->pipeline_slots;
    if (k == "size_grid_to_batch") return ctx->size_grid_to_batch;
    if (k == "num_sms") return ctx->num_sms;
    if (k == "force_wide") return ctx->force_wide;
    if (k == "out_stage") return ctx->out_stage;
    return -1;
}

static int upload_impl(movi_gpu_ctx* ctx, const MoviGpuIndexDesc* d, int layout, uint2* adopt_rows);

int movi_gpu_upload_index(movi_gpu_ctx* ctx, const MoviGpuIndexDesc* d, int layout) {
    return upload_impl(ctx, d, layout, nullptr);
}

int movi_gpu_adopt_device_rows(movi_gpu_ctx* ctx, const MoviGpuIndexDesc* d, void* d_rows) {
    if (!d_rows) return ctx ? fail(ctx, MOVI_GPU_ERR_ARG, "null device row table") : MOVI_GPU_ERR_ARG;
    if (d && d->mode != MOVI_MODE_REGULAR_THRESHOLDS)
        return fail(ctx, MOVI_GPU_ERR_INDEX, "an adopted device table must already be in the 8-byte regular-thresholds row format");
    return upload_impl(ctx, d, MOVI_GPU_LAYOUT_EXPANDED8, static_cast<uint2*>(d_rows));
}

static int upload_impl(movi_gpu_ctx* ctx, const MoviGpuIndexDesc* d, int layout, uint2* adopt_rows) {
    if (!ctx || !d) return MOVI_GPU_ERR_ARG;
    if (layout != MOVI_GPU_LAYOUT_AUTO && layout != MOVI_GPU_LAYOUT_EXPANDED8 && layout != MOVI_GPU_LAYOUT_COMPACT)
        return fail(ctx, MOVI_GPU_ERR_ARG, "unknown device layout");
    if (d->mode != MOVI_MODE_REGULAR_THRESHOLDS && d->mode != MOVI_MODE_BLOCKED_THRESHOLDS && d->mode != MOVI_MODE_SAMPLED_THRESHOLDS &&
        d->mode != MOVI_MODE_REGULAR && d->mode != MOVI_MODE_BLOCKED && d->mode != MOVI_MODE_SAMPLED)
        return fail(ctx, MOVI_GPU_ERR_INDEX, "unsupported index mode " + std::to_string(d->mode) +
                    " (regular-thresholds 6, blocked-thresholds 8, sampled-thresholds 7, regular 3, blocked 2, sampled 5)");
    // NO_THRESHOLDS (include/utils.hpp:146): PML always repositions with the --rpml rule.  "regular" also has 12-bit lengths
    // (its own form of the device row, flag nt); "blocked" and "sampled" fit the ordinary one with an empty threshold mask.
    const bool no_thr = d->mode == MOVI_MODE_REGULAR || d->mode == MOVI_MODE_BLOCKED || d->mode == MOVI_MODE_SAMPLED;
    const bool nt = d->mode == MOVI_MODE_REGULAR;
    const bool blocked = d->mode == MOVI_MODE_BLOCKED_THRESHOLDS || d->mode == MOVI_MODE_BLOCKED;
    const bool sampled = d->mode == MOVI_MODE_SAMPLED_THRESHOLDS || d->mode == MOVI_MODE_SAMPLED;
    if (no_thr && (adopt_rows || layout == MOVI_GPU_LAYOUT_COMPACT))
        return fail(ctx, MOVI_GPU_ERR_INDEX, "the non-threshold index types are uploaded from host rows, in the expanded layout");
    if ((!d->rows && !adopt_rows) || !d->alphamap || d->r == 0) return fail(ctx, MOVI_GPU_ERR_INDEX, "empty index description");
    if (d->r >= (1ull << R8_MAX_ROW_BITS)) return fail(ctx, MOVI_GPU_ERR_INDEX, "more than 2^35 rows");
    if (d->end_bwt_idx >= d->r) return fail(ctx, MOVI_GPU_ERR_INDEX, "end_bwt_idx out of range");
    const bool sep = (d->sigma == 5 && d->alphabet[0] == '%');       // use_separator(), src/move_structure.cpp:547-552
    if (!sep && (d->sigma < 1 || d->sigma > 4))
        return fail(ctx, MOVI_GPU_ERR_INDEX, "alphabet must have at most 4 characters, or be \"%ACGT\" with separators");
    if (blocked && (!d->id_blocks || d->block_size == 0 || d->id_blocks_size < (d->r + d->block_size - 1) / d->block_size))
        return fail(ctx, MOVI_GPU_ERR_INDEX, "blocked-thresholds index without a consistent id_blocks table");
    if (sampled && (!d->tally_ids || d->tally_checkpoints == 0 || d->tally_len < d->r / d->tally_checkpoints + 1))
        return fail(ctx, MOVI_GPU_ERR_INDEX, "sampled-thresholds index without a consistent id checkpoint table");
    if (sep && d->sep_count && (!d->separators_thresholds || (d->sep_map_size && !d->separators_map)))
        return fail(ctx, MOVI_GPU_ERR_INDEX, "separators index without its thresholds table");

    CU(cudaSetDevice(ctx->device));
    free_index(ctx);
    cudaStream_t st = ctx->st_k[0];
    const unsigned long long r = d->r;
    // whatever path leaves this function before the index is complete: nothing stays allocated (adopted rows go back to
    // the caller), and the conversion scratch below frees itself
    struct Rollback {
        movi_gpu_ctx* c; uint2* adopted; unsigned long long r; uint32_t sep; bool permuted = false, done = false;
        ~Rollback() {
            if (done) return;
            if (adopted && c->d_rows == adopted) {
                if (permuted) {      // the caller gets its table back in the format it handed over
                    relayout_regular_undo_kernel<<<c->num_sms * 8, 256, 0, c->st_k[0]>>>(adopted, r, sep);
                    cudaStreamSynchronize(c->st_k[0]);
                }
                c->d_rows = nullptr;
            }
            free_index(c);
        }
    } rollback{ctx, adopt_rows, r, sep ? 1u : 0u};
    struct Scratch { void* p = nullptr; ~Scratch() { if (p) cudaFree(p); } };
    // Device layout: EXPANDED8 unless the caller asks for COMPACT, or AUTO finds that the expanded table (plus the raw
    // copy the conversion needs) does not fit in free HBM.  COMPACT exists for blocked-thresholds (its 6-byte rows are
    // kept as they are on disk); regular-thresholds is 8 bytes on disk anyway; sampled-thresholds is always expanded.
    bool compact = false;
    const bool can_compact = !adopt_rows && ((d->mode == MOVI_MODE_BLOCKED_THRESHOLDS && (d->block_size & (d->block_size - 1)) == 0) ||
                                             d->mode == MOVI_MODE_SAMPLED_THRESHOLDS);
    if (can_compact) {
        if (layout == MOVI_GPU_LAYOUT_COMPACT) compact = true;
        else if (layout == MOVI_GPU_LAYOUT_AUTO) {
            size_t free_b = 0, total_b = 0;
            CU(cudaMemGetInfo(&free_b, &total_b));
            const double raw_b = d->mode == MOVI_MODE_BLOCKED_THRESHOLDS ? 6.0 : 4.0;
            compact = (double)r * (8 + raw_b) + (256ull << 20) > (double)free_b;
        }
    }
    if (layout == MOVI_GPU_LAYOUT_COMPACT && !compact && d->mode == MOVI_MODE_BLOCKED_THRESHOLDS)
        return fail(ctx, MOVI_GPU_ERR_INDEX, "compact layout needs a power-of-two id block size");
    uint32_t n_last_host = 0;
    const bool compact_t3 = compact && d->mode == MOVI_MODE_SAMPLED_THRESHOLDS;
    if (compact_t3) {
        // T4 (dev_index.cuh): the 3-byte rows go up as they are, one pass adds every row's walk-back distance
        const size_t tally_bytes = (size_t)d->sigma * d->tally_len * 5;
        Scratch s_raw;
        CU(cudaMalloc(&s_raw.p, r * 3 + 16));
        uint8_t* d_raw = static_cast<uint8_t*>(s_raw.p);
        CU(cudaMemsetAsync(d_raw + r * 3, 0, 16, st));
        CU(cudaMemcpyAsync(d_raw, d->rows, r * 3, cudaMemcpyHostToDevice, st));
        CU(cudaMalloc(&ctx->d_rows4, (r + 4) * 4));
        CU(cudaMemsetAsync(ctx->d_rows4 + r, 0, 16, st));
        CU(cudaMalloc(&ctx->d_tally, tally_bytes + 16));
        CU(cudaMemsetAsync(ctx->d_tally + tally_bytes, 0, 16, st));
        CU(cudaMemcpyAsync(ctx->d_tally, d->tally_ids, tally_bytes, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(ctx->d_status, 0, sizeof(unsigned int), st));
        convert_sampled_t4_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(d_raw), ctx->d_rows4, r, d->end_bwt_idx, ctx->d_tally,
                                                                  d->tally_len, d->tally_checkpoints, ctx->d_status);
        ctx->launches += 1;
        unsigned int status = 0;
        CU(cudaMemcpyAsync(&status, ctx->d_status, sizeof(status), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        if (status) {
            CU(cudaMemset(ctx->d_status, 0, sizeof(unsigned int)));
            return fail(ctx, MOVI_GPU_ERR_INDEX, "[get id - tally mode] The offset in the destination run should be always smaller than the length of that run.");
        }
        ctx->index_bytes = (r + 4) * 4 + tally_bytes + 16;
        const uint8_t* lastp = static_cast<const uint8_t*>(d->rows) + (r - 1) * 3;
        n_last_host = (uint32_t)lastp[0] | (((uint32_t)lastp[2] >> 1) & 1u) << 8;
    } else if (compact) {
        CU(cudaMalloc(&ctx->d_rows6, r * 6 + 16));
        CU(cudaMemsetAsync(reinterpret_cast<uint8_t*>(ctx->d_rows6) + r * 6, 0, 16, st));
        CU(cudaMemcpyAsync(ctx->d_rows6, d->rows, r * 6, cudaMemcpyHostToDevice, st));
        const size_t nb_bytes = (size_t)d->sigma * d->id_blocks_size * sizeof(uint32_t);
        CU(cudaMalloc(&ctx->d_id_blocks, nb_bytes));
        CU(cudaMemcpyAsync(ctx->d_id_blocks, d->id_blocks, nb_bytes, cudaMemcpyHostToDevice, st));
        ctx->index_bytes = r * 6 + 16 + nb_bytes;
        uint16_t f_n; std::memcpy(&f_n, static_cast<const uint8_t*>(d->rows) + (r - 1) * 6 + 2, 2);
        n_last_host = f_n & 0x3FFu;
    } else {
    if (adopt_rows) ctx->d_rows = adopt_rows;        // already resident (r + 4 rows, 8-byte format); ownership moves to the context
    else CU(cudaMalloc(&ctx->d_rows, (r + 4) * sizeof(uint2)));
    CU(cudaMemsetAsync(ctx->d_rows + r, 0, 4 * sizeof(uint2), st));
    ctx->index_bytes = (r + 4) * sizeof(uint2);
    }

    if (compact) {
        // nothing to convert
    } else if (adopt_rows || d->mode == MOVI_MODE_REGULAR_THRESHOLDS || nt) {
        // the 8-byte rows of the reference: copy (unless already resident), then permute the bit fields into R8 in place
        if (!adopt_rows) CU(cudaMemcpyAsync(ctx->d_rows, d->rows, r * sizeof(uint2), cudaMemcpyHostToDevice, st));
        if (nt) relayout_regular_nt_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(ctx->d_rows, r);
        else { relayout_regular_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(ctx->d_rows, r, sep ? 1u : 0u); rollback.permuted = adopt_rows != nullptr; }
        ctx->launches += 1;
        CU(cudaGetLastError());
    } else if (blocked) {
        Scratch s_raw, s_blocks, s_first;
        CU(cudaMalloc(&s_raw.p, r * 6 + 8));
        CU(cudaMalloc(&s_blocks.p, (size_t)d->sigma * d->id_blocks_size * sizeof(uint32_t)));
        CU(cudaMalloc(&s_first.p, 8 * sizeof(unsigned long long)));
        uint16_t* d_raw = static_cast<uint16_t*>(s_raw.p); uint32_t* d_blocks = static_cast<uint32_t*>(s_blocks.p);
        unsigned long long* d_first = static_cast<unsigned long long*>(s_first.p);
        CU(cudaMemcpyAsync(d_raw, d->rows, r * 6, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_blocks, d->id_blocks, (size_t)d->sigma * d->id_blocks_size * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_first, d->first_runs, 8 * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
        convert_blocked_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(d_raw, ctx->d_rows, r, d->end_bwt_idx, sep ? 1u : 0u, no_thr ? 1u : 0u, d_blocks, d->id_blocks_size,
                                                               d->block_size, d_first);
        ctx->launches += 1;
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
    } else {
        Scratch s_raw, s_tally;
        const size_t tally_bytes = (size_t)d->sigma * d->tally_len * 5;
        CU(cudaMalloc(&s_raw.p, r * 3 + 16));
        uint8_t* d_raw = static_cast<uint8_t*>(s_raw.p);
        CU(cudaMemsetAsync(d_raw + r * 3, 0, 16, st));
        CU(cudaMalloc(&s_tally.p, tally_bytes + 8));
        uint8_t* d_tally = static_cast<uint8_t*>(s_tally.p);
        CU(cudaMemcpyAsync(d_raw, d->rows, r * 3, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_tally, d->tally_ids, tally_bytes, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(ctx->d_status, 0, sizeof(unsigned int), st));
        convert_sampled_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(d_raw), ctx->d_rows, r, d->end_bwt_idx, sep ? 1u : 0u, no_thr ? 1u : 0u, d_tally, d->tally_len,
                                                               d->tally_checkpoints, ctx->d_status);
        ctx->launches += 1;
        unsigned int status = 0;
        CU(cudaMemcpyAsync(&status, ctx->d_status, sizeof(status), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        if (status) {
            CU(cudaMemset(ctx->d_status, 0, sizeof(unsigned int)));
            return fail(ctx, MOVI_GPU_ERR_INDEX, "[get id - tally mode] The offset in the destination run should be always smaller than the length of that run.");
        }
    }

    DevIndex& ix = ctx->ix;
    std::memset(&ix, 0, sizeof(ix));
    ix.rows = ctx->d_rows;
    ix.rows6 = ctx->d_rows6;
    ix.id_blocks = ctx->d_id_blocks;
    ix.layout = compact_t3 ? LAYOUT_T4 : compact ? LAYOUT_B6 : LAYOUT_R8;
    ix.nt = nt ? 1u : 0u;
    ix.rows4 = ctx->d_rows4; ix.tally = ctx->d_tally; ix.tally_len = d->tally_len; ix.tally_cp = d->tally_checkpoints;
    if (compact && !compact_t3) {
        ix.id_blocks_size = (uint32_t)d->id_blocks_size;
        ix.block_shift = 0;
        while ((1ull << ix.block_shift) < d->block_size) ix.block_shift++;
    }
    ix.r = r;
    ix.end_idx = d->end_bwt_idx;
    ix.sep = sep ? 1u : 0u;
    ix.sigma = d->sigma;
    for (int k = 0; k < 4; k++) ix.end_thr[k] = (uint32_t)std::min<uint64_t>(d->end_bwt_idx_thresholds[k], 0xFFFFFFFFull);
    for (int k = 0; k < 8; k++) {
        ix.first_runs[k] = d->first_runs[k]; ix.last_runs[k] = d->last_runs[k];
        ix.first_offsets[k] = (uint32_t)d->first_offsets[k]; ix.last_offsets[k] = (uint32_t)d->last_offsets[k];
    }
    for (int b = 0; b < 256; b++) {
        // check_alphabet, src/move_structure.cpp:383-397: '%' is never a legal query character with separators
        const uint64_t a = d->alphamap[b];
        ix.code_of[b] = (a < d->sigma && !(sep && b == '%')) ? (uint8_t)a : (uint8_t)CODE_ILLEGAL;
    }
    CU(cudaStreamSynchronize(st));   // the table copy / conversion ran on a non-blocking stream
    if (!compact) {
        // a corrupt file must not become out-of-range reads of the side tables: every row's character is below sigma
        CU(cudaMemsetAsync(ctx->d_status, 0, sizeof(unsigned int), st));
        check_row_chars_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(ctx->d_rows, r, d->end_bwt_idx, d->sigma, ctx->d_status);
        ctx->launches += 1;
        unsigned int status = 0;
        CU(cudaMemcpyAsync(&status, ctx->d_status, sizeof(status), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (status) { CU(cudaMemset(ctx->d_status, 0, sizeof(unsigned int))); return fail(ctx, MOVI_GPU_ERR_INDEX, "corrupt index: a row holds a character outside the alphabet"); }
    }
    {   // n(r-1) for the start state of query_pml
        uint2 last = make_uint2(0u, n_last_host);
        if (!compact) CU(cudaMemcpy(&last, ctx->d_rows + (r - 1), sizeof(last), cudaMemcpyDeviceToHost));
        ix.n_last = nt ? row_n<true>(last.y) : (last.y & 0x7FFu);
        if (ix.n_last == 0) return fail(ctx, MOVI_GPU_ERR_INDEX, "last row has length 0");
    }
    if (sep && !no_thr) {
        // separators_thresholds_map: row -> slot; keep it as a sorted key array + thresholds in key order,
        // plus one trailing entry that mirrors slot 0 (what unordered_map::operator[] would yield for a miss)
        std::vector<std::pair<uint64_t, uint64_t>> kv(d->sep_map_size);
        for (uint64_t i = 0; i < d->sep_map_size; i++) kv[i] = {d->separators_map[2 * i], d->separators_map[2 * i + 1]};
        std::sort(kv.begin(), kv.end());
        std::vector<unsigned long long> keys(kv.size() + 1, 0);
        std::vector<uint16_t> vals((kv.size() + 1) * 4, 0);
        for (size_t i = 0; i < kv.size(); i++) {
            keys[i] = kv[i].first;
            if (kv[i].second >= d->sep_count) return fail(ctx, MOVI_GPU_ERR_INDEX, "separators map points outside the thresholds table");
            for (int k = 0; k < 4; k++) vals[i * 4 + k] = d->separators_thresholds[kv[i].second * 4 + k];
        }
        if (d->sep_count) for (int k = 0; k < 4; k++) vals[kv.size() * 4 + k] = d->separators_thresholds[k];
        CU(cudaMalloc(&ctx->d_sep_keys, keys.size() * sizeof(unsigned long long)));
        CU(cudaMalloc(&ctx->d_sep_vals, vals.size() * sizeof(uint16_t)));
        CU(cudaMemcpy(ctx->d_sep_keys, keys.data(), keys.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(ctx->d_sep_vals, vals.data(), vals.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        ix.sep_keys = ctx->d_sep_keys; ix.sep_vals = ctx->d_sep_vals; ix.sep_cnt = (uint32_t)kv.size();
        ctx->index_bytes += keys.size() * 8 + vals.size() * 2;
    }
    CU(cudaStreamSynchronize(st));
    ctx->mode = d->mode;
    ctx->always_rpml = no_thr;
    ctx->has_index = true;
    rollback.done = true;
    return MOVI_GPU_OK;
}

int movi_gpu_share_index(movi_gpu_ctx* ctx, const movi_gpu_ctx* owner) {
    if (!ctx || !owner) return MOVI_GPU_ERR_ARG;
    if (!owner->has_index) return fail(ctx, MOVI_GPU_ERR_INDEX, "the owning context has no index");
    if (owner->device != ctx->device) return fail(ctx, MOVI_GPU_ERR_ARG, "an index can only be shared between contexts of the same GPU");
    if (owner == ctx) return MOVI_GPU_OK;
    CU(cudaSetDevice(ctx->device));
    free_index(ctx);
    ctx->ix = owner->ix;
    ctx->ix.prefix = nullptr;            // (built on demand per context: ensure_prefix)
    ctx->d_rows = owner->d_rows; ctx->d_rows6 = owner->d_rows6; ctx->d_id_blocks = owner->d_id_blocks; ctx->d_rows4 = owner->d_rows4;
    ctx->d_tally = owner->d_tally; ctx->d_sep_keys = owner->d_sep_keys; ctx->d_sep_vals = owner->d_sep_vals;
    ctx->mode = owner->mode; ctx->always_rpml = owner->always_rpml; ctx->index_bytes = 0;
    ctx->borrowed_index = true; ctx->has_index = true;
    return MOVI_GPU_OK;
}

int movi_gpu_index_export_rows(movi_gpu_ctx* ctx, uint64_t first, uint64_t count, void* host_rows) {
    if (!ctx || (count && !host_rows)) return MOVI_GPU_ERR_ARG;
    if (!ctx->has_index || !ctx->d_rows) return fail(ctx, MOVI_GPU_ERR_INDEX, "no 8-byte row table resident");
    if (first > ctx->ix.r || count > ctx->ix.r - first) return fail(ctx, MOVI_GPU_ERR_ARG, "row range out of bounds");
    CU(cudaSetDevice(ctx->device));
    const uint64_t chunk = 1ull << 24;              // 128 MB of scratch at a time
    uint2* d_tmp = nullptr;
    CU(cudaMalloc(&d_tmp, std::min(chunk, std::max<uint64_t>(count, 1)) * sizeof(uint2)));
    int rc = MOVI_GPU_OK;
    for (uint64_t done = 0; done < count && rc == MOVI_GPU_OK; done += chunk) {
        const uint64_t m = std::min(chunk, count - done);
        if (ctx->ix.nt) relayout_nt_to_file_kernel<<<ctx->num_sms * 8, 256, 0, ctx->st_k[0]>>>(ctx->d_rows + first + done, d_tmp, m);
        else relayout_to_file_kernel<<<ctx->num_sms * 8, 256, 0, ctx->st_k[0]>>>(ctx->d_rows + first + done, d_tmp, m, ctx->ix.sep);
        if (cudaMemcpyAsync(static_cast<uint2*>(host_rows) + done, d_tmp, m * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->st_k[0]) != cudaSuccess ||
            cudaStreamSynchronize(ctx->st_k[0]) != cudaSuccess)
            rc = fail(ctx, MOVI_GPU_ERR_CUDA, std::string("export rows: ") + cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(d_tmp);
    return rc;
}

// the flags of the call in progress travel in the context (one call at a time per context)
struct CallFlags { movi_gpu_ctx* c; CallFlags(movi_gpu_ctx* ctx, uint32_t f) : c(ctx) { if (c) c->call_flags = f; } ~CallFlags() { if (c) c->call_flags = 0; } };

int movi_gpu_query_pml(movi_gpu_ctx* ctx, const char* bases, const uint64_t* read_offsets, uint64_t n_reads,
                       uint32_t flags, uint16_t* out_ml) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    CallFlags cf(ctx, flags);
    if (n_reads && !out_ml) return fail(ctx, MOVI_GPU_ERR_ARG, "null output buffer");
    return host_query(ctx, Q_PML, bases, read_offsets, n_reads, out_ml, nullptr, nullptr);
}
int movi_gpu_query_classify(movi_gpu_ctx* ctx, int zml, const char* bases, const uint64_t* read_offsets, uint64_t n_reads, uint32_t flags,
                            uint32_t bin_width, uint32_t max_value_thr, uint32_t* out_bins_above, uint32_t* out_bins_below,
                            uint64_t* out_sum_max, uint16_t* out_ml) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    CallFlags cf(ctx, flags);
    if (bin_width == 0) return fail(ctx, MOVI_GPU_ERR_ARG, "bin_width must be positive");
    if (n_reads && (!out_bins_above || !out_bins_below || !out_sum_max)) return fail(ctx, MOVI_GPU_ERR_ARG, "null output buffer");
    const Kind kind = zml ? Q_ZML : Q_PML;
    int rc = check_query_ready(ctx, kind, 0);
    if (rc) return rc;
    ClassifyOut c; c.bin_width = bin_width; c.thr = max_value_thr; c.above = out_bins_above; c.below = out_bins_below; c.sum = out_sum_max;
    return host_query(ctx, kind, bases, read_offsets, n_reads, out_ml, nullptr, nullptr, &c);
}

int movi_gpu_query_ml8(movi_gpu_ctx* ctx, int zml, const char* bases, const uint64_t* read_offsets, uint64_t n_reads, uint32_t flags,
                       uint8_t* out_ml8) {
    CallFlags cf(ctx, flags);
    const Kind kind = zml ? Q_ZML : Q_PML;
    int rc = check_query_ready(ctx, kind, n_reads);
    if (rc) return rc;
    if (n_reads && (!read_offsets || !out_ml8)) return fail(ctx, MOVI_GPU_ERR_ARG, "null buffer");
    uint64_t longest = 0;
    for (uint64_t i = 0; i < n_reads; i++) longest = std::max(longest, read_offsets[i + 1] - read_offsets[i]);
    if (longest > 255) return fail(ctx, MOVI_GPU_ERR_ARG, "movi_gpu_query_ml8: a read is longer than 255 bases (its lengths may not fit one byte); use the 16-bit entry points");
    return host_query(ctx, kind, bases, read_offsets, n_reads, nullptr, nullptr, nullptr, nullptr, out_ml8);
}

uint64_t movi_gpu_bpf_image_bytes(const uint64_t* read_offsets, const uint64_t* id_offsets, uint64_t n_reads) {
    if (!read_offsets || !id_offsets) return 0;
    return bpf_record_at(read_offsets, id_offsets, n_reads);
}

int movi_gpu_query_bpf(movi_gpu_ctx* ctx, int zml, const char* bases, const uint64_t* read_offsets, uint64_t n_reads, uint32_t flags,
                       const char* ids, const uint64_t* id_offsets, uint8_t* out_image, uint64_t image_capacity) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    CallFlags cf(ctx, flags);
    const Kind kind = zml ? Q_ZML : Q_PML;
    int rc = check_query_ready(ctx, kind, 0);
    if (rc) return rc;
    if (n_reads == 0) return MOVI_GPU_OK;
    if (!read_offsets || !ids || !id_offsets || !out_image) return fail(ctx, MOVI_GPU_ERR_ARG, "null buffer");
    for (uint64_t i = 0; i < n_reads; i++)
        if (id_offsets[i + 1] - id_offsets[i] > 0xFFFFull) return fail(ctx, MOVI_GPU_ERR_ARG, "a read id is longer than 65535 bytes (the .bpf record holds its length in 16 bits)");
    if (movi_gpu_bpf_image_bytes(read_offsets, id_offsets, n_reads) > image_capacity) return fail(ctx, MOVI_GPU_ERR_ARG, "the .bpf image buffer is too small");
    BpfOut b; b.ids = ids; b.id_offsets = id_offsets; b.image = out_image;
    return host_query(ctx, kind, bases, read_offsets, n_reads, nullptr, nullptr, nullptr, nullptr, nullptr, &b);
}

int movi_gpu_bpf_assemble_device(movi_gpu_ctx* ctx, const uint16_t* d_ml, const uint64_t* d_read_offsets, uint64_t n_reads,
                                 const char* d_ids, const uint64_t* d_id_offsets, uint8_t* d_image, void* stream) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    if (n_reads == 0) return MOVI_GPU_OK;
    if (!d_ml || !d_read_offsets || !d_ids || !d_id_offsets || !d_image) return fail(ctx, MOVI_GPU_ERR_ARG, "null buffer");
    const unsigned blocks = (unsigned)std::min<uint64_t>((n_reads * 32 + 255) / 256, (uint64_t)ctx->num_sms * 16);
    bpf_assemble_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_ml, reinterpret_cast<const unsigned long long*>(d_read_offsets), 0ull, n_reads,
                                                                  reinterpret_cast<const uint8_t*>(d_ids),
                                                                  reinterpret_cast<const unsigned long long*>(d_id_offsets), d_image);
    ctx->launches += 1;
    CU(cudaGetLastError());
    return MOVI_GPU_OK;
}

int movi_gpu_query_zml(movi_gpu_ctx* ctx, const char* bases, const uint64_t* read_offsets, uint64_t n_reads,
                       uint32_t, uint16_t* out_ml) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    if (n_reads && !out_ml) return fail(ctx, MOVI_GPU_ERR_ARG, "null output buffer");
    int rc = check_query_ready(ctx, Q_ZML, 0);
    if (rc) return rc;
    return host_query(ctx, Q_ZML, bases, read_offsets, n_reads, out_ml, nullptr, nullptr);
}
int movi_gpu_query_count(movi_gpu_ctx* ctx, const char* bases, const uint64_t* read_offsets, uint64_t n_reads,
                         uint32_t, uint64_t* out_count, int32_t* out_pos_on_r) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    if (n_reads && (!out_count || !out_pos_on_r)) return fail(ctx, MOVI_GPU_ERR_ARG, "null output buffer");
    int rc = check_query_ready(ctx, Q_COUNT, 0);
    if (rc) return rc;
    return host_query(ctx, Q_COUNT, bases, read_offsets, n_reads, nullptr, out_count, out_pos_on_r);
}

static unsigned long long* next_counter(movi_gpu_ctx* ctx) {
    unsigned long long* p = ctx->d_counters + (ctx->counter_slot & 63u);
    ctx->counter_slot++;
    return p;
}

}  // extern "C" (the helpers below are C++)

namespace {
// ---- device-side ingest: a chunk of a FASTA / FASTQ file (starting at a record boundary) parsed in HBM (parse_kernels.cuh)

// raw chunk (host memory) -> ctx->p_bases / p_offsets / p_ids / p_id_offsets on the device; synchronises `st`
int parse_chunk(movi_gpu_ctx* ctx, const char* raw, uint64_t n, int fastq, int final_chunk, cudaStream_t st, ParseResult* out) {
    std::memset(out, 0, sizeof *out);
    if (n >= 0xFFFFFF00ull) return fail(ctx, MOVI_GPU_ERR_ARG, "a file chunk must be smaller than 4 GiB");
    int rc;
    if (!ctx->p_res) CU(cudaMalloc(&ctx->p_res, sizeof(ParseResult) + sizeof(Sum3) + 16));
    Sum3* d_total = reinterpret_cast<Sum3*>(ctx->p_res + 1);
    uint32_t* d_nl_total = reinterpret_cast<uint32_t*>(d_total + 1);
    CU(cudaMemsetAsync(ctx->p_res, 0, sizeof(ParseResult) + sizeof(Sum3) + 16, st));
    if (n == 0) { CU(cudaStreamSynchronize(st)); return MOVI_GPU_OK; }
    if ((rc = grow(ctx, ctx->p_raw, ctx->p_raw_cap, n + 64))) return rc;
    if ((rc = staged_h2d(ctx, ctx->p_raw, raw, n, st))) return rc;
    const uint64_t n_tiles = (n + PARSE_TILE - 1) / PARSE_TILE;
    if ((rc = grow(ctx, ctx->p_tile, ctx->p_tile_cap, n_tiles + 1))) return rc;
    parse_count_newlines<<<(unsigned)n_tiles, 256, 0, st>>>(ctx->p_raw, n, ctx->p_tile);
    parse_scan_u32<<<1, 1024, 0, st>>>(ctx->p_tile, n_tiles, d_nl_total);
    uint32_t n_nl = 0;
    CU(cudaMemcpyAsync(&n_nl, d_nl_total, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    ctx->launches += 2;
    // lines: every '\n'-terminated one, plus -- in the last chunk of a file -- a final line without '\n'
    const uint32_t n_lines = n_nl + ((final_chunk && raw[n - 1] != '\n') ? 1u : 0u);
    if (n_lines == 0) return MOVI_GPU_OK;               // no complete line yet: nothing consumed
    if ((rc = grow(ctx, ctx->p_nlpos, ctx->p_nlpos_cap, (size_t)n_nl + 2))) return rc;
    if ((rc = grow(ctx, ctx->p_lines, ctx->p_lines_cap, (size_t)n_lines + 1))) return rc;
    const uint32_t l_tiles = (n_lines + LINE_TILE - 1) / LINE_TILE;
    if ((rc = grow(ctx, ctx->p_ltile, ctx->p_ltile_cap, (size_t)l_tiles + 1))) return rc;
    parse_fill_newlines<<<(unsigned)n_tiles, 256, 0, st>>>(ctx->p_raw, n, ctx->p_tile, ctx->p_nlpos);
    parse_classify_lines<<<(n_lines + 255) / 256, 256, 0, st>>>(ctx->p_raw, n, ctx->p_nlpos, n_nl, n_lines, fastq, ctx->p_lines, ctx->p_res);
    parse_line_tile_sums<<<l_tiles, LINE_TILE, 0, st>>>(ctx->p_lines, n_lines, ctx->p_ltile);
    parse_scan_tiles<<<1, 32, 0, st>>>(ctx->p_ltile, l_tiles, d_total);
    Sum3 total{};
    ParseResult pr{};
    CU(cudaMemcpyAsync(&total, d_total, sizeof total, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&pr, ctx->p_res, sizeof pr, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    ctx->launches += 4;
    out->flags = pr.flags; out->n_lines = n_lines;
    if (pr.flags) return MOVI_GPU_OK;                   // the caller hands the file to the host parser
    uint32_t n_keep;
    if (fastq) n_keep = n_lines / 4;                    // records whose four lines are all there (a truncated last record is dropped)
    else n_keep = final_chunk ? total.recs : (total.recs ? total.recs - 1 : 0);   // the last record may continue in the next chunk
    if ((rc = grow(ctx, ctx->p_bases, ctx->p_bases_cap, (size_t)total.bases + 96))) return rc;
    if ((rc = grow(ctx, ctx->p_offsets, ctx->p_offsets_cap, (size_t)total.recs + 2))) return rc;
    if ((rc = grow(ctx, ctx->p_ids, ctx->p_ids_cap, (size_t)total.ids + 16))) return rc;
    if ((rc = grow(ctx, ctx->p_id_offsets, ctx->p_id_offsets_cap, (size_t)total.recs + 2))) return rc;
    parse_apply<<<l_tiles, LINE_TILE, 0, st>>>(ctx->p_raw, ctx->p_lines, n_lines, ctx->p_ltile, n_keep, ctx->p_bases, ctx->p_offsets, ctx->p_ids,
                                               ctx->p_id_offsets, ctx->p_res, 0);
    // every record kept: totals from the grand totals; what is consumed ends after the last complete line
    parse_finish<<<1, 32, 0, st>>>(d_total, n_keep, final_chunk ? n : 0ull, ctx->p_offsets, ctx->p_id_offsets, ctx->p_res);
    CU(cudaMemcpyAsync(&pr, ctx->p_res, sizeof pr, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    ctx->launches += 2;
    if (!final_chunk && n_keep == total.recs) {
        // (FASTQ whose last complete record ends exactly at the last newline of the chunk) consumed = one past that newline
        uint32_t last_nl = 0;
        CU(cudaMemcpy(&last_nl, ctx->p_nlpos + (n_nl - 1), 4, cudaMemcpyDeviceToHost));
        pr.consumed = fastq ? (uint64_t)last_nl + 1 : 0;
        if (fastq && (n_lines & 3u)) return fail(ctx, MOVI_GPU_ERR_QUERY, "internal: FASTQ chunk accounting");
    }
    *out = pr;
    out->n_lines = n_lines;
    return MOVI_GPU_OK;
}
}  // namespace

extern "C" {

extern "C" int movi_gpu_parse_chunk(movi_gpu_ctx* ctx, const char* raw, uint64_t n_bytes, int fastq, int final_chunk,
                                    char* out_bases, uint64_t bases_capacity, uint64_t* out_offsets, char* out_ids, uint64_t ids_capacity,
                                    uint64_t* out_id_offsets, uint64_t reads_capacity, MoviGpuChunkResult* res) {
    if (!ctx || !res || (n_bytes && !raw)) return MOVI_GPU_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    ParseResult pr;
    int rc = parse_chunk(ctx, raw, n_bytes, fastq, final_chunk, ctx->st_k[0], &pr);
    if (rc) return rc;
    res->n_reads = pr.n_reads; res->n_bases = pr.n_bases; res->id_bytes = pr.id_bytes; res->consumed_bytes = pr.consumed;
    res->image_bytes = 0; res->parse_flags = pr.flags;
    if (pr.flags || pr.n_reads == 0) return MOVI_GPU_OK;
    if (pr.n_bases > bases_capacity || pr.id_bytes > ids_capacity || pr.n_reads > reads_capacity) return fail(ctx, MOVI_GPU_ERR_ARG, "output buffers too small for the parsed chunk");
    CU(cudaMemcpy(out_bases, ctx->p_bases, pr.n_bases, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out_offsets, ctx->p_offsets, (pr.n_reads + 1) * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out_ids, ctx->p_ids, pr.id_bytes, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out_id_offsets, ctx->p_id_offsets, (pr.n_reads + 1) * 8, cudaMemcpyDeviceToHost));
    return MOVI_GPU_OK;
}

extern "C" int movi_gpu_file_chunk_parse(movi_gpu_ctx* ctx, const char* raw, uint64_t n_bytes, int fastq, int final_chunk, MoviGpuChunkResult* res) {
    if (!ctx || !res || (n_bytes && !raw)) return MOVI_GPU_ERR_ARG;
    if (!ctx->has_index) return fail(ctx, MOVI_GPU_ERR_INDEX, "no index uploaded");
    CU(cudaSetDevice(ctx->device));
    ParseResult pr;
    ctx->p_last = ParseResult{};
    int rc = parse_chunk(ctx, raw, n_bytes, fastq, final_chunk, ctx->st_k[0], &pr);
    if (rc) return rc;
    res->n_reads = pr.n_reads; res->n_bases = pr.n_bases; res->id_bytes = pr.id_bytes; res->consumed_bytes = pr.consumed;
    res->parse_flags = pr.flags; res->reserved = 0;
    res->image_bytes = pr.flags ? 0 : 10ull * pr.n_reads + pr.id_bytes + 2ull * pr.n_bases;
    if (!pr.flags) ctx->p_last = pr;
    return MOVI_GPU_OK;
}

extern "C" int movi_gpu_file_chunk_query(movi_gpu_ctx* ctx, int zml, uint32_t flags, uint8_t* out_image, uint64_t image_capacity, int out_fd,
                                         uint64_t file_offset) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    CallFlags cf(ctx, flags);
    const Kind kind = zml ? Q_ZML : Q_PML;
    int rc = check_query_ready(ctx, kind, 0);
    if (rc) return rc;
    const ParseResult pr = ctx->p_last;
    if (pr.n_reads == 0) return MOVI_GPU_OK;
    const uint64_t image_bytes = 10ull * pr.n_reads + pr.id_bytes + 2ull * pr.n_bases;
    if (out_image && image_bytes > image_capacity) return fail(ctx, MOVI_GPU_ERR_NOMEM, "the .bpf image buffer is too small for this chunk");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->st_k[0];
    if ((rc = grow(ctx, ctx->p_ml, ctx->p_ml_cap, (size_t)pr.n_bases + 96))) return rc;
    CU(cudaMemsetAsync(ctx->d_status, 0, sizeof(unsigned int), st));
    CU(cudaMemsetAsync(ctx->p_bases + pr.n_bases, 0, 64, st));
    rc = device_query(ctx, kind, reinterpret_cast<const char*>(ctx->p_bases), reinterpret_cast<const uint64_t*>(ctx->p_offsets), 0, pr.n_reads,
                      ctx->p_ml, nullptr, nullptr, next_counter(ctx), st);
    if (rc) return rc;
    if (out_image || out_fd >= 0) {
        if ((rc = grow(ctx, ctx->p_image, ctx->p_image_cap, (size_t)image_bytes + 16))) return rc;
        const unsigned blocks = (unsigned)std::min<uint64_t>((pr.n_reads * 32 + 255) / 256, (uint64_t)ctx->num_sms * 16);
        bpf_assemble_kernel<<<blocks, 256, 0, st>>>(ctx->p_ml, ctx->p_offsets, 0ull, pr.n_reads, ctx->p_ids, ctx->p_id_offsets, ctx->p_image);
        ctx->launches += 1;
        CU(cudaGetLastError());
        if ((rc = staged_d2h(ctx, out_image, out_fd, file_offset, ctx->p_image, image_bytes, st))) return rc;
    }
    CU(cudaStreamSynchronize(st));
    return movi_gpu_check_device_status(ctx);
}

// The chunk parsed last: walked, classified on the device (Classifier::classify's binning, src/classifier.cpp:99-143) and, when a sink is
// given, delivered as .bpf records like movi_gpu_file_chunk_query does.  Only 16 bytes per read come back for the report.
extern "C" int movi_gpu_file_chunk_classify(movi_gpu_ctx* ctx, int zml, uint32_t flags, uint32_t bin_width, uint32_t max_value_thr,
                                            uint32_t* out_above, uint32_t* out_below, uint64_t* out_sum,
                                            uint8_t* out_image, uint64_t image_capacity, int out_fd, uint64_t file_offset) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    if (bin_width == 0) return fail(ctx, MOVI_GPU_ERR_ARG, "bin_width must be positive");
    CallFlags cf(ctx, flags);
    const Kind kind = zml ? Q_ZML : Q_PML;
    int rc = check_query_ready(ctx, kind, 0);
    if (rc) return rc;
    const ParseResult pr = ctx->p_last;
    if (pr.n_reads == 0) return MOVI_GPU_OK;
    if (!out_above || !out_below || !out_sum) return fail(ctx, MOVI_GPU_ERR_ARG, "null result buffer");
    const uint64_t image_bytes = 10ull * pr.n_reads + pr.id_bytes + 2ull * pr.n_bases;
    if (out_image && image_bytes > image_capacity) return fail(ctx, MOVI_GPU_ERR_NOMEM, "the .bpf image buffer is too small for this chunk");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->st_k[0];
    if ((rc = grow(ctx, ctx->p_ml, ctx->p_ml_cap, (size_t)pr.n_bases + 96))) return rc;
    if (pr.n_reads > ctx->p_cls_cap) {
        cudaFree(ctx->p_above); cudaFree(ctx->p_below); cudaFree(ctx->p_sum); cudaFree(ctx->p_pos);
        ctx->p_above = nullptr; ctx->p_below = nullptr; ctx->p_sum = nullptr; ctx->p_cls_cap = 0;
        const size_t want = (size_t)pr.n_reads + (size_t)pr.n_reads / 8 + 64;
        CU(cudaMalloc(&ctx->p_above, want * sizeof(uint32_t))); CU(cudaMalloc(&ctx->p_below, want * sizeof(uint32_t)));
        CU(cudaMalloc(&ctx->p_sum, want * sizeof(unsigned long long)));
        ctx->p_cls_cap = want;
    }
    CU(cudaMemsetAsync(ctx->d_status, 0, sizeof(unsigned int), st));
    CU(cudaMemsetAsync(ctx->p_bases + pr.n_bases, 0, 64, st));
    rc = device_query(ctx, kind, reinterpret_cast<const char*>(ctx->p_bases), reinterpret_cast<const uint64_t*>(ctx->p_offsets), 0, pr.n_reads,
                      ctx->p_ml, nullptr, nullptr, next_counter(ctx), st);
    if (rc) return rc;
    {
        const unsigned blocks = (unsigned)std::min<uint64_t>((pr.n_reads * 32 + 255) / 256, (uint64_t)ctx->num_sms * 16);
        classify_kernel<<<blocks, 256, 0, st>>>(ctx->p_ml, ctx->p_offsets, 0ull, pr.n_reads, bin_width, max_value_thr, ctx->p_above, ctx->p_below, ctx->p_sum);
        ctx->launches += 1;
        CU(cudaGetLastError());
    }
    if (out_image || out_fd >= 0) {
        if ((rc = grow(ctx, ctx->p_image, ctx->p_image_cap, (size_t)image_bytes + 16))) return rc;
        const unsigned blocks = (unsigned)std::min<uint64_t>((pr.n_reads * 32 + 255) / 256, (uint64_t)ctx->num_sms * 16);
        bpf_assemble_kernel<<<blocks, 256, 0, st>>>(ctx->p_ml, ctx->p_offsets, 0ull, pr.n_reads, ctx->p_ids, ctx->p_id_offsets, ctx->p_image);
        ctx->launches += 1;
        CU(cudaGetLastError());
        if ((rc = staged_d2h(ctx, out_image, out_fd, file_offset, ctx->p_image, image_bytes, st))) return rc;
    }
    CU(cudaMemcpyAsync(out_above, ctx->p_above, pr.n_reads * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out_below, ctx->p_below, pr.n_reads * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out_sum, ctx->p_sum, pr.n_reads * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return movi_gpu_check_device_status(ctx);
}

// The reads parsed last, copied back to the host (any pointer may be NULL): ids for a report, sequences for --filter.
extern "C" int movi_gpu_file_chunk_fetch(movi_gpu_ctx* ctx, char* out_bases, uint64_t* out_offsets, char* out_ids, uint64_t* out_id_offsets) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    const ParseResult pr = ctx->p_last;
    if (pr.n_reads == 0) return MOVI_GPU_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->st_k[0];
    int rc;
    if (out_bases && (rc = staged_d2h(ctx, reinterpret_cast<uint8_t*>(out_bases), -1, 0, ctx->p_bases, pr.n_bases, st))) return rc;
    if (out_ids && (rc = staged_d2h(ctx, reinterpret_cast<uint8_t*>(out_ids), -1, 0, ctx->p_ids, pr.id_bytes, st))) return rc;
    if (out_offsets) CU(cudaMemcpyAsync(out_offsets, ctx->p_offsets, (pr.n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (out_id_offsets) CU(cudaMemcpyAsync(out_id_offsets, ctx->p_id_offsets, (pr.n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MOVI_GPU_OK;
}

// The chunk parsed last, for --filter: walked, binned, and the echo of the reads Classifier::classify keeps assembled on the device
// (filter_kernels.cuh).  The text stays in HBM until movi_gpu_file_chunk_filter_deliver fetches it.
extern "C" int movi_gpu_file_chunk_filter(movi_gpu_ctx* ctx, int zml, uint32_t flags, uint32_t bin_width, uint32_t max_value_thr, int invert,
                                          uint64_t* text_bytes) {
    if (!ctx || !text_bytes) return MOVI_GPU_ERR_ARG;
    *text_bytes = 0; ctx->p_text_bytes = 0;
    if (bin_width == 0) return fail(ctx, MOVI_GPU_ERR_ARG, "bin_width must be positive");
    CallFlags cf(ctx, flags);
    const Kind kind = zml ? Q_ZML : Q_PML;
    int rc = check_query_ready(ctx, kind, 0);
    if (rc) return rc;
    const ParseResult pr = ctx->p_last;
    if (pr.n_reads == 0) return MOVI_GPU_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->st_k[0];
    if ((rc = grow(ctx, ctx->p_ml, ctx->p_ml_cap, (size_t)pr.n_bases + 96))) return rc;
    if ((rc = grow(ctx, ctx->p_pos, ctx->p_pos_cap, (size_t)pr.n_reads + 2))) return rc;
    if (pr.n_reads > ctx->p_cls_cap) {
        cudaFree(ctx->p_above); cudaFree(ctx->p_below); cudaFree(ctx->p_sum);
        ctx->p_above = nullptr; ctx->p_below = nullptr; ctx->p_sum = nullptr; ctx->p_cls_cap = 0;
        const size_t want = (size_t)pr.n_reads + (size_t)pr.n_reads / 8 + 64;
        CU(cudaMalloc(&ctx->p_above, want * sizeof(uint32_t))); CU(cudaMalloc(&ctx->p_below, want * sizeof(uint32_t)));
        CU(cudaMalloc(&ctx->p_sum, want * sizeof(unsigned long long)));
        ctx->p_cls_cap = want;
    }
    CU(cudaMemsetAsync(ctx->d_status, 0, sizeof(unsigned int), st));
    CU(cudaMemsetAsync(ctx->p_bases + pr.n_bases, 0, 64, st));
    rc = device_query(ctx, kind, reinterpret_cast<const char*>(ctx->p_bases), reinterpret_cast<const uint64_t*>(ctx->p_offsets), 0, pr.n_reads,
                      ctx->p_ml, nullptr, nullptr, next_counter(ctx), st);
    if (rc) return rc;
    const unsigned wblocks = (unsigned)std::min<uint64_t>((pr.n_reads * 32 + 255) / 256, (uint64_t)ctx->num_sms * 16);
    classify_kernel<<<wblocks, 256, 0, st>>>(ctx->p_ml, ctx->p_offsets, 0ull, pr.n_reads, bin_width, max_value_thr, ctx->p_above, ctx->p_below, ctx->p_sum);
    filter_sizes_kernel<<<(unsigned)std::min<uint64_t>((pr.n_reads + 255) / 256, (uint64_t)ctx->num_sms * 8), 256, 0, st>>>(
        ctx->p_above, ctx->p_below, invert ? 1 : 0, ctx->p_offsets, ctx->p_id_offsets, pr.n_reads, ctx->p_pos);
    scan_u64_kernel<<<1, 1024, 0, st>>>(ctx->p_pos, pr.n_reads);
    unsigned long long total = 0;
    CU(cudaMemcpyAsync(&total, ctx->p_pos + pr.n_reads, sizeof total, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    ctx->launches += 3;
    if ((rc = movi_gpu_check_device_status(ctx))) return rc;
    if (total) {
        if ((rc = grow(ctx, ctx->p_image, ctx->p_image_cap, (size_t)total + 16))) return rc;
        filter_emit_kernel<<<wblocks, 256, 0, st>>>(ctx->p_pos, pr.n_reads, ctx->p_bases, ctx->p_offsets, ctx->p_ids, ctx->p_id_offsets, ctx->p_image);
        ctx->launches += 1;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(st));
    }
    ctx->p_text_bytes = total; *text_bytes = total;
    return MOVI_GPU_OK;
}

// ... fetched: into out_text (may be NULL) and / or written to out_fd with write() calls in order (stdout may be a pipe)
extern "C" int movi_gpu_file_chunk_filter_deliver(movi_gpu_ctx* ctx, uint8_t* out_text, uint64_t capacity, int out_fd) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    const uint64_t n = ctx->p_text_bytes;
    if (n == 0) return MOVI_GPU_OK;
    if (out_text && n > capacity) return fail(ctx, MOVI_GPU_ERR_NOMEM, "the text buffer is too small for this chunk's kept reads");
    if (!out_text && out_fd < 0) return MOVI_GPU_OK;
    CU(cudaSetDevice(ctx->device));
    return staged_d2h(ctx, out_text, out_fd, SEQUENTIAL_WRITES, ctx->p_image, n, ctx->st_k[0]);
}

extern "C" int movi_gpu_query_file_chunk(movi_gpu_ctx* ctx, int zml, const char* raw, uint64_t n_bytes, int fastq, int final_chunk, uint32_t flags,
                                         uint8_t* out_image, uint64_t image_capacity, MoviGpuChunkResult* res) {
    int rc = movi_gpu_file_chunk_parse(ctx, raw, n_bytes, fastq, final_chunk, res);
    if (rc || res->parse_flags || res->n_reads == 0) return rc;
    return movi_gpu_file_chunk_query(ctx, zml, flags, out_image, image_capacity, -1, 0);
}

int movi_gpu_query_pml_device(movi_gpu_ctx* ctx, const char* d_bases, const uint64_t* d_read_offsets, uint64_t n_reads,
                              uint32_t flags, uint16_t* d_out_ml, void* stream) {
    CallFlags cf(ctx, flags);
    int rc = check_query_ready(ctx, Q_PML, n_reads);
    if (rc) return rc;
    return device_query(ctx, Q_PML, d_bases, d_read_offsets, 0, n_reads, d_out_ml, nullptr, nullptr, next_counter(ctx), (cudaStream_t)stream);
}
int movi_gpu_query_zml_device(movi_gpu_ctx* ctx, const char* d_bases, const uint64_t* d_read_offsets, uint64_t n_reads,
                              uint32_t, uint16_t* d_out_ml, void* stream) {
    int rc = check_query_ready(ctx, Q_ZML, n_reads);
    if (rc) return rc;
    return device_query(ctx, Q_ZML, d_bases, d_read_offsets, 0, n_reads, d_out_ml, nullptr, nullptr, next_counter(ctx), (cudaStream_t)stream);
}
int movi_gpu_query_count_device(movi_gpu_ctx* ctx, const char* d_bases, const uint64_t* d_read_offsets, uint64_t n_reads,
                                uint32_t, uint64_t* d_out_count, int32_t* d_out_pos_on_r, void* stream) {
    int rc = check_query_ready(ctx, Q_COUNT, n_reads);
    if (rc) return rc;
    return device_query(ctx, Q_COUNT, d_bases, d_read_offsets, 0, n_reads, nullptr, d_out_count, d_out_pos_on_r, next_counter(ctx), (cudaStream_t)stream);
}

int movi_gpu_check_device_status(movi_gpu_ctx* ctx) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    unsigned int status = 0;
    CU(cudaMemcpy(&status, ctx->d_status, sizeof(status), cudaMemcpyDeviceToHost));
    if (status) {
        CU(cudaMemset(ctx->d_status, 0, sizeof(unsigned int)));
        if (status & 4u)
            return fail(ctx, MOVI_GPU_ERR_ARG, "a *_device call was given 2^32 or more bases: split the batch (the host-buffer entry points do)");
        return fail(ctx, MOVI_GPU_ERR_QUERY, (status & 2u)
                    ? "[get id - tally mode] The offset in the destination run should be always smaller than the length of that run. "
                      "(the reference throws here, src/move_structure.cpp:196-199)"
                    : "[reposition thresholds] no row with the read's character in the threshold direction: "
                      "This should not happen! (the reference throws here, src/move_structure_query.cpp:348,582-598)");
    }
    return MOVI_GPU_OK;
}

int movi_gpu_classify_device(movi_gpu_ctx* ctx, const uint16_t* d_ml, const uint64_t* d_read_offsets, uint64_t n_reads,
                             uint32_t bin_width, uint32_t max_value_thr, uint32_t* d_above, uint32_t* d_below,
                             uint64_t* d_sum_max, void* stream) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    if (bin_width == 0) return fail(ctx, MOVI_GPU_ERR_ARG, "bin_width must be positive");
    if (n_reads == 0) return MOVI_GPU_OK;
    const int threads = 256;
    const uint64_t warps = n_reads;
    const unsigned blocks = (unsigned)std::min<uint64_t>((warps * 32 + threads - 1) / threads, (uint64_t)ctx->num_sms * 16);
    classify_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(d_ml, reinterpret_cast<const unsigned long long*>(d_read_offsets), 0ull, n_reads,
                                                                  bin_width, max_value_thr, d_above, d_below,
                                                                  reinterpret_cast<unsigned long long*>(d_sum_max));
    ctx->launches += 1;
    CU(cudaGetLastError());
    return MOVI_GPU_OK;
}

int movi_gpu_host_alloc(void** p, uint64_t bytes) {
    if (!p) return MOVI_GPU_ERR_ARG;
    cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); *p = nullptr; return MOVI_GPU_ERR_NOMEM; }
    return MOVI_GPU_OK;
}
int movi_gpu_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? MOVI_GPU_OK : MOVI_GPU_ERR_CUDA; }

int movi_gpu_gather_probe(movi_gpu_ctx* ctx, uint64_t table_bytes, uint32_t load_bytes, uint32_t dependent, uint32_t steps,
                          uint32_t blocks_per_sm, uint32_t threads_per_block, double* seconds, uint64_t* loads_done) {
    if (!ctx || !seconds) return MOVI_GPU_ERR_ARG;
    if (load_bytes != 8 && load_bytes != 16 && load_bytes != 32) return fail(ctx, MOVI_GPU_ERR_ARG, "load_bytes: 8, 16 or 32");
    if (threads_per_block == 0 || threads_per_block > 1024 || blocks_per_sm == 0) return fail(ctx, MOVI_GPU_ERR_ARG, "bad launch shape");
    CU(cudaSetDevice(ctx->device));
    const uint2* table; uint64_t n_rows;
    if (table_bytes == 0) {
        if (!ctx->has_index || !ctx->d_rows) return fail(ctx, MOVI_GPU_ERR_ARG, "no 8-byte index table uploaded and table_bytes == 0");
        table = ctx->d_rows; n_rows = ctx->ix.r & ~3ull;
    } else {
        table_bytes &= ~255ull;
        if (ctx->probe_bytes != table_bytes || (ctx->probe_vmm_align_mb > 0) != (ctx->probe_vmm[0] != 0)) {
            if (ctx->probe_vmm[0]) { VmmBlock b; b.ptr = ctx->probe_vmm[0]; b.size = ctx->probe_vmm[1]; b.handle = ctx->probe_vmm[2]; vmm_free(&b); ctx->probe_vmm[0] = 0; }
            else cudaFree(ctx->d_probe);
            ctx->d_probe = nullptr; ctx->probe_bytes = 0;
            if (ctx->probe_vmm_align_mb > 0) {
                VmmBlock b;
                if (!vmm_alloc(ctx->device, table_bytes, (size_t)ctx->probe_vmm_align_mb << 20, &b)) return fail(ctx, MOVI_GPU_ERR_NOMEM, "VMM allocation of the probe table failed");
                ctx->probe_vmm[0] = b.ptr; ctx->probe_vmm[1] = b.size; ctx->probe_vmm[2] = b.handle;
                ctx->d_probe = reinterpret_cast<uint2*>(b.ptr);
            } else
            CU(cudaMalloc(&ctx->d_probe, table_bytes));
            ctx->probe_bytes = table_bytes;
            probe_fill_kernel<<<ctx->num_sms * 8, 256>>>(ctx->d_probe, table_bytes / 8);
            CU(cudaDeviceSynchronize());
        }
        table = ctx->d_probe; n_rows = table_bytes / 8;
    }
    unsigned long long* d_sink = nullptr;
    CU(cudaMalloc(&d_sink, sizeof(unsigned long long)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const unsigned blocks = blocks_per_sm * ctx->num_sms;
    auto run = [&](uint32_t nsteps) {
        if (load_bytes == 8) { if (dependent) gather_probe_kernel<8, true><<<blocks, threads_per_block>>>(table, n_rows, nsteps, d_sink); else gather_probe_kernel<8, false><<<blocks, threads_per_block>>>(table, n_rows, nsteps, d_sink); }
        else if (load_bytes == 16) { if (dependent) gather_probe_kernel<16, true><<<blocks, threads_per_block>>>(table, n_rows, nsteps, d_sink); else gather_probe_kernel<16, false><<<blocks, threads_per_block>>>(table, n_rows, nsteps, d_sink); }
        else { if (dependent) gather_probe_kernel<32, true><<<blocks, threads_per_block>>>(table, n_rows, nsteps, d_sink); else gather_probe_kernel<32, false><<<blocks, threads_per_block>>>(table, n_rows, nsteps, d_sink); }
    };
    run(std::max(1u, steps / 8));     // warm-up
    CU(cudaEventRecord(e0));
    run(steps);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    CU(cudaGetLastError());
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    ctx->launches += 2;
    *seconds = ms * 1e-3;
    if (loads_done) *loads_done = (uint64_t)blocks * threads_per_block * steps;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_sink);
    return MOVI_GPU_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- Movi Color: --multi-classify, --sa-entries (ext_kernel.cuh)
namespace {

int ext_ready(movi_gpu_ctx* ctx, uint32_t flags) {
    if (!ctx) return MOVI_GPU_ERR_ARG;
    if (!ctx->has_index) return fail(ctx, MOVI_GPU_ERR_INDEX, "no index uploaded");
    if (ctx->ix.layout != LAYOUT_R8 || ctx->ix.nt || ctx->always_rpml)
        return fail(ctx, MOVI_GPU_ERR_INDEX, "multi-class classification and SA entries need a threshold index in the expanded device layout");
    if (flags & MOVI_GPU_FLAG_RPML) return fail(ctx, MOVI_GPU_ERR_ARG, "--rpml is not available with multi-class classification / SA entries");
    return MOVI_GPU_OK;
}

// one batch through movi_ext_kernel: slices of whole reads (a device call holds < 2^32 bases), copy in, walk, copy out
int ext_query(movi_gpu_ctx* ctx, int ext, const char* bases, const uint64_t* offsets, uint64_t n_reads, uint32_t min_len,
              MoviGpuColorResult* out_color, uint16_t* out_ml, uint64_t* out_sa) {
    if (n_reads == 0) return MOVI_GPU_OK;
    if (!bases || !offsets) return fail(ctx, MOVI_GPU_ERR_ARG, "null read batch");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->st_k[0];
    const bool wide = ctx->ix.r > 0xFFFFFFF0ull || ctx->force_wide;
    const bool sep = ctx->ix.sep != 0;
    const uint64_t slice_bases = 1ull << 28, slice_reads = 1ull << 22;
    const bool color = (ext & EXT_COLOR) != 0;
    const size_t cnt_item = (ext & EXT_PVALUE) ? sizeof(double) : sizeof(uint32_t);
    uint64_t i0 = 0;
    while (i0 < n_reads) {
        uint64_t i1 = i0 + 1;
        while (i1 < n_reads && i1 - i0 < slice_reads && offsets[i1 + 1] - offsets[i0] <= slice_bases) i1++;
        const uint64_t nr = i1 - i0, nb = offsets[i1] - offsets[i0];
        if (nb >= 0xFFFFFFF0ull) return fail(ctx, MOVI_GPU_ERR_ARG, "a read of 2^32 bases or more");
        int rc;
        if ((rc = grow(ctx, ctx->e_bases, ctx->e_bases_cap, nb + 64))) return rc;
        if ((rc = grow(ctx, ctx->e_offsets, ctx->e_offsets_cap, nr + 1))) return rc;
        if (out_ml && (rc = grow(ctx, ctx->e_ml, ctx->e_ml_cap, nb + 32))) return rc;
        if (out_sa && (rc = grow(ctx, ctx->e_sa, ctx->e_sa_cap, nb + 8))) return rc;
        if (color && (rc = grow(ctx, ctx->e_res, ctx->e_res_cap, nr))) return rc;
        // grid: one read per lane, at most what fits the machine; the per-lane counters bound it too (4 GiB of them at most)
        uint64_t blocks = (nr + QUERY_THREADS - 1) / QUERY_THREADS;
        blocks = std::min<uint64_t>(blocks, (uint64_t)ctx->num_sms * 6);
        if (color) {
            const uint64_t per_block = (uint64_t)QUERY_THREADS * ctx->c_species * cnt_item;
            blocks = std::max<uint64_t>(1, std::min<uint64_t>(blocks, (4ull << 30) / std::max<uint64_t>(per_block, 1)));
            if ((rc = grow(ctx, ctx->e_cnt, ctx->e_cnt_cap, blocks * per_block))) return rc;
        }
        // (pageable caller buffers pass through the ring of pinned pieces: staged_h2d / staged_d2h)
        if ((rc = staged_h2d(ctx, ctx->e_bases, bases + offsets[i0], nb, st))) return rc;
        CU(cudaMemcpyAsync(ctx->e_offsets, offsets + i0, (nr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        unsigned long long* counter = next_counter(ctx);
        CU(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
        ExtArgs a{};
        a.bases = ctx->e_bases; a.offsets = ctx->e_offsets; a.origin = offsets[i0]; a.n_reads = (uint32_t)nr;
        a.out_ml = out_ml ? ctx->e_ml : nullptr;
        a.counter = counter; a.status = ctx->d_status;
        a.color_inds = ctx->c_inds; a.flat = ctx->c_flat; a.flat_size = ctx->c_flat_size; a.num_species = ctx->c_species; a.min_len = min_len;
        a.lk = ctx->c_lk; a.cnt = ctx->e_cnt; a.cnt_stride = blocks * QUERY_THREADS; a.out_color = ctx->e_res;
        a.all_p = ctx->s_all_p; a.sampled = ctx->s_sampled; a.rate = ctx->s_rate; a.out_sa = ctx->e_sa;
#define EXT_LAUNCH(IDX, SEPV, EXTV) movi_ext_kernel<IDX, SEPV, EXTV><<<(unsigned)blocks, QUERY_THREADS, 0, st>>>(ctx->ix, a)
#define EXT_PICK(EXTV) do { if (wide) { if (sep) EXT_LAUNCH(unsigned long long, true, EXTV); else EXT_LAUNCH(unsigned long long, false, EXTV); } \
                            else      { if (sep) EXT_LAUNCH(uint32_t, true, EXTV); else EXT_LAUNCH(uint32_t, false, EXTV); } } while (0)
        if (ext == EXT_COLOR) EXT_PICK(EXT_COLOR);
        else if (ext == (EXT_COLOR | EXT_PVALUE)) EXT_PICK(EXT_COLOR | EXT_PVALUE);
        else EXT_PICK(EXT_SA);
#undef EXT_PICK
#undef EXT_LAUNCH
        ctx->launches += 1;
        CU(cudaGetLastError());
        if (out_ml && (rc = staged_d2h(ctx, reinterpret_cast<uint8_t*>(out_ml + offsets[i0]), -1, 0, reinterpret_cast<const uint8_t*>(ctx->e_ml), nb * sizeof(uint16_t), st))) return rc;
        if (out_sa && (rc = staged_d2h(ctx, reinterpret_cast<uint8_t*>(out_sa + offsets[i0]), -1, 0, reinterpret_cast<const uint8_t*>(ctx->e_sa), nb * sizeof(uint64_t), st))) return rc;
        if (color && (rc = staged_d2h(ctx, reinterpret_cast<uint8_t*>(out_color + i0), -1, 0, reinterpret_cast<const uint8_t*>(ctx->e_res), nr * sizeof(MoviGpuColorResult), st))) return rc;
        CU(cudaStreamSynchronize(st));
        int qs = movi_gpu_check_device_status(ctx);
        if (qs) return qs;
        i0 = i1;
    }
    return MOVI_GPU_OK;
}

}  // namespace

extern "C" {

int movi_gpu_upload_colors(movi_gpu_ctx* ctx, const MoviGpuColorDesc* d) {
    if (!ctx || !d) return MOVI_GPU_ERR_ARG;
    if (!ctx->has_index) return fail(ctx, MOVI_GPU_ERR_INDEX, "upload the index before its colour table");
    if (!d->flat_colors || !d->doc_set_flat_inds || d->num_species == 0 || d->num_species >= 0xFFFFu)
        return fail(ctx, MOVI_GPU_ERR_INDEX, "empty colour table");
    // every colour must lie inside the table and name documents below num_species (the kernels index per-lane counters with them)
    for (uint64_t p = 0; p < d->flat_size;) {
        const uint64_t k = d->flat_colors[p];
        if (p + 1 + k > d->flat_size) return fail(ctx, MOVI_GPU_ERR_INDEX, "corrupt colour table: a colour runs past the end");
        for (uint64_t j = 0; j < k; j++)
            if (d->flat_colors[p + 1 + j] >= d->num_species) return fail(ctx, MOVI_GPU_ERR_INDEX, "corrupt colour table: document id out of range");
        p += 1 + k;
    }
    const uint8_t* inds = static_cast<const uint8_t*>(d->doc_set_flat_inds);
    for (uint64_t i = 0; i < ctx->ix.r; i++) {
        const uint8_t* e = inds + 5 * i;
        const uint64_t off = (uint64_t)e[0] | ((uint64_t)e[1] << 8) | ((uint64_t)e[2] << 16) | ((uint64_t)e[3] << 24) | ((uint64_t)e[4] << 32);
        // (an offset at or beyond the end means "colour dropped by compression", move_structure_query.cpp:395; one inside must start a colour:
        //  it does when it was written by flat_and_serialize_colors_vectors -- bounds are what protects the device)
        if (off < d->flat_size && off + 1 + d->flat_colors[off] > d->flat_size) return fail(ctx, MOVI_GPU_ERR_INDEX, "corrupt colour table: a row points into the tail of the table");
    }
    CU(cudaSetDevice(ctx->device));
    cudaFree(ctx->c_inds); cudaFree(ctx->c_flat); cudaFree(ctx->c_lk);
    ctx->c_inds = nullptr; ctx->c_flat = nullptr; ctx->c_lk = nullptr; ctx->c_species = 0;
    CU(cudaMalloc(&ctx->c_inds, ctx->ix.r * 5 + 16));
    CU(cudaMalloc(&ctx->c_flat, (d->flat_size + 8) * sizeof(uint16_t)));
    CU(cudaMemcpy(ctx->c_inds, inds, ctx->ix.r * 5, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->c_flat, d->flat_colors, d->flat_size * sizeof(uint16_t), cudaMemcpyHostToDevice));
    ctx->c_pvalue_ok = d->log_lens != nullptr;
    if (d->log_lens) {
        const double log4 = std::log(4.0);                                 // include/move_structure.hpp:315
        std::vector<double> lk(d->num_species);
        for (uint32_t i = 0; i < d->num_species; i++) lk[i] = d->log_lens[i] / log4;   // the quotient of move_structure_query.cpp:421
        CU(cudaMalloc(&ctx->c_lk, lk.size() * sizeof(double)));
        CU(cudaMemcpy(ctx->c_lk, lk.data(), lk.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    ctx->c_flat_size = d->flat_size; ctx->c_species = d->num_species;
    return MOVI_GPU_OK;
}

int movi_gpu_query_multi_classify(movi_gpu_ctx* ctx, const char* bases, const uint64_t* read_offsets, uint64_t n_reads,
                                  uint32_t flags, uint32_t min_len, int pvalue, MoviGpuColorResult* out, uint16_t* out_ml) {
    int rc = ext_ready(ctx, flags);
    if (rc) return rc;
    if (!ctx->c_species) return fail(ctx, MOVI_GPU_ERR_INDEX, "no colour table uploaded (movi_gpu_upload_colors)");
    if (pvalue && !ctx->c_pvalue_ok) return fail(ctx, MOVI_GPU_ERR_INDEX, "the colour table was uploaded without document lengths: no --pvalue-scoring");
    if (n_reads && !out) return fail(ctx, MOVI_GPU_ERR_ARG, "null result buffer");
    return ext_query(ctx, pvalue ? (EXT_COLOR | EXT_PVALUE) : EXT_COLOR, bases, read_offsets, n_reads, min_len, out, out_ml, nullptr);
}

int movi_gpu_upload_sampled_sa(movi_gpu_ctx* ctx, uint64_t rate, const uint64_t* sampled, uint64_t n_sampled, const uint64_t* all_p) {
    if (!ctx || !sampled || !all_p) return MOVI_GPU_ERR_ARG;
    if (!ctx->has_index) return fail(ctx, MOVI_GPU_ERR_INDEX, "upload the index before its sampled suffix array");
    if (rate == 0 || n_sampled == 0) return fail(ctx, MOVI_GPU_ERR_INDEX, "empty sampled suffix array");
    // all_p[i] = BWT offset of run i; every BWT offset / rate must have its sample
    const uint64_t r = ctx->ix.r;
    for (uint64_t i = 1; i < r; i++) if (all_p[i] <= all_p[i - 1]) return fail(ctx, MOVI_GPU_ERR_INDEX, "corrupt ssa.movi: run offsets must increase");
    if ((all_p[r - 1] + ctx->ix.n_last - 1) / rate >= n_sampled) return fail(ctx, MOVI_GPU_ERR_INDEX, "corrupt ssa.movi: fewer samples than the BWT needs");
    CU(cudaSetDevice(ctx->device));
    cudaFree(ctx->s_all_p); cudaFree(ctx->s_sampled); ctx->s_all_p = nullptr; ctx->s_sampled = nullptr; ctx->s_rate = 0;
    CU(cudaMalloc(&ctx->s_all_p, r * sizeof(uint64_t)));
    CU(cudaMalloc(&ctx->s_sampled, n_sampled * sizeof(uint64_t)));
    CU(cudaMemcpy(ctx->s_all_p, all_p, r * sizeof(uint64_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->s_sampled, sampled, n_sampled * sizeof(uint64_t), cudaMemcpyHostToDevice));
    ctx->s_rate = rate; ctx->s_n = n_sampled;
    return MOVI_GPU_OK;
}

int movi_gpu_query_pml_sa(movi_gpu_ctx* ctx, const char* bases, const uint64_t* read_offsets, uint64_t n_reads,
                          uint32_t flags, uint16_t* out_ml, uint64_t* out_sa) {
    int rc = ext_ready(ctx, flags);
    if (rc) return rc;
    if (!ctx->s_rate) return fail(ctx, MOVI_GPU_ERR_INDEX, "no sampled suffix array uploaded (movi_gpu_upload_sampled_sa)");
    if (n_reads && !out_sa) return fail(ctx, MOVI_GPU_ERR_ARG, "null result buffer");
    return ext_query(ctx, EXT_SA, bases, read_offsets, n_reads, 0, nullptr, out_ml, out_sa);
}

}  // extern "C"
