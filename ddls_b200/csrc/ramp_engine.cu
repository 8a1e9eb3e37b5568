// ramp_engine.cu -- host side of the C ABI declared in include/ramp_b200.h.
//
// Owns the HBM-resident state (templates, memo table, result slots, trace pool, per-episode tables,
// per-CTA scratch slabs) and launches the three kernels of a batched RampClusterEnvironment.step:
//   plan (memo) -> lookahead (persistent CTAs, one lookahead per CTA at a time) -> step (one thread per episode).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "ramp_kernels.cuh"
#include "ramp_env.cuh"

using namespace ramp;

namespace {

thread_local std::string g_last_error;

int set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return set_error(RAMP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

constexpr int MAX_EVENT_PAIRS = 64;
constexpr int RAMP_SMEM_CARVEOUT = 100;   // percent of the unified L1/shared array given to shared memory

// The lookahead kernel is instantiated for several CTA sizes; small CTAs (1-2 warps) keep more lookaheads
// resident per SM and waste fewer lanes on the small per-tick frontiers, large CTAs finish one lookahead sooner.
using LookaheadKernel = void (*)(const LookaheadArgs);
LookaheadKernel lookahead_kernel_for(int nt) {   // nt = threads per CTA = 32 x (independent lookahead warps per CTA)
    switch (nt) {
        case 32: return ramp_lookahead_kernel<1>;
        case 64: return ramp_lookahead_kernel<2>;
        case 128: return ramp_lookahead_kernel<4>;
        case 256: return ramp_lookahead_kernel<8>;
        default: return nullptr;
    }
}

// the dense shape of the warp kernel: smaller shared-memory frontiers, 16 instead of 12 lookahead warps per SM
constexpr int DENSE_F_CAP = 256, DENSE_OPS_CAP = 32, DENSE_NT = 128;
LookaheadKernel lookahead_dense_kernel() { return ramp_lookahead_kernel<DENSE_NT / 32, DENSE_F_CAP, DENSE_OPS_CAP>; }

LookaheadKernel lookahead_cta_kernel_for(int nt) {   // one CTA of nt threads per lookahead
    switch (nt) {
        case 64: return ramp_lookahead_cta_kernel<2>;
        case 256: return ramp_lookahead_cta_kernel<8>;
        case 128: return ramp_lookahead_cta_kernel<4>;
        default: return nullptr;
    }
}

struct HostTemplate {
    TemplateDev dev;             // copy of what sits in the device array
    void* blob = nullptr;        // single device allocation holding all arrays
    std::vector<unsigned char> bytes;  // canonical host bytes for exact-duplicate detection
    uint64_t hash = 0;
    void* res_blob = nullptr;    // device copy of the quotient blob (thread-per-lookahead kernel), null if not resident
};

}  // namespace

struct ramp_engine {
    ramp_config_t cfg{};
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    // templates
    std::vector<HostTemplate> templates;
    TemplateDev* d_templates = nullptr;
    uint64_t max_scratch = 0;
    int32_t max_w = 1, max_c = 1;
    int32_t par_cap = 0;         // bytes of shared-memory parent counters per lookahead
    // memo + results
    uint32_t memo_cap = 0;
    unsigned long long* d_memo_keys = nullptr;
    int32_t* d_memo_vals = nullptr;
    unsigned long long* d_memo_keys2 = nullptr;
    uint32_t memo_cap2 = 0;
    ResultSlots res{};
    int32_t n_slots = 0;
    TracePool pool{};
    // per-step
    WorkItem* d_items = nullptr;
    Counters* d_counters = nullptr;
    MemoStats* d_stats = nullptr;
    ramp_action_t* d_actions = nullptr;
    double* d_step_stats = nullptr;
    int32_t* d_n_cluster_steps = nullptr;
    double* d_ep_export = nullptr;
    // episode state
    EpisodeState ep{};
    ramp_arrival_t* d_arrivals = nullptr;
    int32_t* d_n_jobs_ep = nullptr;
    // lookahead scratch
    unsigned char* d_scratch = nullptr;
    uint64_t scratch_stride = 0;
    int scratch_grid = 0;
    int grid = 0;
    int nt = 128;                // threads per lookahead CTA = 32 x warps, one lookahead per warp (RAMP_LOOKAHEAD_THREADS overrides)
    int max_ctas_per_sm = 0;     // optional cap (RAMP_LOOKAHEAD_CTAS_PER_SM)
    size_t smem_bytes = 0;
    // CTA-per-lookahead variant (lower latency; used when a launch has fewer work items than warp slots)
    int cta_nt = 0;              // 0 = pick 128 or 64 threads per launch; RAMP_LOOKAHEAD_CTA_THREADS forces one
    int dense_grid = 0;          // resident CTAs of the dense warp-kernel shape
    size_t dense_smem_bytes = 0;
    double dense_factor = 2.0;   // the dense shape runs a step's lookaheads when there are more than dense_factor x warp slots
    int cta_grid = 0;            // resident CTAs of the 128-thread variant
    int cta64_grid = 0;          // resident CTAs of the 64-thread variant
    int cta256_grid = 0;         // resident CTAs of the 256-thread variant
    int cta_grid_for(int nt) const { return nt == 256 ? cta256_grid : nt == 128 ? cta_grid : cta64_grid; }
    size_t cta_smem_bytes = 0;
    int split_warp_nt = 32;      // threads per CTA of the warp kernel when it runs beside a CTA kernel
    size_t smem2_bytes = 0;      // dynamic shared memory of the 2-warp CTAs used beside a CTA kernel
    int use_cta256 = 1;          // RAMP_USE_CTA256=0 turns off: 256-thread CTAs for the big lookaheads when 8 n_big + n_small fit the warp slots
    double split_alpha = 1.3;    // 64-thread-CTA split while 2 n_big + n_small <= split_alpha x resident warp slots
    int debug = 0;               // RAMP_DEBUG=1 prints the per-step launch decisions to stderr
    int mode = 0;                // 0 auto, 1 warp-per-lookahead, 2 CTA-per-lookahead (RAMP_LOOKAHEAD_MODE)
    int32_t* h_n_work = nullptr; // pinned [4]
    int64_t big_threshold = 60000;   // N + E from which one CTA per lookahead beats one warp (RAMP_BIG_THRESHOLD); measured
                                     // on B200: N+E=35k 3.4 vs 3.8 ms, N+E=137k 8.4 vs 4.5 ms (profiles/r1_latency_by_degree.txt)
    WorkItem* d_items_big = nullptr;
    cudaStream_t stream2 = nullptr;  // big lookaheads run concurrently with the small ones
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // resident (quotient) templates: thread-per-lookahead kernel
    int use_resident = 1;        // RAMP_RESIDENT=0: register every template for the warp / CTA kernels only
    int use_quotient = 1;        // RAMP_QUOTIENT=0: resident blobs are built from the unfolded job (identity quotient)
    int32_t res_max_bytes = 96 * 1024;   // largest quotient blob that goes resident (RAMP_RESIDENT_MAX_KB)
    int n_resident = 0, n_nonresident = 0;
    int32_t res_tmpl_cap = 0, res_n_cap = 0, res_spill_ops = 0, res_spill_deps = 0;
    int res_grid = 0;
    size_t res_smem = 0;
    unsigned char* d_res_scratch = nullptr;
    uint64_t res_scratch_stride = 0;
    int res_scratch_grid = 0;
    WorkItem* d_items_res = nullptr;
    WorkItem* d_chunk_items = nullptr;   // [B][32]
    ChunkDesc* d_chunks = nullptr;       // [B]
    int32_t* d_tcount = nullptr;         // [max_templates + 1]
    int32_t* d_tbase = nullptr;
    int32_t* d_rank = nullptr;           // [B]
    TemplateHints* d_hints = nullptr;    // [max_templates]
    double* d_hint_jct = nullptr;        // [max_templates]
    // device-resident rollouts (ramp_env_*)
    bool has_env = false;
    EnvDev env{};
    std::vector<void*> env_allocs;
    int32_t* env_h_need = nullptr;       // pinned: [0] = count, [1] = error flag; [2], [3]: the same, read by ramp_env_read; [4] engine error episode
    void* env_h_mirror = nullptr;        // pinned host arrays (ramp_env_host_mirror)
    bool env_unchecked_decide = false;   // a ramp_env_decide without need_host_out has not been looked at yet
    // standalone lookahead buffers
    WorkItem* sa_chunk_items = nullptr;
    ChunkDesc* sa_chunks = nullptr;
    ResultSlots sa_res{};
    int32_t sa_cap = 0;
    WorkItem* sa_items = nullptr;
    Counters* sa_counters = nullptr;
    // instrumentation
    int64_t launches = 0;
    cudaEvent_t ev_a[MAX_EVENT_PAIRS]{}, ev_b[MAX_EVENT_PAIRS]{};
    int ev_pending = 0;
    double la_ms_total = 0.0;
    int64_t la_launches = 0;
    unsigned long long la_items_base = 0, la_bytes_base = 0, la_qbytes_base = 0;
    MemoStats memo_base{};       // device counters at the last ramp_reset (memo statistics are reported since the reset)
};

namespace {

int alloc_result_slots(ResultSlots& r, int32_t n) {
    CUDA_TRY(cudaMalloc(&r.jct, sizeof(double) * n));
    CUDA_TRY(cudaMalloc(&r.comm, sizeof(double) * n));
    CUDA_TRY(cudaMalloc(&r.comp, sizeof(double) * n));
    CUDA_TRY(cudaMalloc(&r.n_ticks, sizeof(int32_t) * n));
    CUDA_TRY(cudaMalloc(&r.status, sizeof(int32_t) * n));
    CUDA_TRY(cudaMalloc(&r.trace_off, sizeof(int64_t) * n));
    CUDA_TRY(cudaMalloc(&r.util, sizeof(double) * n));
    CUDA_TRY(cudaMalloc(&r.util_nmw, sizeof(int32_t) * n));
    return RAMP_OK;
}

void free_result_slots(ResultSlots& r) {
    cudaFree(r.jct); cudaFree(r.comm); cudaFree(r.comp); cudaFree(r.n_ticks); cudaFree(r.status); cudaFree(r.trace_off);
    cudaFree(r.util); cudaFree(r.util_nmw);
    r = ResultSlots{};
}

int resolve_events(ramp_engine* e) {
    for (int k = 0; k < e->ev_pending; ++k) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e->ev_a[k], e->ev_b[k]));
        e->la_ms_total += ms;
        e->la_launches++;
    }
    e->ev_pending = 0;
    return RAMP_OK;
}

// (re)allocates the per-CTA scratch slabs for the largest registered template and picks the grid
int ensure_scratch(ramp_engine* e) {
    const uint64_t trace_bytes = align_up((uint64_t)e->cfg.trace_cap * 12, 16);
    const uint64_t stride = align_up(std::max<uint64_t>(e->max_scratch, 16), 256) + align_up(trace_bytes, 256);
    const size_t smem = lookahead_smem_per_warp(e->max_w, e->max_c, e->par_cap) * (size_t)(e->nt / 32);
    if (smem > 200 * 1024)
        return set_error(RAMP_ERR_CAPACITY, "a template needs %zu B of shared memory for its worker/channel key arrays (max 200 KiB)", smem);
    if (smem != e->smem_bytes || e->grid == 0) {
        LookaheadKernel kern = lookahead_kernel_for(e->nt);
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // all lookahead kernels ask for the same L1/shared split: kernels with different carve-outs cannot share an SM,
        // which would serialise the CTA kernel and the warp kernel when they are launched side by side
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, RAMP_SMEM_CARVEOUT));
        int occ = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, e->nt, smem));
        if (occ < 1) occ = 1;
        if (e->max_ctas_per_sm > 0 && occ > e->max_ctas_per_sm) occ = e->max_ctas_per_sm;
        e->grid = e->sm_count * occ;     // persistent CTAs: a whole number of waves (148 SMs x resident CTAs per SM)
        e->smem_bytes = smem;
        // beside a CTA kernel the small lookaheads run in 2-warp CTAs: they fill the registers and shared memory the CTA
        // kernel leaves on every SM at a finer grain (the block scheduler spreads both kernels over all SMs)
        const size_t smem2 = lookahead_smem_per_warp(e->max_w, e->max_c, e->par_cap) * (size_t)(e->split_warp_nt / 32);
        LookaheadKernel kern2 = lookahead_kernel_for(e->split_warp_nt);
        CUDA_TRY(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        CUDA_TRY(cudaFuncSetAttribute(kern2, cudaFuncAttributePreferredSharedMemoryCarveout, RAMP_SMEM_CARVEOUT));
        e->smem2_bytes = smem2;
        const size_t smem_d = lookahead_smem_per_warp(e->max_w, e->max_c, e->par_cap, DENSE_F_CAP, DENSE_OPS_CAP) * (size_t)(DENSE_NT / 32);
        LookaheadKernel kern_d = lookahead_dense_kernel();
        CUDA_TRY(cudaFuncSetAttribute(kern_d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));
        CUDA_TRY(cudaFuncSetAttribute(kern_d, cudaFuncAttributePreferredSharedMemoryCarveout, RAMP_SMEM_CARVEOUT));
        int occ_d = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_d, kern_d, DENSE_NT, smem_d));
        e->dense_grid = e->sm_count * std::max(occ_d, 1);
        e->dense_smem_bytes = smem_d;
    }
    const size_t cta_smem = lookahead_cta_smem(e->max_w, e->max_c, e->par_cap);
    if (cta_smem > 200 * 1024)
        return set_error(RAMP_ERR_CAPACITY, "a template needs %zu B of shared memory (max 200 KiB)", cta_smem);
    if (cta_smem != e->cta_smem_bytes || e->cta_grid == 0) {
        for (int nt : {256, 128, 64}) {
            LookaheadKernel kern = lookahead_cta_kernel_for(nt);
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cta_smem));
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, RAMP_SMEM_CARVEOUT));
            int occ = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, cta_smem));
            if (occ < 1) occ = 1;
            (nt == 256 ? e->cta256_grid : nt == 128 ? e->cta_grid : e->cta64_grid) = e->sm_count * occ;
        }
        e->cta_smem_bytes = cta_smem;
    }
    // one slab per lookahead in flight; the warp kernel and one of the CTA kernels may run side by side
    const int n_slabs = std::max(e->grid * (e->nt / 32) + std::max(e->cta_grid, e->cta64_grid), e->dense_grid * (DENSE_NT / 32));
    if (stride != e->scratch_stride || n_slabs != e->scratch_grid || e->d_scratch == nullptr) {
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        if (e->d_scratch) cudaFree(e->d_scratch);
        e->d_scratch = nullptr;
        CUDA_TRY(cudaMalloc(&e->d_scratch, stride * (uint64_t)n_slabs));
        e->scratch_stride = stride;
        e->scratch_grid = n_slabs;
    }
    return RAMP_OK;
}

LookaheadArgs make_lookahead_args(ramp_engine* e, const WorkItem* items, Counters* counters, const ResultSlots& res,
                                  bool use_pool, MemoStats* stats) {
    LookaheadArgs a{};
    a.templates = e->d_templates;
    a.items = items;
    a.n_work = &counters->n_work;
    a.items_b = nullptr;
    a.n_work_b = nullptr;
    a.cursor = &counters->work_cursor;
    a.scratch = e->d_scratch;
    a.scratch_stride = e->scratch_stride;
    a.res = res;
    a.pool = e->pool;
    if (!use_pool) a.pool.top = nullptr;
    a.trace_cap = e->cfg.trace_cap;
    a.w_cap = e->max_w;
    a.c_cap = e->max_c;
    a.par_cap = e->par_cap;
    a.stats = stats;
    return a;
}

// Launches ONE lookahead kernel for n_items work items of a standalone run (ramp_run_lookaheads): one CTA per lookahead
// when there are big lookaheads among them and the items fit in about one wave of CTAs (latency-bound regime), one warp
// per lookahead otherwise (small lookaheads are fastest on one warp; many lookaheads need the warp kernel's density).
// Measured on B200 (ResNet-50-like job): degree 16 alone 2.7 ms on a 128-thread CTA, 4.0 ms on a 64-thread CTA, 4.5-5 ms on
// one warp; degree 2: 1.8 / 1.8 / 1.5 ms.
void launch_lookahead(ramp_engine* e, const LookaheadArgs& a, int n_items, int n_big, cudaStream_t st) {
    int cta_nt = 0;
    if (e->mode == 2) cta_nt = e->cta_nt ? e->cta_nt : (n_items <= e->cta_grid ? 128 : 64);
    else if (e->mode != 1 && n_big > 0) {
        if (e->cta_nt) { if (n_items <= e->cta_grid_for(e->cta_nt)) cta_nt = e->cta_nt; }
        else if (n_items <= e->cta_grid) cta_nt = 128;
        else if (n_items <= e->cta64_grid) cta_nt = 64;
    }
    if (cta_nt) {
        const int grid = std::max(1, std::min(e->cta_grid_for(cta_nt), n_items));
        lookahead_cta_kernel_for(cta_nt)<<<grid, cta_nt, e->cta_smem_bytes, st>>>(a);
    } else {
        const int wpb = e->nt / 32;
        const int grid = std::max(1, std::min(e->grid, (n_items + wpb - 1) / wpb));
        lookahead_kernel_for(e->nt)<<<grid, e->nt, e->smem_bytes, st>>>(a);
    }
}

uint64_t fnv1a(const unsigned char* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

// rank keys: larger key wins.  Sorting by (priority desc, index asc) and numbering from the top reproduces
// "iterate in sorted() order, replace only on strictly greater priority" (RCE:56-66, RCE:672-685).
void make_rank_keys(const int64_t* prio, int32_t n, std::vector<uint32_t>& key) {
    std::vector<int32_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return prio[a] > prio[b]; });
    key.resize(n);
    for (int32_t r = 0; r < n; ++r) key[order[r]] = (uint32_t)(n - r);
}


// ---- resident (quotient) templates ---------------------------------------------------------------------------
int bits_for_u64(uint64_t max_value) { int b = 1; while ((max_value >> b) != 0) ++b; return b; }

// Packs a quotient job (ramp_quotient.cpp) into the blob the thread-per-lookahead kernel bulk-copies into shared memory
// (layout: ResHeader, ramp_lookahead_thread.cuh).  Returns false when the job is not eligible: blob larger than
// max_bytes, class sizes / counters beyond 16 bits, or a dep word wider than 64 bits.
bool build_resident_blob(const ramp_lowered_job_t* j, const ramp_quotient_t& q, int32_t max_bytes, std::vector<unsigned char>& blob) {
    const int32_t N = q.n_ops, E = q.n_deps;
    if (N < 1 || q.n_workers > 0xFFFF || q.n_channels >= 0xFFFF) return false;
    if (!q.masks_valid) return false;
    std::vector<uint64_t> in_total(N, 0);
    uint32_t max_inc = 1;
    // keys only order entries: re-rank them densely so that the key and the group set share one 32-bit word
    std::vector<uint32_t> distinct(q.dep_key, q.dep_key + E);
    std::sort(distinct.begin(), distinct.end());
    distinct.erase(std::unique(distinct.begin(), distinct.end()), distinct.end());
    std::vector<uint32_t> dense_key(E);
    for (int32_t k = 0; k < E; ++k) {
        in_total[q.dep_dst[k]] += q.dep_inc[k];
        max_inc = std::max(max_inc, q.dep_inc[k]);
        dense_key[k] = (uint32_t)(std::lower_bound(distinct.begin(), distinct.end(), q.dep_key[k]) - distinct.begin()) + 1u;
    }
    for (int32_t c = 0; c < N; ++c)
        if (q.op_weight[c] > 0xFFFF || in_total[c] > 0xFFFF || q.op_threshold[c] > 0xFFFFFFFFu) return false;   // u16 counters never wrap
    const int kbits = bits_for_u64((uint64_t)distinct.size() + 1), cbits = std::max(q.n_channels, 1);
    const int ibits = bits_for_u64(max_inc), nbits = bits_for_u64((uint64_t)N);
    if (kbits + cbits > 32 || 1 + ibits + nbits > 32) return false;     // dep word = two 32-bit halves: key | group set, flow | inc | child
    std::vector<int32_t> in_deg(N, 0), src;
    for (int32_t k = 0; k < E; ++k) in_deg[q.dep_dst[k]]++;
    for (int32_t c = 0; c < N; ++c) if (in_deg[c] == 0) src.push_back(c);
    ResHeader h{};
    h.n_ops = N; h.n_deps = E; h.n_workers = q.n_workers; h.n_channels = q.n_channels;
    h.n_src = (int32_t)src.size(); h.num_training_steps = j->num_training_steps; h.orig_workers = j->n_workers;
    h.kmask = (uint32_t)((1ull << kbits) - 1ull); h.cmask = (uint32_t)((1ull << cbits) - 1ull); h.imask = (uint32_t)((1ull << ibits) - 1ull);
    h.cshift = kbits; h.fshift = 0; h.ishift = 1; h.dshift = 1 + ibits;
    size_t off = sizeof(ResHeader);
    const size_t off_rec = off;                 off += align_up((uint64_t)N * 16, 16);
    h.off_op_row = (int32_t)off;                off += align_up((uint64_t)N * 8, 16);
    h.off_op_thr = (int32_t)off;                off += align_up((uint64_t)N * 4, 16);
    h.off_dep_kd = (int32_t)off;                off += align_up((uint64_t)std::max(E, 1) * 8, 16);
    h.off_dep_rt = (int32_t)off;                off += align_up((uint64_t)std::max(E, 1) * 8, 16);
    h.off_src = (int32_t)off;                   off += align_up((uint64_t)std::max<size_t>(src.size(), 1) * 4, 16);
    if (off > (size_t)max_bytes) return false;
    h.total_bytes = (int32_t)off;
    blob.assign(off, 0);
    memcpy(blob.data(), &h, sizeof(h));
    struct OpRec { double cost; uint32_t key; uint32_t worker; };
    OpRec* rec = reinterpret_cast<OpRec*>(blob.data() + off_rec);
    int32_t* row = reinterpret_cast<int32_t*>(blob.data() + h.off_op_row);
    uint32_t* thr = reinterpret_cast<uint32_t*>(blob.data() + h.off_op_thr);
    unsigned long long* kd = reinterpret_cast<unsigned long long*>(blob.data() + h.off_dep_kd);
    double* rt = reinterpret_cast<double*>(blob.data() + h.off_dep_rt);
    for (int32_t c = 0; c < N; ++c) {
        rec[c] = OpRec{q.op_cost[c] + 0.0, q.op_key[c], q.op_worker[c] | (q.op_weight[c] << 16)};
        row[2 * c] = q.row_ptr[c]; row[2 * c + 1] = q.row_ptr[c + 1] - q.row_ptr[c];
        thr[c] = q.op_threshold[c];
    }
    for (int32_t k = 0; k < E; ++k) {
        const uint32_t lo = dense_key[k] | (uint32_t)(q.dep_group_mask[k] << h.cshift);
        const uint32_t hi = (q.dep_is_flow[k] ? 1u : 0u) | (q.dep_inc[k] << h.ishift) | ((uint32_t)q.dep_dst[k] << h.dshift);
        kd[k] = (unsigned long long)lo | ((unsigned long long)hi << 32);
        rt[k] = q.dep_run_time[k] + 0.0;
    }
    if (!src.empty()) memcpy(blob.data() + h.off_src, src.data(), sizeof(int32_t) * src.size());
    return true;
}

// the identity quotient: every op its own class (RAMP_QUOTIENT=0; lets the tests run the thread kernel on unfolded jobs)
int identity_quotient(const ramp_lowered_job_t* j, ramp_quotient_t* q) {
    memset(q, 0, sizeof(*q));
    const int32_t N = j->n_ops, E = j->n_deps;
    std::vector<uint32_t> ok, dk;
    make_rank_keys(j->op_prio, N, ok);
    make_rank_keys(j->dep_prio, E, dk);
    q->n_ops = N; q->n_deps = E; q->n_workers = j->n_workers; q->n_channels = j->n_channels;
    q->op_cost = (double*)malloc(sizeof(double) * N); q->op_key = (uint32_t*)malloc(4 * (size_t)N); q->op_worker = (uint32_t*)malloc(4 * (size_t)N);
    q->op_weight = (uint32_t*)malloc(4 * (size_t)N); q->op_threshold = (uint32_t*)malloc(4 * (size_t)N); q->row_ptr = (int32_t*)malloc(4 * ((size_t)N + 1));
    const size_t e1 = std::max(E, 1);
    q->dep_dst = (int32_t*)malloc(4 * e1); q->dep_run_time = (double*)malloc(8 * e1); q->dep_key = (uint32_t*)malloc(4 * e1);
    q->dep_channel = (uint32_t*)malloc(4 * e1); q->dep_is_flow = (uint8_t*)malloc(e1); q->dep_inc = (uint32_t*)malloc(4 * e1);
    q->dep_group_mask = (uint64_t*)malloc(8 * e1); q->merged = 0; q->masks_valid = j->n_channels <= 64 ? 1 : 0;
    q->op_class = (int32_t*)malloc(4 * (size_t)N); q->dep_entry = (int32_t*)malloc(4 * e1);
    for (int32_t i = 0; i < N; ++i) {
        q->op_cost[i] = j->op_cost[i]; q->op_key[i] = ok[i]; q->op_worker[i] = j->op_worker[i]; q->op_weight[i] = 1;
        q->op_threshold[i] = j->op_n_parents[i]; q->row_ptr[i] = j->row_ptr[i]; q->op_class[i] = i;
    }
    q->row_ptr[N] = j->row_ptr[N];
    for (int32_t k = 0; k < E; ++k) {
        q->dep_dst[k] = j->dep_dst[k]; q->dep_run_time[k] = j->dep_run_time[k]; q->dep_key[k] = dk[k];
        q->dep_channel[k] = (j->dep_channel[k] == RAMP_NO_CHANNEL) ? 0xFFFFFFFFu : (uint32_t)j->dep_channel[k];
        q->dep_group_mask[k] = (j->dep_channel[k] == RAMP_NO_CHANNEL || !q->masks_valid) ? 0ull : (1ull << j->dep_channel[k]);
        q->dep_is_flow[k] = j->dep_is_flow[k] ? 1 : 0; q->dep_inc[k] = 1; q->dep_entry[k] = k;
    }
    return RAMP_OK;
}

// shared memory / grid / HBM slabs of the thread-per-lookahead kernel for the resident templates registered so far
int ensure_thread_scratch(ramp_engine* e) {
    const size_t smem = thread_smem_bytes(e->res_tmpl_cap, e->res_n_cap);
    if (smem > 220 * 1024) return set_error(RAMP_ERR_CAPACITY, "resident templates need %zu B of shared memory (max 220 KiB)", smem);
    if (smem != e->res_smem || e->res_grid == 0) {
        CUDA_TRY(cudaFuncSetAttribute(ramp_lookahead_thread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(ramp_lookahead_thread_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, RAMP_SMEM_CARVEOUT));
        int occ = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ramp_lookahead_thread_kernel, 32, smem));
        if (occ < 1) occ = 1;
        // one chunk = up to 32 lookaheads; enough resident CTAs for every episode to miss at once, at most 2 per SM
        const int want = (e->cfg.n_episodes + 31) / 32 + 8;
        e->res_grid = std::max(1, std::min(std::min(e->sm_count * occ, e->sm_count * 2), want));
        e->res_smem = smem;
    }
    const uint64_t stride = thread_scratch_bytes(e->res_spill_ops, e->res_spill_deps, e->cfg.trace_cap);
    if (stride != e->res_scratch_stride || e->res_grid != e->res_scratch_grid || !e->d_res_scratch) {
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        if (e->d_res_scratch) cudaFree(e->d_res_scratch);
        e->d_res_scratch = nullptr;
        CUDA_TRY(cudaMalloc(&e->d_res_scratch, stride * (uint64_t)e->res_grid));
        e->res_scratch_stride = stride;
        e->res_scratch_grid = e->res_grid;
    }
    return RAMP_OK;
}

ThreadArgs make_thread_args(ramp_engine* e, const ChunkDesc* chunks, const int32_t* n_chunks, int32_t* cursor, const WorkItem* items,
                            const ResultSlots& res, const TracePool& pool, MemoStats* stats) {
    ThreadArgs a{};
    a.templates = e->d_templates; a.chunks = chunks; a.n_chunks = n_chunks; a.cursor = cursor; a.items = items;
    a.scratch = e->d_res_scratch; a.scratch_stride = e->res_scratch_stride;
    a.res = res; a.pool = pool; a.trace_cap = e->cfg.trace_cap;
    a.tmpl_cap = e->res_tmpl_cap; a.n_cap = e->res_n_cap; a.spill_ops = e->res_spill_ops; a.spill_deps = e->res_spill_deps;
    a.stats = stats; a.hints = e->d_hints; a.hint_jct = e->d_hint_jct;
    return a;
}

}  // namespace

// hooks for the other translation units of the library (ramp_policy.cu); not part of the C ABI
int ramp_internal_set_error(int code, const char* msg) { g_last_error = msg; return code; }
cudaStream_t ramp_internal_stream(ramp_engine_t* e) { return e->stream; }
int ramp_internal_device(ramp_engine_t* e) { return e->cfg.device; }
void ramp_internal_count_launches(ramp_engine_t* e, int n) { e->launches += n; }

extern "C" {

const char* ramp_last_error(void) { return g_last_error.c_str(); }

int ramp_engine_create(const ramp_config_t* cfg_in, ramp_engine_t** out) {
    if (!cfg_in || !out) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    ramp_config_t cfg = *cfg_in;
    if (cfg.n_episodes < 1 || cfg.max_jobs < 1 || cfg.n_cluster_workers < 1)
        return set_error(RAMP_ERR_BAD_ARG, "n_episodes, max_jobs and n_cluster_workers must be >= 1");
    if (cfg.max_running < 1) cfg.max_running = cfg.n_cluster_workers;
    if (cfg.max_templates < 1) cfg.max_templates = 1024;
    if (cfg.trace_cap < 1) cfg.trace_cap = 16384;
    if (cfg.job_queue_capacity < 0) cfg.job_queue_capacity = 10;
    if (cfg.machine_epsilon == 0.0) cfg.machine_epsilon = 1e-7;
    if (cfg.memo_capacity_log2 <= 0) {
        int lg = 10;
        while ((1ll << lg) < (long long)cfg.n_episodes * 16 && lg < 26) ++lg;
        cfg.memo_capacity_log2 = lg;
    }
    CUDA_TRY(cudaSetDevice(cfg.device));
    ramp_engine* e = new ramp_engine();
    e->cfg = cfg;
    if (const char* v = getenv("RAMP_LOOKAHEAD_THREADS")) {
        const int nt = atoi(v);
        if (lookahead_kernel_for(nt) == nullptr) { delete e; return set_error(RAMP_ERR_BAD_ARG, "RAMP_LOOKAHEAD_THREADS must be 32, 64, 128 or 256"); }
        e->nt = nt;
    }
    if (const char* v = getenv("RAMP_LOOKAHEAD_CTAS_PER_SM")) e->max_ctas_per_sm = atoi(v);
    if (const char* v = getenv("RAMP_LOOKAHEAD_CTA_THREADS")) {
        const int nt = atoi(v);
        if (lookahead_cta_kernel_for(nt) == nullptr) { delete e; return set_error(RAMP_ERR_BAD_ARG, "RAMP_LOOKAHEAD_CTA_THREADS must be 64, 128 or 256"); }
        e->cta_nt = nt;
    }
    if (const char* v = getenv("RAMP_LOOKAHEAD_MODE")) {
        // warp / cta: every template goes to that kernel (no resident quotient blobs); thread_unfolded: the thread kernel on
        // the unfolded job (identity quotient); anything else: automatic (resident whenever the quotient blob fits)
        e->mode = !strcmp(v, "warp") ? 1 : !strcmp(v, "cta") ? 2 : 0;
        if (e->mode != 0) e->use_resident = 0;
        if (!strcmp(v, "thread_unfolded")) e->use_quotient = 0;
    }
    if (const char* v = getenv("RAMP_BIG_THRESHOLD")) e->big_threshold = atoll(v);
    if (const char* v = getenv("RAMP_DEBUG")) e->debug = atoi(v);
    if (const char* v = getenv("RAMP_RESIDENT")) e->use_resident = atoi(v);
    if (const char* v = getenv("RAMP_QUOTIENT")) e->use_quotient = atoi(v);
    if (const char* v = getenv("RAMP_RESIDENT_MAX_KB")) e->res_max_bytes = atoi(v) * 1024;
    if (const char* v = getenv("RAMP_SPLIT_ALPHA")) e->split_alpha = atof(v);
    if (const char* v = getenv("RAMP_USE_CTA256")) e->use_cta256 = atoi(v);
    if (const char* v = getenv("RAMP_DENSE_FACTOR")) e->dense_factor = atof(v);
    if (const char* v = getenv("RAMP_SPLIT_WARP_THREADS")) {
        const int nt = atoi(v);
        if (lookahead_kernel_for(nt) == nullptr) { delete e; return set_error(RAMP_ERR_BAD_ARG, "RAMP_SPLIT_WARP_THREADS must be 32, 64, 128 or 256"); }
        e->split_warp_nt = nt;
    }
    cudaDeviceProp prop{};
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg.device));
    e->sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    const int B = cfg.n_episodes;

    CUDA_TRY(cudaMalloc(&e->d_templates, sizeof(TemplateDev) * cfg.max_templates));
    e->memo_cap = 1u << cfg.memo_capacity_log2;
    CUDA_TRY(cudaMalloc(&e->d_memo_keys, sizeof(unsigned long long) * e->memo_cap));
    CUDA_TRY(cudaMemset(e->d_memo_keys, 0, sizeof(unsigned long long) * e->memo_cap));
    CUDA_TRY(cudaMalloc(&e->d_memo_vals, sizeof(int32_t) * e->memo_cap));
    e->memo_cap2 = 1;
    while (e->memo_cap2 < (uint32_t)cfg.max_templates * 2u) e->memo_cap2 <<= 1;
    CUDA_TRY(cudaMalloc(&e->d_memo_keys2, sizeof(unsigned long long) * e->memo_cap2));
    CUDA_TRY(cudaMemset(e->d_memo_keys2, 0, sizeof(unsigned long long) * e->memo_cap2));
    e->n_slots = (int32_t)e->memo_cap + B + (int32_t)e->memo_cap2;
    if (alloc_result_slots(e->res, e->n_slots) != RAMP_OK) return RAMP_ERR_CUDA;
    CUDA_TRY(cudaMemset(e->res.status, 0, sizeof(int32_t) * e->n_slots));

    // trace pool: exact-size allocations; 4 Mi entries at least, else enough for 8 un-memoised lookaheads of 2,048 ticks per
    // episode between two resets (the pool is reset with the memo)
    e->pool.len = std::max<uint64_t>(4ull << 20, (uint64_t)B * 8ull * 2048ull);      // B=1: 48 MiB; B=4096: 805 MiB
    CUDA_TRY(cudaMalloc(&e->pool.n_active, sizeof(int32_t) * e->pool.len));
    CUDA_TRY(cudaMalloc(&e->pool.tick, sizeof(double) * e->pool.len));
    CUDA_TRY(cudaMalloc(&e->pool.top, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(e->pool.top, 0, sizeof(unsigned long long)));

    CUDA_TRY(cudaMalloc(&e->d_items, sizeof(WorkItem) * B));
    CUDA_TRY(cudaMalloc(&e->d_items_big, sizeof(WorkItem) * B));
    CUDA_TRY(cudaMalloc(&e->d_items_res, sizeof(WorkItem) * B));
    CUDA_TRY(cudaMalloc(&e->d_chunk_items, sizeof(WorkItem) * (size_t)B * 32));
    CUDA_TRY(cudaMalloc(&e->d_chunks, sizeof(ChunkDesc) * B));
    CUDA_TRY(cudaMalloc(&e->d_tcount, sizeof(int32_t) * ((size_t)cfg.max_templates + 1)));
    CUDA_TRY(cudaMemset(e->d_tcount, 0, sizeof(int32_t) * ((size_t)cfg.max_templates + 1)));
    CUDA_TRY(cudaMalloc(&e->d_tbase, sizeof(int32_t) * ((size_t)cfg.max_templates + 1)));
    CUDA_TRY(cudaMalloc(&e->d_rank, sizeof(int32_t) * B));
    CUDA_TRY(cudaMalloc(&e->d_hints, sizeof(TemplateHints) * (size_t)cfg.max_templates));
    CUDA_TRY(cudaMemset(e->d_hints, 0, sizeof(TemplateHints) * (size_t)cfg.max_templates));
    CUDA_TRY(cudaMalloc(&e->d_hint_jct, sizeof(double) * (size_t)cfg.max_templates));
    CUDA_TRY(cudaMemset(e->d_hint_jct, 0, sizeof(double) * (size_t)cfg.max_templates));
    CUDA_TRY(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    CUDA_TRY(cudaMalloc(&e->d_counters, sizeof(Counters)));
    CUDA_TRY(cudaMemset(e->d_counters, 0, sizeof(Counters)));
    CUDA_TRY(cudaMalloc(&e->d_stats, sizeof(MemoStats)));
    CUDA_TRY(cudaMemset(e->d_stats, 0, sizeof(MemoStats)));
    CUDA_TRY(cudaMalloc(&e->d_actions, sizeof(ramp_action_t) * B));
    CUDA_TRY(cudaMalloc(&e->d_step_stats, sizeof(double) * RAMP_STEP_STATS_LEN * B));
    CUDA_TRY(cudaMalloc(&e->d_n_cluster_steps, sizeof(int32_t) * B));
    CUDA_TRY(cudaMalloc(&e->d_ep_export, sizeof(double) * RAMP_EP_LEN * B));
    CUDA_TRY(cudaMallocHost(&e->h_n_work, sizeof(int32_t) * 4));

    EpisodeState& ep = e->ep;
    ep.B = B; ep.max_running = cfg.max_running; ep.max_jobs = cfg.max_jobs; ep.n_jobs = 0;
    ep.n_cluster_workers = cfg.n_cluster_workers; ep.queue_capacity = cfg.job_queue_capacity;
    ep.eps = cfg.machine_epsilon; ep.max_sim_time = cfg.max_simulation_run_time;
    CUDA_TRY(cudaMalloc(&ep.ef, sizeof(double) * EF_COUNT * B));
    CUDA_TRY(cudaMalloc(&ep.ei, sizeof(int32_t) * EI_COUNT * B));
    CUDA_TRY(cudaMalloc(&ep.rf, sizeof(double) * RF_COUNT * (size_t)cfg.max_running * B));
    CUDA_TRY(cudaMalloc(&ep.ri, sizeof(int32_t) * RI_COUNT * (size_t)cfg.max_running * B));
    CUDA_TRY(cudaMalloc(&ep.rec, sizeof(ramp_job_record_t) * (size_t)cfg.max_jobs * B));
    CUDA_TRY(cudaMalloc(&e->d_arrivals, sizeof(ramp_arrival_t) * (size_t)cfg.max_jobs * B));
    CUDA_TRY(cudaMemset(ep.ef, 0, sizeof(double) * EF_COUNT * B));
    CUDA_TRY(cudaMemset(ep.ei, 0, sizeof(int32_t) * EI_COUNT * B));
    CUDA_TRY(cudaMemset(ep.rec, 0, sizeof(ramp_job_record_t) * (size_t)cfg.max_jobs * B));
    ep.arr = e->d_arrivals;
    CUDA_TRY(cudaMalloc(&e->d_n_jobs_ep, sizeof(int32_t) * B));
    CUDA_TRY(cudaMemset(e->d_n_jobs_ep, 0, sizeof(int32_t) * B));
    ep.n_jobs_ep = e->d_n_jobs_ep;

    for (int k = 0; k < MAX_EVENT_PAIRS; ++k) {
        CUDA_TRY(cudaEventCreate(&e->ev_a[k]));
        CUDA_TRY(cudaEventCreate(&e->ev_b[k]));
    }
    *out = e;
    return RAMP_OK;
}

int ramp_engine_destroy(ramp_engine_t* e) {
    if (!e) return RAMP_OK;
    cudaSetDevice(e->cfg.device);
    cudaStreamSynchronize(e->stream);
    for (auto& t : e->templates) { cudaFree(t.blob); cudaFree(t.res_blob); }
    for (void* pa : e->env_allocs) cudaFree(pa);
    cudaFree(e->ep.tick_util); cudaFree(e->ep.tick_util_n);
    if (e->env_h_need) cudaFreeHost(e->env_h_need);
    if (e->env_h_mirror) cudaFreeHost(e->env_h_mirror);
    cudaFree(e->d_items_res); cudaFree(e->d_chunk_items); cudaFree(e->d_chunks); cudaFree(e->d_tcount); cudaFree(e->d_tbase); cudaFree(e->d_rank); cudaFree(e->d_hints); cudaFree(e->d_hint_jct);
    cudaFree(e->d_res_scratch); cudaFree(e->sa_chunk_items); cudaFree(e->sa_chunks);
    cudaFree(e->d_templates); cudaFree(e->d_memo_keys); cudaFree(e->d_memo_vals); cudaFree(e->d_memo_keys2);
    free_result_slots(e->res); free_result_slots(e->sa_res);
    cudaFree(e->pool.n_active); cudaFree(e->pool.tick); cudaFree(e->pool.top);
    cudaFree(e->d_items); cudaFree(e->d_items_big); cudaStreamDestroy(e->stream2); cudaEventDestroy(e->ev_fork); cudaEventDestroy(e->ev_join); cudaFree(e->d_counters); cudaFree(e->d_stats); cudaFree(e->d_actions);
    cudaFree(e->d_step_stats); cudaFree(e->d_n_cluster_steps); cudaFree(e->d_ep_export);
    cudaFree(e->ep.ef); cudaFree(e->ep.ei); cudaFree(e->ep.rf); cudaFree(e->ep.ri); cudaFree(e->ep.rec);
    cudaFree(e->d_arrivals); cudaFree(e->d_n_jobs_ep); cudaFree(e->d_scratch); cudaFree(e->sa_items); cudaFree(e->sa_counters); cudaFreeHost(e->h_n_work);
    for (int k = 0; k < MAX_EVENT_PAIRS; ++k) { cudaEventDestroy(e->ev_a[k]); cudaEventDestroy(e->ev_b[k]); }
    cudaStreamDestroy(e->stream);
    delete e;
    return RAMP_OK;
}

void* ramp_engine_stream(ramp_engine_t* e) { return e ? (void*)e->stream : nullptr; }

int ramp_template_count(ramp_engine_t* e) { return e ? (int)e->templates.size() : 0; }

int ramp_register_template(ramp_engine_t* e, const ramp_lowered_job_t* j, int32_t* id_out) {
    if (!e || !j || !id_out) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    if ((int)e->templates.size() >= e->cfg.max_templates)
        return set_error(RAMP_ERR_CAPACITY, "max_templates (%d) reached", e->cfg.max_templates);
    const int32_t N = j->n_ops, E = j->n_deps, W = j->n_workers, C = j->n_channels;
    if (N < 1 || E < 0 || W < 1 || C < 0) return set_error(RAMP_ERR_BAD_ARG, "bad template sizes N=%d E=%d W=%d C=%d", N, E, W, C);
    if (W > 0xFFFF || C >= 0xFFFF) return set_error(RAMP_ERR_BAD_ARG, "too many mounted workers/channels");
    if (j->model_id < 0 || j->model_id > 0xFFFF || j->degree < 0 || j->degree > 0xFFFF)
        return set_error(RAMP_ERR_BAD_ARG, "model_id and degree must fit 16 bits");
    // ---- validate (the reference would KeyError / misbehave on these) ----
    if (j->row_ptr[0] != 0 || j->row_ptr[N] != E) return set_error(RAMP_ERR_BAD_ARG, "row_ptr is not a CSR offset array");
    for (int32_t i = 0; i < N; ++i) {
        if (j->row_ptr[i + 1] < j->row_ptr[i]) return set_error(RAMP_ERR_BAD_ARG, "row_ptr not monotone at %d", i);
        if (j->op_worker[i] >= W) return set_error(RAMP_ERR_BAD_ARG, "op %d on worker %d >= n_workers %d", i, j->op_worker[i], W);
        if (!(j->op_cost[i] >= 0.0)) return set_error(RAMP_ERR_BAD_ARG, "op %d has a negative or NaN cost", i);
    }
    std::vector<int32_t> in_deg(N, 0);
    for (int32_t k = 0; k < E; ++k) {
        if (j->dep_dst[k] < 0 || j->dep_dst[k] >= N) return set_error(RAMP_ERR_BAD_ARG, "dep %d has dst out of range", k);
        if (j->dep_channel[k] != RAMP_NO_CHANNEL && j->dep_channel[k] >= C)
            return set_error(RAMP_ERR_BAD_ARG, "dep %d on channel %d >= n_channels %d", k, j->dep_channel[k], C);
        if (!(j->dep_run_time[k] >= 0.0)) return set_error(RAMP_ERR_BAD_ARG, "dep %d has a negative or NaN run time", k);
        // RCE:542-560 zeroes the run time of every non-flow dep when the job is mounted; with a non-zero one the reference's
        // zero-length ticks (RCE:412-422) would never complete it and _run_lookahead would spin forever
        if (!j->dep_is_flow[k] && j->dep_run_time[k] != 0.0)
            return set_error(RAMP_ERR_BAD_ARG, "non-flow dep %d has a non-zero run time (RCE:542-560 zeroes it)", k);
        in_deg[j->dep_dst[k]]++;
    }
    for (int32_t i = 0; i < N; ++i)
        if ((int32_t)j->op_n_parents[i] > in_deg[i])
            return set_error(RAMP_ERR_BAD_ARG, "op %d has n_parents %d > in-degree %d", i, (int)j->op_n_parents[i], in_deg[i]);
    // ---- derive ----
    std::vector<uint32_t> op_key, dep_key;
    make_rank_keys(j->op_prio, N, op_key);
    make_rank_keys(j->dep_prio, E, dep_key);
    std::vector<int32_t> src;
    for (int32_t i = 0; i < N; ++i) if (in_deg[i] == 0) src.push_back(i);
    std::vector<double> op_cost(j->op_cost, j->op_cost + N), dep_rt(j->dep_run_time, j->dep_run_time + E);
    for (auto& x : op_cost) x = x + 0.0;   // -0.0 -> +0.0 so the u64 bit pattern orders like the value
    for (auto& x : dep_rt) x = x + 0.0;

    // ---- records in the layout the tick loop streams (see TemplateDev) ----
    struct OpRec { double cost; uint32_t key; uint32_t worker; };
    static_assert(sizeof(OpRec) == 16, "op record must be 16 bytes");
    std::vector<OpRec> op_rec(N);
    std::vector<int32_t> op_row((size_t)N * 2, 0);
    for (int32_t i = 0; i < N; ++i) {
        op_rec[i] = OpRec{op_cost[i], op_key[i], (uint32_t)j->op_worker[i]};
        op_row[(size_t)i * 2] = j->row_ptr[i];
        op_row[(size_t)i * 2 + 1] = j->row_ptr[i + 1] - j->row_ptr[i];
    }
    int32_t max_in_deg = 0;
    for (int32_t i = 0; i < N; ++i) max_in_deg = std::max(max_in_deg, in_deg[i]);
    const bool par_in_smem = (max_in_deg <= 255 && N <= 8192);       // byte parent counters in shared memory
    // the whole dep in one word for the kernels' 16-byte frontier entries (TemplateDev::dep_kd)
    auto bits_for = [](uint64_t max_value) { int b = 1; while ((max_value >> b) != 0) ++b; return b; };
    const int kbits = bits_for((uint64_t)std::max(E, 1));                 // keys are 1..E, 0 = "none"
    const int cbits = bits_for((uint64_t)C + 1);                          // channels 0..C-1, all ones = "none"
    const int nbits = bits_for((uint64_t)N);
    if (kbits + cbits + 9 + nbits > 64)
        return set_error(RAMP_ERR_CAPACITY, "template too large for the packed dep word: %d key + %d channel + 9 + %d op bits > 64",
                         kbits, cbits, nbits);
    const uint32_t kd_kmask = (uint32_t)((1ull << kbits) - 1ull), kd_cmask = (uint32_t)((1ull << cbits) - 1ull);
    const int kd_cshift = kbits, kd_fshift = kbits + cbits, kd_dshift = kbits + cbits + 9;
    std::vector<unsigned long long> dep_kd(E);
    for (int32_t k = 0; k < E; ++k) {
        const unsigned long long chan = (j->dep_channel[k] == RAMP_NO_CHANNEL) ? (unsigned long long)kd_cmask : (unsigned long long)j->dep_channel[k];
        dep_kd[k] = (unsigned long long)dep_key[k] | (chan << kd_cshift)
                    | ((unsigned long long)(j->dep_is_flow[k] ? 1 : 0) << kd_fshift)
                    | (par_in_smem ? ((unsigned long long)j->op_n_parents[j->dep_dst[k]] << (kd_fshift + 1)) : 0ull)
                    | ((unsigned long long)j->dep_dst[k] << kd_dshift);
    }

    // ---- pack one blob ----
    struct Seg { const void* p; size_t bytes; size_t off; };
    Seg segs[6] = {
        {op_rec.data(), sizeof(OpRec) * (size_t)N, 0}, {j->op_n_parents, sizeof(uint16_t) * (size_t)N, 0},
        {op_row.data(), sizeof(int32_t) * op_row.size(), 0}, {dep_kd.data(), sizeof(unsigned long long) * (size_t)E, 0},
        {dep_rt.data(), sizeof(double) * (size_t)E, 0}, {src.data(), sizeof(int32_t) * src.size(), 0}};
    size_t total = 0;
    for (auto& s : segs) { s.off = total; total += align_up(std::max<size_t>(s.bytes, 1), 256); }
    HostTemplate ht;
    ht.bytes.assign(total + 8 * sizeof(int32_t), 0);
    for (auto& s : segs) if (s.bytes) memcpy(ht.bytes.data() + s.off, s.p, s.bytes);
    int32_t hdr[8] = {N, E, W, C, j->num_training_steps, 0, 0, (int32_t)src.size()};
    memcpy(ht.bytes.data() + total, hdr, sizeof(hdr));
    ht.hash = fnv1a(ht.bytes.data(), ht.bytes.size());
    int32_t canon = (int32_t)e->templates.size();
    for (size_t t = 0; t < e->templates.size(); ++t)
        if (e->templates[t].hash == ht.hash && e->templates[t].bytes == ht.bytes) { canon = e->templates[t].dev.canon_id; break; }

    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaMalloc(&ht.blob, total));
    CUDA_TRY(cudaMemcpy(ht.blob, ht.bytes.data(), total, cudaMemcpyHostToDevice));
    unsigned char* base = (unsigned char*)ht.blob;
    TemplateDev& d = ht.dev;
    d.n_ops = N; d.n_deps = E; d.n_workers = W; d.n_channels = C;
    d.num_training_steps = j->num_training_steps; d.model_id = j->model_id; d.degree = j->degree;
    d.n_src = (int32_t)src.size(); d.canon_id = canon;
    d.size_class = ((int64_t)N + (int64_t)E >= e->big_threshold) ? 1 : 0;
    d.par_in_smem = par_in_smem ? 1 : 0;
    d._pad0 = 0;
    d.op_rec = (const int4*)(base + segs[0].off); d.op_n_parents = (const uint16_t*)(base + segs[1].off);
    d.op_row = (const int2*)(base + segs[2].off); d.dep_kd = (const unsigned long long*)(base + segs[3].off);
    d.dep_rt = (const double*)(base + segs[4].off); d.src_ops = (const int32_t*)(base + segs[5].off);
    d.kd_kmask = kd_kmask; d.kd_cmask = kd_cmask; d.kd_cshift = kd_cshift; d.kd_fshift = kd_fshift; d.kd_dshift = kd_dshift; d._pad1 = 0;
    d.scratch_bytes = scratch_bytes_for(N, E);
    d.algorithmic_bytes_static = 20ull * (uint64_t)N + 19ull * (uint64_t)E + 24ull;
    // ---- symmetry quotient -> resident blob for the thread-per-lookahead kernel ----
    d.res_blob = nullptr; d.res_bytes = 0; d.res_n_ops = 0; d.res_n_deps = 0; d._pad2 = 0;
    if (e->use_resident) {
        ramp_quotient_t q{};
        const int qrc = e->use_quotient ? ramp_quotient_template(j, &q) : identity_quotient(j, &q);
        if (qrc != RAMP_OK) { cudaFree(ht.blob); return set_error(qrc, "ramp_quotient_template failed (%d)", qrc); }
        std::vector<unsigned char> rblob;
        if (build_resident_blob(j, q, e->res_max_bytes, rblob)) {
            cudaError_t ce = cudaMalloc(&ht.res_blob, rblob.size());
            if (ce == cudaSuccess) ce = cudaMemcpy(ht.res_blob, rblob.data(), rblob.size(), cudaMemcpyHostToDevice);
            if (ce != cudaSuccess) { ramp_free_quotient(&q); cudaFree(ht.blob); return set_error(RAMP_ERR_CUDA, "resident blob upload failed: %s", cudaGetErrorString(ce)); }
            d.res_blob = (const unsigned char*)ht.res_blob; d.res_bytes = (int32_t)rblob.size();
            d.res_n_ops = q.n_ops; d.res_n_deps = q.n_deps;
            d.size_class = 2;
            e->res_tmpl_cap = std::max(e->res_tmpl_cap, (int32_t)align_up((uint64_t)rblob.size(), 128));
            e->res_n_cap = std::max(e->res_n_cap, (int32_t)align_up((uint64_t)q.n_ops, 2));
            e->res_spill_ops = std::max(e->res_spill_ops, q.n_ops);
            e->res_spill_deps = std::max(e->res_spill_deps, std::max(q.n_deps, 1));
        }
        if (e->debug) fprintf(stderr, "[ramp] template %d: N=%d E=%d W=%d C=%d -> quotient N=%d E=%d W=%d C=%d, %s (%zu B)\n", (int)e->templates.size(),
                              N, E, W, C, q.n_ops, q.n_deps, q.n_workers, q.n_channels, d.res_blob ? "resident" : "not resident", rblob.size());
        ramp_free_quotient(&q);
    }
    if (d.size_class == 2) e->n_resident++; else e->n_nonresident++;
    const int32_t id = (int32_t)e->templates.size();
    CUDA_TRY(cudaMemcpy(e->d_templates + id, &d, sizeof(TemplateDev), cudaMemcpyHostToDevice));
    if (d.size_class != 2) {       // the warp / CTA kernels' slabs and shared-memory tables are sized by the jobs that use them
        e->max_scratch = std::max(e->max_scratch, d.scratch_bytes);
        e->max_w = std::max(e->max_w, W);
        e->max_c = std::max(e->max_c, std::max(C, 1));
        if (d.par_in_smem) e->par_cap = std::max(e->par_cap, (int32_t)align_up((uint64_t)N, 16));
    }
    e->templates.push_back(std::move(ht));
    *id_out = id;
    return RAMP_OK;
}

int ramp_reset(ramp_engine_t* e, const ramp_arrival_t* arrivals, int32_t n_jobs) {
    if (!e || !arrivals) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    if (n_jobs < 1 || n_jobs > e->cfg.max_jobs) return set_error(RAMP_ERR_BAD_ARG, "n_jobs %d not in [1, max_jobs=%d]", n_jobs, e->cfg.max_jobs);
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    const int B = e->cfg.n_episodes;
    if (n_jobs == e->cfg.max_jobs) {
        CUDA_TRY(cudaMemcpyAsync(e->d_arrivals, arrivals, sizeof(ramp_arrival_t) * (size_t)n_jobs * B, cudaMemcpyHostToDevice, e->stream));
    } else {
        CUDA_TRY(cudaMemcpy2DAsync(e->d_arrivals, sizeof(ramp_arrival_t) * (size_t)e->cfg.max_jobs, arrivals,
                                   sizeof(ramp_arrival_t) * (size_t)n_jobs, sizeof(ramp_arrival_t) * (size_t)n_jobs, B,
                                   cudaMemcpyHostToDevice, e->stream));
    }
    e->ep.n_jobs = n_jobs;
    {
        std::vector<int32_t> nj(B, n_jobs);
        CUDA_TRY(cudaMemcpyAsync(e->d_n_jobs_ep, nj.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));     // nj goes out of scope
    }
    // memo is per env instance per episode: cleared on reset (RCE:269-275)
    CUDA_TRY(cudaMemsetAsync(e->d_memo_keys, 0, sizeof(unsigned long long) * e->memo_cap, e->stream));
    // the batch-wide cache of RAMP_MEMO_SHARED (level-2 keys, its result slots and traces) is a pure function of the
    // lowered job and survives the reset; every other mode starts from an empty trace pool
    if (e->cfg.memo_mode != RAMP_MEMO_SHARED) CUDA_TRY(cudaMemsetAsync(e->pool.top, 0, sizeof(unsigned long long), e->stream));
    CUDA_TRY(cudaMemcpyAsync(&e->memo_base, e->d_stats, sizeof(MemoStats), cudaMemcpyDeviceToHost, e->stream));   // counters stay cumulative
    CUDA_TRY(cudaMemsetAsync(e->d_counters, 0, sizeof(Counters), e->stream));
    ramp_reset_kernel<<<(B + 127) / 128, 128, 0, e->stream>>>(e->ep);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return RAMP_OK;
}

int ramp_set_arrivals(ramp_engine_t* e, int32_t episode, int32_t first_job, const ramp_arrival_t* rows, int32_t n) {
    if (!e || !rows) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    if (episode < 0 || episode >= e->cfg.n_episodes || first_job < 0 || n < 0 || first_job + n > e->cfg.max_jobs)
        return set_error(RAMP_ERR_BAD_ARG, "arrival rows [%d, %d) of episode %d out of range (max_jobs=%d)", first_job, first_job + n, episode, e->cfg.max_jobs);
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(e->d_arrivals + (size_t)episode * e->cfg.max_jobs + first_job, rows, sizeof(ramp_arrival_t) * (size_t)n,
                             cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return RAMP_OK;
}

int ramp_set_job_count(ramp_engine_t* e, int32_t episode, int32_t n_jobs) {
    if (!e) return set_error(RAMP_ERR_BAD_ARG, "null engine");
    if (episode < 0 || episode >= e->cfg.n_episodes || n_jobs < 0 || n_jobs > e->cfg.max_jobs)
        return set_error(n_jobs > e->cfg.max_jobs ? RAMP_ERR_CAPACITY : RAMP_ERR_BAD_ARG,
                         "job count %d of episode %d out of range (max_jobs=%d)", n_jobs, episode, e->cfg.max_jobs);
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(e->d_n_jobs_ep + episode, &n_jobs, sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return RAMP_OK;
}

int ramp_set_limits(ramp_engine_t* e, double max_sim_time, int32_t queue_capacity) {
    if (!e) return set_error(RAMP_ERR_BAD_ARG, "null engine");
    if (queue_capacity < 0 || !(max_sim_time > 0.0)) return set_error(RAMP_ERR_BAD_ARG, "bad limits");
    e->cfg.max_simulation_run_time = max_sim_time; e->cfg.job_queue_capacity = queue_capacity;
    e->ep.max_sim_time = max_sim_time; e->ep.queue_capacity = queue_capacity;
    return RAMP_OK;
}

int ramp_step_device(ramp_engine_t* e, const ramp_action_t* d_actions, int32_t fuse, double* d_stats_out, int32_t* d_ncs_out) {
    if (!e || !d_actions) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    if (e->ep.n_jobs < 1) return set_error(RAMP_ERR_BAD_ARG, "ramp_reset must be called before ramp_step");
    if (e->templates.empty()) {
        // allowed: every action must then be Action(); still need a scratch-less plan
    }
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    if (e->n_nonresident > 0) { int rc = ensure_scratch(e); if (rc != RAMP_OK) return rc; }
    if (e->n_resident > 0) { int rc = ensure_thread_scratch(e); if (rc != RAMP_OK) return rc; }
    const int B = e->cfg.n_episodes;
    cudaStream_t st = e->stream;
    CUDA_TRY(cudaMemsetAsync(e->d_counters, 0, 4 * sizeof(int32_t), st));   // both work lists' counts and cursors
    PlanArgs p{};
    p.actions = d_actions; p.templates = e->d_templates; p.n_templates = (int32_t)e->templates.size();
    p.ep = e->ep; p.memo.keys = e->d_memo_keys; p.memo.mask = e->memo_cap - 1; p.memo.mode = e->cfg.memo_mode;
    p.memo.vals = e->d_memo_vals; p.memo.keys2 = e->d_memo_keys2; p.memo.mask2 = e->memo_cap2 - 1;
    p.memo.slot2_base = (int32_t)e->memo_cap + e->cfg.n_episodes;
    p.items = e->d_items; p.items_big = e->d_items_big; p.items_res = e->d_items_res; p.counters = e->d_counters; p.stats = e->d_stats;
    ramp_plan_kernel<<<(B + 127) / 128, 128, 0, st>>>(p);
    e->launches++;
    if (!e->templates.empty()) {
        if (e->ev_pending >= MAX_EVENT_PAIRS) { CUDA_TRY(cudaStreamSynchronize(st)); int rc = resolve_events(e); if (rc) return rc; }
        CUDA_TRY(cudaEventRecord(e->ev_a[e->ev_pending], st));
        if (e->n_resident > 0) {
            // memo misses on resident templates: grouped by template into chunks of <= 32, one THREAD per lookahead.  No
            // host read-back: the counts stay on the device, idle CTAs find the chunk cursor exhausted and exit.
            BucketArgs ba{};
            ba.items = e->d_items_res; ba.n_items = &e->d_counters->n_work_res; ba.n_templates = (int32_t)e->templates.size();
            ba.tcount = e->d_tcount; ba.tbase = e->d_tbase; ba.chunk_items = e->d_chunk_items; ba.chunks = e->d_chunks;
            ba.n_chunks = &e->d_counters->n_chunks; ba.cursor = &e->d_counters->chunk_cursor; ba.rank = e->d_rank;
            ramp_bucket_kernel<<<1, 1024, 0, st>>>(ba);
            ThreadArgs ta = make_thread_args(e, e->d_chunks, &e->d_counters->n_chunks, &e->d_counters->chunk_cursor, e->d_chunk_items,
                                             e->res, e->pool, e->d_stats);
            ramp_lookahead_thread_kernel<<<e->res_grid, 32, e->res_smem, st>>>(ta);
            e->launches += 2;
        }
        if (e->n_nonresident > 0) {
        // the number of memo misses of each size class decides the kernel shapes: a 16-byte read-back (~10 us) against
        // multi-ms kernels
        CUDA_TRY(cudaMemcpyAsync(e->h_n_work, &e->d_counters->n_work, sizeof(int32_t) * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        const int n_small = e->h_n_work[0], n_big = e->h_n_work[2];
        if (n_small + n_big > 0) {
            LookaheadArgs a = make_lookahead_args(e, e->d_items, e->d_counters, e->res, true, e->d_stats);
            LookaheadArgs ab = a;
            ab.items = e->d_items_big; ab.n_work = &e->d_counters->n_work_big; ab.cursor = &e->d_counters->work_cursor_big;
            const int wpb = e->nt / 32;
            const int warp_slots = e->grid * wpb;
            // The big lookaheads set the step's latency, the small ones its load.  While everything fits the SMs' warp slots
            // at once (registers cap every mix at cta_grid x 4 warps) each big lookahead gets a 256-thread CTA if 8 warps per
            // big one still fit (2.4 ms instead of 2.9 ms for the bench job), else a 128-thread CTA; while the big
            // ones still fit as 64-thread CTAs and the small ones need at most a short second wave they get those; beyond
            // that everything goes through the warp kernel, the big list first (longest-processing-time-first keeps the
            // tail short).  The CTA kernel runs on a second stream beside the warp kernel for the small list.
            const int reg_slots = e->cta_grid * 4;
            int split_nt = 0;
            if (e->mode != 1 && n_big > 0) {
                if (e->use_cta256 && n_big <= e->cta256_grid && 8 * n_big + n_small <= reg_slots) split_nt = 256;
                else if (n_big <= e->cta_grid && 4 * n_big + n_small <= reg_slots) split_nt = 128;
                else if (n_big <= e->cta64_grid && 2 * n_big + n_small <= (int)(e->split_alpha * reg_slots)) split_nt = 64;
            }
            const bool split = split_nt != 0;
            if (e->debug) fprintf(stderr, "[ramp] step lookaheads: small=%d big=%d warp_slots=%d reg_slots=%d -> %s %d\n", n_small, n_big,
                                  warp_slots, reg_slots, split ? "split (CTA || warp), CTA threads" : "single warp kernel, big first", split_nt);
            if (split) {
                const int cgrid = e->cta_grid_for(split_nt);
                const int grid = std::min(n_big, cgrid);
                // `st` is idle here (synchronised for the read-back above), so the second stream needs no fork event and the
                // CTA kernel's blocks are always placed before the warp kernel's
                lookahead_cta_kernel_for(split_nt)<<<grid, split_nt, e->cta_smem_bytes, e->stream2>>>(ab);
                CUDA_TRY(cudaEventRecord(e->ev_join, e->stream2));
                e->launches++;
                if (n_small > 0) {
                    LookaheadArgs as = a;
                    as.scratch = a.scratch + (uint64_t)std::max(e->cta_grid, e->cta64_grid) * a.scratch_stride;   // slabs past the CTA kernel's
                    const int wpb2 = e->split_warp_nt / 32;
                    const int g2 = std::max(1, std::min(e->grid * wpb / wpb2, (n_small + wpb2 - 1) / wpb2));   // never more warps than slabs
                    lookahead_kernel_for(e->split_warp_nt)<<<g2, e->split_warp_nt, e->smem2_bytes, st>>>(as);
                    e->launches++;
                }
                CUDA_TRY(cudaStreamWaitEvent(st, e->ev_join, 0));
            } else if (e->mode == 2) {
                if (n_big > 0) { launch_lookahead(e, ab, n_big, n_big, st); e->launches++; }
                if (n_small > 0) { launch_lookahead(e, a, n_small, 0, st); e->launches++; }
            } else {
                LookaheadArgs all = ab;                      // list A = big, list B = small, one cursor
                all.items_b = e->d_items; all.n_work_b = &e->d_counters->n_work;
                const int n_all = n_small + n_big;
                if (e->mode == 0 && e->dense_grid * (DENSE_NT / 32) > warp_slots && n_all > (int)(e->dense_factor * warp_slots)) {
                    // far more lookaheads than slots: throughput matters, not the latency of one -> 16 warps per SM
                    const int wd = DENSE_NT / 32;
                    const int g = std::max(1, std::min(e->dense_grid, (n_all + wd - 1) / wd));
                    lookahead_dense_kernel()<<<g, DENSE_NT, e->dense_smem_bytes, st>>>(all);
                } else {
                    const int g = std::max(1, std::min(e->grid, (n_all + wpb - 1) / wpb));
                    lookahead_kernel_for(e->nt)<<<g, e->nt, e->smem_bytes, st>>>(all);
                }
                e->launches++;
            }
        }
        }
        CUDA_TRY(cudaEventRecord(e->ev_b[e->ev_pending], st));
        e->ev_pending++;
    }
    StepArgs s{};
    s.actions = d_actions; s.ep = e->ep; s.res = e->res; s.pool = e->pool; s.counters = e->d_counters;
    s.stats_out = d_stats_out; s.n_cluster_steps_out = d_ncs_out; s.fuse_empty_steps = fuse;
    ramp_step_kernel<<<(B + 63) / 64, 64, 0, st>>>(s);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return RAMP_OK;
}

int ramp_sync(ramp_engine_t* e) {
    if (!e) return set_error(RAMP_ERR_BAD_ARG, "null engine");
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return resolve_events(e);
}

int ramp_step_host(ramp_engine_t* e, const ramp_action_t* actions, int32_t fuse, double* stats_out, int32_t* ncs_out) {
    if (!e || !actions) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    const int B = e->cfg.n_episodes;
    CUDA_TRY(cudaMemcpyAsync(e->d_actions, actions, sizeof(ramp_action_t) * B, cudaMemcpyHostToDevice, e->stream));
    int rc = ramp_step_device(e, e->d_actions, fuse, stats_out ? e->d_step_stats : nullptr, ncs_out ? e->d_n_cluster_steps : nullptr);
    if (rc != RAMP_OK) return rc;
    if (stats_out)
        CUDA_TRY(cudaMemcpyAsync(stats_out, e->d_step_stats, sizeof(double) * RAMP_STEP_STATS_LEN * B, cudaMemcpyDeviceToHost, e->stream));
    if (ncs_out)
        CUDA_TRY(cudaMemcpyAsync(ncs_out, e->d_n_cluster_steps, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, e->stream));
    return ramp_sync(e);
}

int ramp_check_status(ramp_engine_t* e, int32_t* ep_out, int32_t* st_out) {
    if (!e) return set_error(RAMP_ERR_BAD_ARG, "null engine");
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    Counters c{};
    CUDA_TRY(cudaMemcpy(&c, e->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost));
    if (c.err_episode == 0) { if (ep_out) *ep_out = -1; if (st_out) *st_out = 0; return RAMP_OK; }
    const int b = c.err_episode - 1;
    int32_t st = 0;
    CUDA_TRY(cudaMemcpy(&st, e->ep.ei + (size_t)EI_STATUS * e->cfg.n_episodes + b, sizeof(int32_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemset(&e->d_counters->err_episode, 0, sizeof(int32_t)));
    if (ep_out) *ep_out = b;
    if (st_out) *st_out = st;
    const char* what = st == RAMP_ST_INFINITE_TICK ? "ERROR: Last tick was infinite, a bug has occurred somewhere."
                     : st == RAMP_ST_TRACE_OVERFLOW ? "lookahead needed more ticks than trace_cap (or the trace pool is full)"
                     : st == RAMP_ST_TABLE_FULL ? "running-job table or memo table is full"
                     : st == RAMP_ST_NO_QUEUED_JOB ? "an action was given for an episode whose job queue is empty"
                     : st == RAMP_ST_BAD_TEMPLATE ? "an action names a template id that was never registered"
                     : "simulation error";
    return set_error(RAMP_ERR_SIM, "episode %d: %s (status %d)", b, what, st);
}

int ramp_get_job_records(ramp_engine_t* e, ramp_job_record_t* out) {
    if (!e || !out) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    CUDA_TRY(cudaMemcpy(out, e->ep.rec, sizeof(ramp_job_record_t) * (size_t)e->cfg.max_jobs * e->cfg.n_episodes, cudaMemcpyDeviceToHost));
    return RAMP_OK;
}

int ramp_episode_state_device(ramp_engine_t* e, double** d_out) {
    if (!e || !d_out) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    const int B = e->cfg.n_episodes;
    ramp_export_episode_state_kernel<<<(B + 127) / 128, 128, 0, e->stream>>>(e->ep, e->d_ep_export);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    *d_out = e->d_ep_export;
    return RAMP_OK;
}

int ramp_export_episode_state_to(ramp_engine_t* e, double* d_dst) {
    if (!e || !d_dst) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    const int B = e->cfg.n_episodes;
    ramp_export_episode_state_kernel<<<(B + 127) / 128, 128, 0, e->stream>>>(e->ep, d_dst);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return RAMP_OK;
}

int ramp_get_episode_state(ramp_engine_t* e, double* out) {
    double* d = nullptr;
    int rc = ramp_episode_state_device(e, &d);
    if (rc != RAMP_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, d, sizeof(double) * RAMP_EP_LEN * e->cfg.n_episodes, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return RAMP_OK;
}

int ramp_get_memo_stats(ramp_engine_t* e, int64_t* lookups, int64_t* hits, int64_t* lookaheads) {
    if (!e) return set_error(RAMP_ERR_BAD_ARG, "null engine");
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    MemoStats s{};
    CUDA_TRY(cudaMemcpy(&s, e->d_stats, sizeof(MemoStats), cudaMemcpyDeviceToHost));
    if (lookups) *lookups = (int64_t)(s.lookups - e->memo_base.lookups);
    if (hits) *hits = (int64_t)(s.hits - e->memo_base.hits);
    if (lookaheads) *lookaheads = (int64_t)(s.lookaheads - e->memo_base.lookaheads);
    return RAMP_OK;
}

int ramp_get_memo_stats_ex(ramp_engine_t* e, int64_t out[4]) {
    if (!e || !out) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    MemoStats s{};
    CUDA_TRY(cudaMemcpy(&s, e->d_stats, sizeof(MemoStats), cudaMemcpyDeviceToHost));
    out[0] = (int64_t)(s.lookups - e->memo_base.lookups);
    out[1] = (int64_t)(s.hits - e->memo_base.hits);
    out[2] = (int64_t)(s.shared_hits - e->memo_base.shared_hits);
    out[3] = (int64_t)(s.lookaheads - e->memo_base.lookaheads);
    return RAMP_OK;
}

int ramp_get_last_lookahead(ramp_engine_t* e, int32_t episode, ramp_lookahead_result_t* res, int32_t* tn, double* tt, int32_t cap) {
    if (!e || !res) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    if (episode < 0 || episode >= e->cfg.n_episodes) return set_error(RAMP_ERR_BAD_ARG, "episode out of range");
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    int32_t slot = -1;
    CUDA_TRY(cudaMemcpy(&slot, e->ep.ei + (size_t)EI_LAST_SLOT * e->cfg.n_episodes + episode, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (slot < 0) return set_error(RAMP_ERR_BAD_ARG, "episode %d has not mounted a job yet", episode);
    int64_t off = -1;
    CUDA_TRY(cudaMemcpy(&res->jct, e->res.jct + slot, sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&res->comm, e->res.comm + slot, sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&res->comp, e->res.comp + slot, sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&res->n_ticks, e->res.n_ticks + slot, sizeof(int32_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&res->status, e->res.status + slot, sizeof(int32_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&off, e->res.trace_off + slot, sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (tn && tt && off >= 0) {
        const int32_t n = std::min(std::min(res->n_ticks, cap), e->cfg.trace_cap);
        CUDA_TRY(cudaMemcpy(tn, e->pool.n_active + off, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(tt, e->pool.tick + off, sizeof(double) * n, cudaMemcpyDeviceToHost));
    }
    return RAMP_OK;
}

int ramp_run_lookaheads(ramp_engine_t* e, const int32_t* template_ids, int32_t n, ramp_lookahead_result_t* results,
                        int32_t* trace_n, double* trace_tick, int32_t trace_cap, float* kernel_ms_out) {
    if (!e || !template_ids || !results || n < 0) return set_error(RAMP_ERR_BAD_ARG, "bad argument");
    if (n == 0) return RAMP_OK;
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    for (int32_t k = 0; k < n; ++k)
        if (template_ids[k] < 0 || template_ids[k] >= (int32_t)e->templates.size())
            return set_error(RAMP_ERR_BAD_ARG, "template id %d at %d is not registered", template_ids[k], k);
    std::vector<int32_t> res_idx, old_idx;
    for (int32_t k = 0; k < n; ++k) (e->templates[template_ids[k]].dev.size_class == 2 ? res_idx : old_idx).push_back(k);
    int rc = RAMP_OK;
    if (!old_idx.empty()) { rc = ensure_scratch(e); if (rc != RAMP_OK) return rc; }
    if (!res_idx.empty()) { rc = ensure_thread_scratch(e); if (rc != RAMP_OK) return rc; }
    cudaStream_t st = e->stream;
    if (n > e->sa_cap) {
        CUDA_TRY(cudaStreamSynchronize(st));
        free_result_slots(e->sa_res); cudaFree(e->sa_items); cudaFree(e->sa_chunk_items); cudaFree(e->sa_chunks);
        if (alloc_result_slots(e->sa_res, n) != RAMP_OK) return RAMP_ERR_CUDA;
        CUDA_TRY(cudaMalloc(&e->sa_items, sizeof(WorkItem) * n));
        CUDA_TRY(cudaMalloc(&e->sa_chunk_items, sizeof(WorkItem) * (size_t)n * 32));
        CUDA_TRY(cudaMalloc(&e->sa_chunks, sizeof(ChunkDesc) * n));
        e->sa_cap = n;
    }
    if (!e->sa_counters) CUDA_TRY(cudaMalloc(&e->sa_counters, sizeof(Counters)));
    std::vector<WorkItem> items(old_idx.size());
    for (size_t q = 0; q < old_idx.size(); ++q) {
        const int32_t k = old_idx[q];
        items[q].template_id = template_ids[k]; items[q].slot = k; items[q].episode = -1; items[q].n_mounted_workers = 0;
    }
    // resident templates: chunks of <= 32 items of one template, built here (the step path builds them on the device)
    std::vector<ChunkDesc> chunks;
    std::vector<WorkItem> chunk_items;
    {
        std::vector<int32_t> order(res_idx);
        std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return template_ids[x] < template_ids[y]; });
        for (size_t q = 0; q < order.size();) {
            const int32_t t = template_ids[order[q]];
            size_t r = q;
            while (r < order.size() && template_ids[order[r]] == t && r - q < 32) ++r;
            ChunkDesc cd; cd.template_id = t; cd.count = (int32_t)(r - q);
            chunks.push_back(cd);
            const size_t base = chunk_items.size();
            chunk_items.resize(base + 32);
            for (size_t x = q; x < r; ++x) {
                WorkItem& it = chunk_items[base + (x - q)];
                it.template_id = t; it.slot = order[x]; it.episode = -1; it.n_mounted_workers = 0;
            }
            q = r;
        }
    }
    Counters c{}; c.n_work = (int32_t)old_idx.size(); c.n_chunks = (int32_t)chunks.size();
    if (!items.empty()) CUDA_TRY(cudaMemcpyAsync(e->sa_items, items.data(), sizeof(WorkItem) * items.size(), cudaMemcpyHostToDevice, st));
    if (!chunks.empty()) {
        CUDA_TRY(cudaMemcpyAsync(e->sa_chunks, chunks.data(), sizeof(ChunkDesc) * chunks.size(), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->sa_chunk_items, chunk_items.data(), sizeof(WorkItem) * chunk_items.size(), cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(e->sa_counters, &c, sizeof(Counters), cudaMemcpyHostToDevice, st));
    // traces of standalone runs go to a private pool sized n x trace_cap when requested
    TracePool priv{};
    const bool want_trace = trace_n && trace_tick && trace_cap > 0;
    unsigned long long* d_top = nullptr;
    if (want_trace) {
        priv.len = (uint64_t)n * (uint64_t)e->cfg.trace_cap;      // the kernels allocate n_ticks entries per lookahead; truncated on copy-out
        CUDA_TRY(cudaMalloc(&priv.n_active, sizeof(int32_t) * priv.len));
        CUDA_TRY(cudaMalloc(&priv.tick, sizeof(double) * priv.len));
        CUDA_TRY(cudaMalloc(&d_top, sizeof(unsigned long long)));
        CUDA_TRY(cudaMemsetAsync(d_top, 0, sizeof(unsigned long long), st));
        priv.top = d_top;
    }
    LookaheadArgs a = make_lookahead_args(e, e->sa_items, e->sa_counters, e->sa_res, false, nullptr);
    if (want_trace) a.pool = priv;
    cudaEvent_t ea = e->ev_a[MAX_EVENT_PAIRS - 1], eb = e->ev_b[MAX_EVENT_PAIRS - 1];
    if (e->ev_pending >= MAX_EVENT_PAIRS - 1) { CUDA_TRY(cudaStreamSynchronize(st)); rc = resolve_events(e); if (rc) return rc; }
    CUDA_TRY(cudaEventRecord(ea, st));
    if (!chunks.empty()) {
        TracePool tp = want_trace ? priv : e->pool;
        if (!want_trace) tp.top = nullptr;
        ThreadArgs ta = make_thread_args(e, e->sa_chunks, &e->sa_counters->n_chunks, &e->sa_counters->chunk_cursor, e->sa_chunk_items,
                                         e->sa_res, tp, nullptr);
        const int g = std::max(1, std::min(e->res_grid, (int)chunks.size()));
        ramp_lookahead_thread_kernel<<<g, 32, e->res_smem, st>>>(ta);
        e->launches++;
    }
    if (!old_idx.empty()) {
        int n_big = 0;
        for (int32_t k : old_idx) n_big += e->templates[template_ids[k]].dev.size_class == 1 ? 1 : 0;
        launch_lookahead(e, a, (int)old_idx.size(), n_big, st);
        e->launches++;
    }
    CUDA_TRY(cudaEventRecord(eb, st));
    CUDA_TRY(cudaGetLastError());
    std::vector<double> jct(n), comm(n), comp(n);
    std::vector<int32_t> nt(n), stt(n);
    std::vector<int64_t> off(n);
    CUDA_TRY(cudaMemcpyAsync(jct.data(), e->sa_res.jct, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(comm.data(), e->sa_res.comm, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(comp.data(), e->sa_res.comp, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(nt.data(), e->sa_res.n_ticks, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(stt.data(), e->sa_res.status, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(off.data(), e->sa_res.trace_off, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (kernel_ms_out) CUDA_TRY(cudaEventElapsedTime(kernel_ms_out, ea, eb));
    for (int32_t k = 0; k < n; ++k) {
        results[k].jct = jct[k]; results[k].comm = comm[k]; results[k].comp = comp[k];
        results[k].n_ticks = nt[k]; results[k].status = stt[k];
    }
    if (want_trace) {
        for (int32_t k = 0; k < n; ++k) {
            if (off[k] < 0) continue;
            const int32_t m = std::min(std::min(nt[k], trace_cap), e->cfg.trace_cap);
            CUDA_TRY(cudaMemcpy(trace_n + (size_t)k * trace_cap, priv.n_active + off[k], sizeof(int32_t) * m, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(trace_tick + (size_t)k * trace_cap, priv.tick + off[k], sizeof(double) * m, cudaMemcpyDeviceToHost));
        }
        cudaFree(priv.n_active); cudaFree(priv.tick); cudaFree(d_top);
    }
    return RAMP_OK;
}

int64_t ramp_launch_count(ramp_engine_t* e) { return e ? e->launches : 0; }

int ramp_get_lookahead_kernel_time(ramp_engine_t* e, double* total_ms, int64_t* launches, int64_t* work_items,
                                   int64_t* alg_bytes, int32_t reset) {
    if (!e) return set_error(RAMP_ERR_BAD_ARG, "null engine");
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    int rc = resolve_events(e);
    if (rc) return rc;
    MemoStats s{};
    CUDA_TRY(cudaMemcpy(&s, e->d_stats, sizeof(MemoStats), cudaMemcpyDeviceToHost));
    if (total_ms) *total_ms = e->la_ms_total;
    if (launches) *launches = e->la_launches;
    if (work_items) *work_items = (int64_t)(s.lookaheads - e->la_items_base);
    if (alg_bytes) *alg_bytes = (int64_t)(s.alg_bytes - e->la_bytes_base);
    if (reset) { e->la_ms_total = 0.0; e->la_launches = 0; e->la_items_base = s.lookaheads; e->la_bytes_base = s.alg_bytes; e->la_qbytes_base = s.quotient_bytes; }
    return RAMP_OK;
}


int ramp_get_quotient_bytes(ramp_engine_t* e, int64_t* quotient_bytes) {
    if (!e || !quotient_bytes) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    MemoStats s{};
    CUDA_TRY(cudaMemcpy(&s, e->d_stats, sizeof(MemoStats), cudaMemcpyDeviceToHost));
    *quotient_bytes = (int64_t)(s.quotient_bytes - e->la_qbytes_base);
    return RAMP_OK;
}


// ---- device-resident rollouts ---------------------------------------------------------------------------------------
extern "C++" {
namespace {
template <class T> int env_upload(ramp_engine* e, T** dst, const T* src, size_t n) {
    CUDA_TRY(cudaMalloc(dst, sizeof(T) * std::max<size_t>(n, 1)));
    e->env_allocs.push_back(*dst);
    if (src && n) CUDA_TRY(cudaMemcpy(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice));
    else CUDA_TRY(cudaMemset(*dst, 0, sizeof(T) * std::max<size_t>(n, 1)));
    return RAMP_OK;
}
}  // namespace
}  // extern "C++"

int ramp_env_create(ramp_engine_t* e, const ramp_env_config_t* c) {
    if (!e || !c) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    if (e->has_env) return set_error(RAMP_ERR_BAD_ARG, "the engine already has an environment");
    const int B = e->cfg.n_episodes, J = c->jobs_per_episode, M = c->n_models, D = c->max_degree, G = c->n_geoms, nw = c->n_words;
    const int n_workers = c->shape[0] * c->shape[1] * c->shape[2];
    if (J != e->cfg.max_jobs) return set_error(RAMP_ERR_BAD_ARG, "jobs_per_episode %d != the engine's max_jobs %d", J, e->cfg.max_jobs);
    if (n_workers != e->cfg.n_cluster_workers || nw * 64 < n_workers || M < 1 || D < 1 || G < 1)
        return set_error(RAMP_ERR_BAD_ARG, "bad environment shape");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    EnvDev& v = e->env;
    v.B = B; v.J = J; v.n_words = nw; v.n_models = M; v.max_degree = D; v.n_geoms = G; v.n_workers = n_workers;
    v.apply_mask = c->apply_action_mask; v.fail_reward = c->fail_reward; v.success_reward = c->success_reward;
    v.num_training_steps = (double)c->num_training_steps;
    const int n_cand = c->cand_ptr[D + 1];
    int rc;
    int32_t* cand_ptr; unsigned long long* cand_mask; int32_t* cand_geom; uint8_t* uniform; uint8_t* shape_ok; double* mp; double* jp;
    if ((rc = env_upload(e, &cand_ptr, c->cand_ptr, (size_t)D + 2))) return rc;
    if ((rc = env_upload(e, &cand_mask, (const unsigned long long*)c->cand_mask, (size_t)n_cand * nw))) return rc;
    if ((rc = env_upload(e, &cand_geom, c->cand_geom, (size_t)n_cand))) return rc;
    if ((rc = env_upload(e, &uniform, c->uniform, (size_t)M * (D + 1)))) return rc;
    if ((rc = env_upload(e, &shape_ok, c->shape_ok, (size_t)D + 1))) return rc;
    if ((rc = env_upload(e, &mp, c->model_params, (size_t)M * 5))) return rc;
    if ((rc = env_upload(e, &jp, c->jobs_params, (size_t)16))) return rc;
    v.cand_ptr = cand_ptr; v.cand_mask = cand_mask; v.cand_geom = cand_geom; v.uniform = uniform; v.shape_ok = shape_ok;
    v.model_params = mp; v.jobs_params = jp;
    std::vector<int32_t> minus1((size_t)M * (D + 1) * G, -1);
    if ((rc = env_upload(e, &v.tmpl_of, minus1.data(), minus1.size()))) return rc;
    if ((rc = env_upload<double>(e, &v.tmpl_mount, nullptr, (size_t)e->cfg.max_templates * 6))) return rc;
    int32_t* model_of; double* frac; double* macc;
    if ((rc = env_upload<int32_t>(e, &model_of, nullptr, (size_t)B * J))) return rc;
    if ((rc = env_upload<double>(e, &frac, nullptr, (size_t)B * J))) return rc;
    if ((rc = env_upload<double>(e, &macc, nullptr, (size_t)B * J))) return rc;
    v.model_of = model_of; v.frac = frac; v.macc = macc;
    if ((rc = env_upload<unsigned long long>(e, &v.busy, nullptr, (size_t)B * nw))) return rc;
    if ((rc = env_upload<unsigned long long>(e, &v.job_mask, nullptr, (size_t)B * J * nw))) return rc;
    if ((rc = env_upload<unsigned long long>(e, &v.placed, nullptr, (size_t)B * nw))) return rc;
    if ((rc = env_upload<int32_t>(e, &v.tid, nullptr, (size_t)B))) return rc;
    if ((rc = env_upload<int32_t>(e, &v.decided_job, nullptr, (size_t)B))) return rc;
    if ((rc = env_upload<int32_t>(e, &v.n_decided, nullptr, (size_t)B))) return rc;
    if ((rc = env_upload<int32_t>(e, &v.actions, nullptr, (size_t)B))) return rc;
    if ((rc = env_upload<double>(e, &v.reward, nullptr, (size_t)B))) return rc;
    if ((rc = env_upload<uint8_t>(e, &v.done, nullptr, (size_t)B))) return rc;
    if ((rc = env_upload<int32_t>(e, &v.queued_model, nullptr, (size_t)B))) return rc;
    if ((rc = env_upload<float>(e, &v.obs_dyn, nullptr, (size_t)B * 11))) return rc;
    if ((rc = env_upload<uint8_t>(e, &v.action_mask, nullptr, (size_t)B * (D + 1)))) return rc;
    if ((rc = env_upload<int32_t>(e, &v.need_host, nullptr, (size_t)B))) return rc;
    if ((rc = env_upload<int32_t>(e, &v.n_need_host, nullptr, 1))) return rc;
    if ((rc = env_upload<int32_t>(e, &v.err, nullptr, 1))) return rc;
    CUDA_TRY(cudaMallocHost(&e->env_h_need, sizeof(int32_t) * 8));
    e->has_env = true;
    return RAMP_OK;
}

int ramp_env_set_template(ramp_engine_t* e, int32_t model, int32_t degree, int32_t geom, int32_t template_id, const double mount[6]) {
    if (!e || !e->has_env || !mount) return set_error(RAMP_ERR_BAD_ARG, "no environment");
    const EnvDev& v = e->env;
    if (model < 0 || model >= v.n_models || degree < 0 || degree > v.max_degree || geom < 0 || geom >= v.n_geoms ||
        template_id < 0 || template_id >= (int32_t)e->templates.size())
        return set_error(RAMP_ERR_BAD_ARG, "bad template table entry (model %d degree %d geometry %d template %d)", model, degree, geom, template_id);
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaMemcpy(v.tmpl_of + ((size_t)model * (v.max_degree + 1) + degree) * v.n_geoms + geom, &template_id, sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(v.tmpl_mount + (size_t)template_id * 6, mount, sizeof(double) * 6, cudaMemcpyHostToDevice));
    return RAMP_OK;
}

int ramp_env_reset(ramp_engine_t* e, const int32_t* model_of, const double* frac, const double* macc, const ramp_arrival_t* arrivals) {
    if (!e || !e->has_env || !model_of || !frac || !macc || !arrivals) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    EnvDev& v = e->env;
    const size_t n = (size_t)v.B * v.J;
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaMemcpyAsync((void*)v.model_of, model_of, sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync((void*)v.frac, frac, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync((void*)v.macc, macc, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemsetAsync(v.job_mask, 0, sizeof(unsigned long long) * n * v.n_words, e->stream));
    CUDA_TRY(cudaMemsetAsync(v.done, 0, (size_t)v.B, e->stream));
    CUDA_TRY(cudaMemsetAsync(v.n_decided, 0, sizeof(int32_t) * (size_t)v.B, e->stream));
    CUDA_TRY(cudaMemsetAsync(v.err, 0, sizeof(int32_t), e->stream));
    int rc = ramp_reset(e, arrivals, v.J);
    if (rc != RAMP_OK) return rc;
    ramp_env_update_kernel<<<(v.B + 127) / 128, 128, 0, e->stream>>>(v, e->ep, nullptr, 1);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return RAMP_OK;
}

int ramp_env_buffers(ramp_engine_t* e, ramp_env_buffers_t* out) {
    if (!e || !e->has_env || !out) return set_error(RAMP_ERR_BAD_ARG, "no environment");
    const EnvDev& v = e->env;
    out->actions = v.actions; out->reward = v.reward; out->done = v.done; out->queued_model = v.queued_model;
    out->obs_dynamic = v.obs_dyn; out->action_mask = v.action_mask; out->busy = (uint64_t*)v.busy; out->template_id = v.tid;
    out->n_episodes = v.B; out->n_actions = v.max_degree + 1; out->n_models = v.n_models;
    return RAMP_OK;
}

int ramp_env_host_mirror(ramp_engine_t* e, ramp_env_buffers_t* out) {
    if (!e || !e->has_env || !out) return set_error(RAMP_ERR_BAD_ARG, "no environment");
    const EnvDev& v = e->env;
    const size_t B = (size_t)v.B, A = (size_t)v.max_degree + 1;
    const size_t o_reward = 0, o_obs = o_reward + 8 * B, o_act = o_obs + 44 * B, o_qm = o_act + 4 * B, o_done = o_qm + 4 * B, o_mask = o_done + B;
    if (!e->env_h_mirror) {
        CUDA_TRY(cudaSetDevice(e->cfg.device));
        CUDA_TRY(cudaMallocHost(&e->env_h_mirror, o_mask + B * A + 64));
        memset(e->env_h_mirror, 0, o_mask + B * A + 64);
    }
    unsigned char* m = (unsigned char*)e->env_h_mirror;
    memset(out, 0, sizeof(*out));
    out->reward = (double*)(m + o_reward); out->obs_dynamic = (float*)(m + o_obs); out->actions = (int32_t*)(m + o_act);
    out->queued_model = (int32_t*)(m + o_qm); out->done = m + o_done; out->action_mask = m + o_mask;
    out->n_episodes = v.B; out->n_actions = v.max_degree + 1; out->n_models = v.n_models;
    return RAMP_OK;
}

int ramp_env_decide(ramp_engine_t* e, const int32_t* actions, int32_t* n_need_host_out, int32_t* need_host_out) {
    if (!e || !e->has_env) return set_error(RAMP_ERR_BAD_ARG, "no environment");
    EnvDev& v = e->env;
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    cudaStream_t st = e->stream;
    if (actions) CUDA_TRY(cudaMemcpyAsync(v.actions, actions, sizeof(int32_t) * v.B, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(v.n_need_host, 0, sizeof(int32_t), st));
    ramp_env_decide_kernel<<<(v.B + 127) / 128, 128, 0, st>>>(v, e->ep, e->d_actions);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    if (n_need_host_out) {
        CUDA_TRY(cudaMemcpyAsync(e->env_h_need, v.n_need_host, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(e->env_h_need + 1, v.err, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (e->env_h_need[1] != 0) {
            CUDA_TRY(cudaMemset(v.err, 0, sizeof(int32_t)));
            return set_error(RAMP_ERR_BAD_ARG, "episode %d: the action is invalid given its action mask (RJPE:314-319)", e->env_h_need[1] - 1);
        }
        *n_need_host_out = e->env_h_need[0];
        if (need_host_out && e->env_h_need[0] > 0)
            CUDA_TRY(cudaMemcpy(need_host_out, v.need_host, sizeof(int32_t) * e->env_h_need[0], cudaMemcpyDeviceToHost));
    } else {
        e->env_unchecked_decide = true;      // looked at by the next ramp_env_read
    }
    return RAMP_OK;
}

int ramp_env_patch(ramp_engine_t* e, int32_t episode, int32_t template_id, const uint64_t* server_mask, const double mount[6]) {
    if (!e || !e->has_env || !server_mask || !mount) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    EnvDev& v = e->env;
    if (episode < 0 || episode >= v.B || template_id >= (int32_t)e->templates.size()) return set_error(RAMP_ERR_BAD_ARG, "bad patch");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    int32_t q = -1;
    CUDA_TRY(cudaMemcpy(&q, v.decided_job + episode, sizeof(int32_t), cudaMemcpyDeviceToHost));
    ramp_action_t row{};
    row.template_id = template_id;
    if (template_id >= 0 && q >= 0) {
        double fr = 0.0, ov = 0.0;
        CUDA_TRY(cudaMemcpy(&fr, v.frac + (size_t)episode * v.J + q, sizeof(double), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(&ov, v.macc + (size_t)episode * v.J + q, sizeof(double), cudaMemcpyDeviceToHost));
        row.max_acceptable_jct = std::isnan(ov) ? fr * mount[0] : ov;
        row.part_op_mem = mount[1]; row.part_dep_size = mount[2]; row.flow_size = mount[3];
        row.n_mounted_workers = (int32_t)mount[4]; row.n_mounted_channels = (int32_t)mount[5];
    } else {
        row.template_id = -1;
    }
    CUDA_TRY(cudaMemcpy(e->d_actions + episode, &row, sizeof(row), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(v.tid + episode, &row.template_id, sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(v.placed + (size_t)episode * v.n_words, server_mask, sizeof(uint64_t) * v.n_words, cudaMemcpyHostToDevice));
    return RAMP_OK;
}

int ramp_env_advance(ramp_engine_t* e) {
    if (!e || !e->has_env) return set_error(RAMP_ERR_BAD_ARG, "no environment");
    EnvDev& v = e->env;
    int rc = ramp_step_device(e, e->d_actions, 1, e->d_step_stats, e->d_n_cluster_steps);
    if (rc != RAMP_OK) return rc;
    ramp_env_update_kernel<<<(v.B + 127) / 128, 128, 0, e->stream>>>(v, e->ep, e->d_n_cluster_steps, 0);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return RAMP_OK;
}

int ramp_env_read(ramp_engine_t* e, double* reward, uint8_t* done, int32_t* queued_model, float* obs_dynamic, uint8_t* action_mask) {
    if (!e || !e->has_env) return set_error(RAMP_ERR_BAD_ARG, "no environment");
    const EnvDev& v = e->env;
    cudaStream_t st = e->stream;
    if (reward) CUDA_TRY(cudaMemcpyAsync(reward, v.reward, sizeof(double) * v.B, cudaMemcpyDeviceToHost, st));
    if (done) CUDA_TRY(cudaMemcpyAsync(done, v.done, (size_t)v.B, cudaMemcpyDeviceToHost, st));
    if (queued_model) CUDA_TRY(cudaMemcpyAsync(queued_model, v.queued_model, sizeof(int32_t) * v.B, cudaMemcpyDeviceToHost, st));
    if (obs_dynamic) CUDA_TRY(cudaMemcpyAsync(obs_dynamic, v.obs_dyn, sizeof(float) * 11 * v.B, cudaMemcpyDeviceToHost, st));
    if (action_mask) CUDA_TRY(cudaMemcpyAsync(action_mask, v.action_mask, (size_t)v.B * (v.max_degree + 1), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(e->env_h_need + 4, &e->d_counters->err_episode, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (e->env_unchecked_decide) {
        // decisions taken without the host looking (a device-resident policy): an invalid action under apply_action_mask is sticky
        // in `err`; episodes the tables could not decide were left unplaced, which only the LAST decide's count can show
        CUDA_TRY(cudaMemcpyAsync(e->env_h_need + 2, v.n_need_host, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(e->env_h_need + 3, v.err, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        int rc = ramp_sync(e);
        if (rc != RAMP_OK) return rc;
        e->env_unchecked_decide = false;
        if (e->env_h_need[3] != 0) {
            CUDA_TRY(cudaMemset(v.err, 0, sizeof(int32_t)));
            return set_error(RAMP_ERR_BAD_ARG, "episode %d: the action is invalid given its action mask (RJPE:314-319)", e->env_h_need[3] - 1);
        }
        if (e->env_h_need[2] != 0)
            return set_error(RAMP_ERR_BAD_ARG, "%d episodes needed the host's placer but ramp_env_decide was called without need_host_out", e->env_h_need[2]);
        return e->env_h_need[4] != 0 ? ramp_check_status(e, nullptr, nullptr) : RAMP_OK;
    }
    {
        int rc = ramp_sync(e);
        if (rc != RAMP_OK) return rc;
    }
    return e->env_h_need[4] != 0 ? ramp_check_status(e, nullptr, nullptr) : RAMP_OK;
}


int ramp_enable_tick_lists(ramp_engine_t* e, int32_t cap) {
    if (!e || cap < 1) return set_error(RAMP_ERR_BAD_ARG, "ramp_enable_tick_lists: cap must be >= 1");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (e->ep.tick_util) { cudaFree(e->ep.tick_util); cudaFree(e->ep.tick_util_n); e->ep.tick_util = nullptr; e->ep.tick_util_n = nullptr; }
    const size_t B = (size_t)e->cfg.n_episodes;
    CUDA_TRY(cudaMalloc(&e->ep.tick_util, sizeof(double) * B * (size_t)cap * 2));
    CUDA_TRY(cudaMalloc(&e->ep.tick_util_n, sizeof(int32_t) * B));
    CUDA_TRY(cudaMemset(e->ep.tick_util_n, 0, sizeof(int32_t) * B));
    e->ep.tick_util_cap = cap;
    return RAMP_OK;
}

int ramp_get_tick_lists(ramp_engine_t* e, int32_t episode, double* mounted_out, double* cluster_out, int32_t cap, int32_t* n_out) {
    if (!e || !n_out) return set_error(RAMP_ERR_BAD_ARG, "null argument");
    if (!e->ep.tick_util) return set_error(RAMP_ERR_BAD_ARG, "per-tick lists are not recorded (ramp_enable_tick_lists)");
    if (episode < 0 || episode >= e->cfg.n_episodes) return set_error(RAMP_ERR_BAD_ARG, "bad episode %d", episode);
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    int32_t n = 0;
    CUDA_TRY(cudaMemcpy(&n, e->ep.tick_util_n + episode, sizeof(int32_t), cudaMemcpyDeviceToHost));
    *n_out = n;
    if (n > e->ep.tick_util_cap)
        return set_error(RAMP_ERR_CAPACITY, "episode %d: the step had %d outer-loop iterations, the per-tick lists hold %d", episode, n, e->ep.tick_util_cap);
    const int m = std::min(n, cap);
    if (m > 0 && mounted_out && cluster_out) {
        std::vector<double> rows((size_t)m * 2);
        CUDA_TRY(cudaMemcpy(rows.data(), e->ep.tick_util + (size_t)episode * e->ep.tick_util_cap * 2, sizeof(double) * rows.size(), cudaMemcpyDeviceToHost));
        for (int k = 0; k < m; ++k) { mounted_out[k] = rows[2 * k]; cluster_out[k] = rows[2 * k + 1]; }
    }
    return RAMP_OK;
}

int ramp_get_last_step_stats(ramp_engine_t* e, double* stats_out, int32_t* n_cluster_steps_out) {
    if (!e) return set_error(RAMP_ERR_BAD_ARG, "null engine");
    const int B = e->cfg.n_episodes;
    if (stats_out) CUDA_TRY(cudaMemcpyAsync(stats_out, e->d_step_stats, sizeof(double) * RAMP_STEP_STATS_LEN * B, cudaMemcpyDeviceToHost, e->stream));
    if (n_cluster_steps_out) CUDA_TRY(cudaMemcpyAsync(n_cluster_steps_out, e->d_n_cluster_steps, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, e->stream));
    return ramp_sync(e);
}


int ramp_env_read_state(ramp_engine_t* e, uint64_t* busy_out, int32_t* actions_out, int32_t* n_decided_out) {
    if (!e || !e->has_env) return set_error(RAMP_ERR_BAD_ARG, "no environment");
    const EnvDev& v = e->env;
    if (busy_out) CUDA_TRY(cudaMemcpyAsync(busy_out, v.busy, sizeof(uint64_t) * (size_t)v.B * v.n_words, cudaMemcpyDeviceToHost, e->stream));
    if (actions_out) CUDA_TRY(cudaMemcpyAsync(actions_out, v.actions, sizeof(int32_t) * v.B, cudaMemcpyDeviceToHost, e->stream));
    if (n_decided_out) CUDA_TRY(cudaMemcpyAsync(n_decided_out, v.n_decided, sizeof(int32_t) * v.B, cudaMemcpyDeviceToHost, e->stream));
    return ramp_sync(e);
}

}  // extern "C"
