// ramp_lookahead_thread.cuh -- _run_lookahead (RCE:379-467) on QUOTIENT templates: ONE THREAD per lookahead.
//
// ramp_register_template folds every lowered job by its symmetries (ramp_quotient.cpp): the bench's ResNet-50-like job
// is 330 op classes and 887-1,413 dep entries at every partition degree, and its ready frontiers hold 0-6 items per
// tick.  There is nothing left for a warp to share, so each lookahead runs on ONE lane, scalar, with no shuffles,
// ballots, atomics or barriers in the tick loop:
//
//   * a CTA is one warp; it pulls CHUNKS of up to 32 work items that use the same template (ramp_bucket_kernel groups the
//     step's memo misses by template), so its lanes run the same instruction stream over the same template -- no
//     divergence, and every template read is a shared-memory broadcast;
//   * the template blob (header + op records + rows + thresholds + packed dep words + run times) is copied into shared
//     memory ONCE per chunk with a bulk async copy (cp.async.bulk -> UBLKCP, completion on an mbarrier), so a tick never
//     waits for L2;
//   * per-lane state is lane-interleaved in shared memory ([slot][lane]: conflict-free): u16 parent counters per op
//     class, the ready-op / ready-flow / ready-non-flow frontiers (first OCAP / FCAP / NFCAP entries; the rest spills to an
//     HBM slab laid out the same way), and small per-worker-group / per-channel-group winner tables.
//
// Per tick (letters as in SURVEY.md 3.3; same arithmetic, same order as ramp_lookahead_kernel):
//   A,B  winners among the ready op classes per worker group (largest rank key), t_op = min remaining, active workers =
//        sum of the winners' class sizes                                                        (RCE:562-606, 44-67, 709-715)
//   C    a ready non-flow dep makes this a zero-length tick that completes exactly the non-flows  (RCE:412-422, 520-540, 718-731)
//   D    else t_comm = min remaining over the per-channel-group winners among the ready flows     (RCE:608-663)
//   E    tick = min(t_op, t_comm)                                                                (RCE:426)
//   H    every ready flow: rem -= min(tick, rem); == 0 -> its child class's counter += entry size; the class is readied
//        when the counter passes through its threshold (n_parents x class size)                 (RCE:733-775, JOB:525-536)
//   G    winning op classes tick; == 0 -> completed, their out-entries become ready (first ticked next tick, RCE:429)
//   I,J  t / comm / comp and the trace, f64, tick order                                           (RCE:442-445, 777-791)
#pragma once

namespace ramp {

#ifndef RAMP_T_OCAP
#define RAMP_T_OCAP 8       // ready op classes kept in shared memory per lane
#endif
#ifndef RAMP_T_FCAP
#define RAMP_T_FCAP 16      // ready flow entries kept in shared memory per lane
#endif
#ifndef RAMP_T_NFCAP
#define RAMP_T_NFCAP 8      // ready non-flow entries kept in shared memory per lane
#endif
#define RAMP_T_WCAP 8       // worker groups / channel groups with a per-lane winner table (more: pairwise comparison)
#define RAMP_T_CCAP 8

// header of a resident template blob (the blob is what the bulk copy moves: 16-byte aligned, size a multiple of 16)
struct ResHeader {
    int32_t n_ops, n_deps, n_workers, n_channels;      // classes, entries, worker groups, channel groups
    int32_t n_src, num_training_steps, orig_workers, _pad0;
    uint32_t kmask, cmask, imask, _pad1;               // dep word: key | chan << cshift | flow << fshift | inc << ishift | child << dshift
    int32_t cshift, fshift, ishift, dshift;
    int32_t off_op_row, off_op_thr, off_dep_kd, off_dep_rt;   // byte offsets from the blob start (op records follow the header)
    int32_t off_src, total_bytes, _pad2, _pad3;
};
static_assert(sizeof(ResHeader) == 96, "resident header is 96 bytes");

struct ChunkDesc { int32_t template_id, count; };      // up to 32 work items of one template; items at [chunk * 32 + lane]

struct ThreadArgs {
    const TemplateDev* templates;
    const ChunkDesc* chunks;
    const int32_t* n_chunks;        // device-side count
    int32_t* cursor;                // device-side chunk cursor (persistent CTAs pull chunks)
    const WorkItem* items;          // [chunk][32]
    unsigned char* scratch;         // [gridDim.x][scratch_stride]: per-CTA spill + temp trace
    uint64_t scratch_stride;
    ResultSlots res;
    TracePool pool;
    int32_t trace_cap;
    int32_t tmpl_cap;               // bytes of shared memory reserved for the template blob
    int32_t n_cap;                  // parent-counter slots per lane in shared memory
    int32_t spill_ops, spill_deps;  // per-lane spill capacities (entries) in the HBM slab
    MemoStats* stats;
};

__host__ __device__ inline size_t thread_smem_bytes(int tmpl_cap, int n_cap) {
    size_t per_lane = (size_t)RAMP_T_FCAP * 16 + (size_t)RAMP_T_OCAP * 20 + (size_t)RAMP_T_NFCAP * 8
                      + (size_t)(RAMP_T_WCAP + RAMP_T_CCAP) * 4 + (size_t)n_cap * 2;
    return (size_t)tmpl_cap + 32 * per_lane + 64;
}
__host__ __device__ inline uint64_t thread_scratch_bytes(int spill_ops, int spill_deps, int trace_cap) {
    // per lane: ops spill (record 16 + index 4), flows spill (16), non-flow spill (8), temp trace (tick 8 + n 4)
    const uint64_t per_lane = (uint64_t)spill_ops * 20 + (uint64_t)spill_deps * 24 + (uint64_t)trace_cap * 12;
    return align_up(per_lane * 32, 256);
}

// ---------------------------------------------------------------------------------------------------
// groups a step's resident work items by template: chunks of up to 32 items of one template (single CTA)
struct BucketArgs {
    const WorkItem* items;          // unsorted memo misses whose template is resident
    int32_t* n_items;               // zeroed on exit (the next step's plan kernel appends from 0)
    int32_t n_templates;
    int32_t* tcount;                // [n_templates + 1] zero on entry, zero on exit
    int32_t* tbase;                 // [n_templates + 1] scratch: first chunk of each template
    WorkItem* chunk_items;          // [max_chunks][32]
    ChunkDesc* chunks;              // [max_chunks]
    int32_t* n_chunks;
    int32_t* cursor;                // reset to 0 here
    int32_t* rank;                  // [B] scratch
};

__global__ void __launch_bounds__(1024) ramp_bucket_kernel(const BucketArgs a) {
    __shared__ int s_scan[1024];
    __shared__ int s_carry;
    const int n = *a.n_items;
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += blockDim.x) a.rank[i] = atomicAdd(&a.tcount[a.items[i].template_id], 1);
    if (tid == 0) s_carry = 0;
    __syncthreads();
    // exclusive scan of chunks-per-template, 1024 templates per round
    for (int base = 0; base < a.n_templates; base += blockDim.x) {
        const int t = base + tid;
        const int cnt = (t < a.n_templates) ? a.tcount[t] : 0;
        const int nch = (cnt + 31) >> 5;
        s_scan[tid] = nch;
        __syncthreads();
        for (int o = 1; o < (int)blockDim.x; o <<= 1) {
            const int v = (tid >= o) ? s_scan[tid - o] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int excl = s_carry + s_scan[tid] - nch;
        if (t < a.n_templates) {
            a.tbase[t] = excl;
            for (int k = 0; k < nch; ++k) {
                ChunkDesc d; d.template_id = t; d.count = (cnt - 32 * k < 32) ? (cnt - 32 * k) : 32;
                a.chunks[excl + k] = d;
            }
        }
        __syncthreads();
        if (tid == blockDim.x - 1) s_carry += s_scan[tid];
        __syncthreads();
    }
    for (int i = tid; i < n; i += blockDim.x) {
        const WorkItem it = a.items[i];
        const int r = a.rank[i];
        a.chunk_items[(size_t)(a.tbase[it.template_id] + (r >> 5)) * 32 + (r & 31)] = it;
    }
    __syncthreads();
    for (int t = tid; t < a.n_templates; t += blockDim.x) a.tcount[t] = 0;
    if (tid == 0) { *a.n_chunks = s_carry; *a.cursor = 0; *a.n_items = 0; }
}

// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// lane-interleaved arrays: element k of this lane at base[k * 32 + lane]; the first CAP entries in shared memory,
// the rest in the CTA's HBM slab
// one ready op class: {remaining.lo, remaining.hi, key, worker group | class size << 16} + its class index;
// one ready flow entry: {remaining.lo, remaining.hi, dep word.lo, dep word.hi}: nothing the tick loop needs is behind a second load
struct LaneOps {
    int4* a_sm; int32_t* i_sm; int4* a_gl; int32_t* i_gl; int lane;
    __device__ __forceinline__ int4 rec(int k) const { return (k < RAMP_T_OCAP) ? a_sm[k * 32 + lane] : a_gl[(size_t)(k - RAMP_T_OCAP) * 32 + lane]; }
    __device__ __forceinline__ int idx(int k) const { return (k < RAMP_T_OCAP) ? i_sm[k * 32 + lane] : i_gl[(size_t)(k - RAMP_T_OCAP) * 32 + lane]; }
    __device__ __forceinline__ void put(int k, const int4 r, int i) const {
        if (k < RAMP_T_OCAP) { a_sm[k * 32 + lane] = r; i_sm[k * 32 + lane] = i; }
        else { a_gl[(size_t)(k - RAMP_T_OCAP) * 32 + lane] = r; i_gl[(size_t)(k - RAMP_T_OCAP) * 32 + lane] = i; }
    }
};
struct LaneFlows {
    int4* sm; int4* gl; int lane;
    __device__ __forceinline__ int4 get(int k) const { return (k < RAMP_T_FCAP) ? sm[k * 32 + lane] : gl[(size_t)(k - RAMP_T_FCAP) * 32 + lane]; }
    __device__ __forceinline__ void put(int k, const int4 v) const { if (k < RAMP_T_FCAP) sm[k * 32 + lane] = v; else gl[(size_t)(k - RAMP_T_FCAP) * 32 + lane] = v; }
};
struct LaneNF {
    unsigned long long* sm; unsigned long long* gl; int lane;
    __device__ __forceinline__ unsigned long long get(int k) const { return (k < RAMP_T_NFCAP) ? sm[k * 32 + lane] : gl[(size_t)(k - RAMP_T_NFCAP) * 32 + lane]; }
    __device__ __forceinline__ void put(int k, unsigned long long v) const { if (k < RAMP_T_NFCAP) sm[k * 32 + lane] = v; else gl[(size_t)(k - RAMP_T_NFCAP) * 32 + lane] = v; }
};
__device__ __forceinline__ unsigned long long kd_of(const int4 f) { return ((unsigned long long)(uint32_t)f.w << 32) | (unsigned long long)(uint32_t)f.z; }

__global__ void __launch_bounds__(32) ramp_lookahead_thread_kernel(const ThreadArgs a) {
    extern __shared__ __align__(128) unsigned char smem_thr[];
    __shared__ __align__(8) unsigned long long mbar;
    const int lane = threadIdx.x;
    unsigned char* tm = smem_thr;                                   // template blob
    unsigned char* st = smem_thr + a.tmpl_cap;                      // per-lane state, by decreasing alignment
    int4* f_sm = reinterpret_cast<int4*>(st);                                            // [FCAP][32]
    int4* o_sm = f_sm + RAMP_T_FCAP * 32;                                                // [OCAP][32]
    unsigned long long* nf_sm = reinterpret_cast<unsigned long long*>(o_sm + RAMP_T_OCAP * 32);   // [NFCAP][32]
    int32_t*  oi_sm = reinterpret_cast<int32_t*>(nf_sm + RAMP_T_NFCAP * 32);            // [OCAP][32]
    uint32_t* wk_sm = reinterpret_cast<uint32_t*>(oi_sm + RAMP_T_OCAP * 32);             // [WCAP][32]
    uint32_t* ck_sm = wk_sm + RAMP_T_WCAP * 32;                                          // [CCAP][32]
    uint16_t* cnt_sm = reinterpret_cast<uint16_t*>(ck_sm + RAMP_T_CCAP * 32);            // [n_cap][32]

    unsigned char* slab = a.scratch + (uint64_t)blockIdx.x * a.scratch_stride;
    int4* f_gl = reinterpret_cast<int4*>(slab);                                          // [spill_deps][32]
    int4* o_gl = f_gl + (size_t)a.spill_deps * 32;                                       // [spill_ops][32]
    unsigned long long* nf_gl = reinterpret_cast<unsigned long long*>(o_gl + (size_t)a.spill_ops * 32);   // [spill_deps][32]
    double*  tr_tick = reinterpret_cast<double*>(nf_gl + (size_t)a.spill_deps * 32);    // [trace_cap][32]
    int32_t* oi_gl = reinterpret_cast<int32_t*>(tr_tick + (size_t)a.trace_cap * 32);    // [spill_ops][32]
    int32_t* tr_n = oi_gl + (size_t)a.spill_ops * 32;                                    // [trace_cap][32]

    const LaneOps ops{o_sm, oi_sm, o_gl, oi_gl, lane};
    const LaneFlows flows{f_sm, f_gl, lane};
    const LaneNF nfs{nf_sm, nf_gl, lane};
    const double INF = __longlong_as_double(RAMP_INF_BITS);

    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;
    int loaded = -1;

    for (;;) {
        int c = 0;
        if (lane == 0) c = atomicAdd(a.cursor, 1);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= *a.n_chunks) break;
        const ChunkDesc ch = a.chunks[c];
        const TemplateDev& TD = a.templates[ch.template_id];
        if (ch.template_id != loaded) {
            // every lane is done with the previous template (convergence point above); order those generic-proxy accesses
            // before the async-proxy writes of the bulk copy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)TD.res_bytes;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(tm)), "l"(TD.res_blob), "r"(bytes), "r"(smem_u32(&mbar)) : "memory");
            }
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "RAMP_WAIT_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra RAMP_DONE_%=;\n"
                "bra RAMP_WAIT_%=;\n"
                "RAMP_DONE_%=:\n"
                "}\n" ::"r"(smem_u32(&mbar)), "r"(phase) : "memory");
            phase ^= 1u;
            loaded = ch.template_id;
        }
        if (lane < ch.count) {
            const WorkItem item = a.items[(size_t)c * 32 + lane];
            const ResHeader& H = *reinterpret_cast<const ResHeader*>(tm);
            const int4* op_rec = reinterpret_cast<const int4*>(tm + sizeof(ResHeader));           // {cost.lo, cost.hi, key, worker | weight << 16}
            const int2* op_row = reinterpret_cast<const int2*>(tm + H.off_op_row);
            const uint32_t* op_thr = reinterpret_cast<const uint32_t*>(tm + H.off_op_thr);
            const unsigned long long* dep_kd = reinterpret_cast<const unsigned long long*>(tm + H.off_dep_kd);
            const double* dep_rt = reinterpret_cast<const double*>(tm + H.off_dep_rt);
            const int32_t* src_ops = reinterpret_cast<const int32_t*>(tm + H.off_src);
            const int N = H.n_ops, E = H.n_deps, W = H.n_workers, C = H.n_channels;
            const uint32_t kmask = H.kmask, cmask = H.cmask, imask = H.imask;
            const int csh = H.cshift, fsh = H.fshift, ish = H.ishift, dsh = H.dshift;
            const bool one_w = (W == 1), one_c = (C <= 1);
            const bool tab_w = (W <= RAMP_T_WCAP), tab_c = (C <= RAMP_T_CCAP);
            for (int i = 0; i < N && i < a.n_cap; ++i) cnt_sm[i * 32 + lane] = 0;

            int nO = H.n_src, nF = 0, nNF = 0;
            for (int k = 0; k < nO; ++k) { const int op = src_ops[k]; ops.put(k, op_rec[op], op); }   // RCE:1334
            int ops_completed = 0, deps_completed = 0, tick_no = 0;
            int status = (N <= a.n_cap) ? RAMP_ST_OK : RAMP_ST_TABLE_FULL;                           // cannot happen (eligibility)
            double t = 0.0, comm = 0.0, comp = 0.0;

            while (status == RAMP_ST_OK) {
                // ---- A, B: winners per worker group: largest key; t_op = min of their remaining times ----
                double t_op = INF;
                int n_active = 0;
                uint32_t best_w = 0u;                 // one worker group: the winner's key
                if (nO > 0) {
                    if (one_w) {
                        for (int k = 0; k < nO; ++k) {
                            const int4 r = ops.rec(k);
                            if ((uint32_t)r.z > best_w) { best_w = (uint32_t)r.z; t_op = __hiloint2double(r.y, r.x); n_active = (int)((uint32_t)r.w >> 16); }
                        }
                    } else if (tab_w) {
                        for (int w = 0; w < W; ++w) wk_sm[w * 32 + lane] = 0u;
                        for (int k = 0; k < nO; ++k) {
                            const int4 r = ops.rec(k);
                            const int w = r.w & 0xffff;
                            if ((uint32_t)r.z > wk_sm[w * 32 + lane]) wk_sm[w * 32 + lane] = (uint32_t)r.z;
                        }
                        for (int k = 0; k < nO; ++k) {
                            const int4 r = ops.rec(k);
                            if (wk_sm[(r.w & 0xffff) * 32 + lane] == (uint32_t)r.z) {
                                const double rem = __hiloint2double(r.y, r.x);
                                t_op = (rem < t_op) ? rem : t_op;
                                n_active += (int)((uint32_t)r.w >> 16);
                            }
                        }
                    } else {
                        // more worker groups than table slots: pairwise comparison; the winners are marked in bit 31 of the key
                        // (keys are ranks <= N < 2^31) for phase G, which clears the mark
                        for (int k = 0; k < nO; ++k) {
                            int4 r = ops.rec(k);
                            bool win = true;
                            for (int j = 0; j < nO && win; ++j) {
                                const int4 r2 = ops.rec(j);
                                if ((r2.w & 0xffff) == (r.w & 0xffff) && ((uint32_t)r2.z & 0x7fffffffu) > (uint32_t)r.z) win = false;
                            }
                            if (win) {
                                const double rem = __hiloint2double(r.y, r.x);
                                t_op = (rem < t_op) ? rem : t_op;
                                n_active += (int)((uint32_t)r.w >> 16);
                                r.z = (int)((uint32_t)r.z | 0x80000000u);
                                ops.put(k, r, ops.idx(k));
                            }
                        }
                    }
                }
                // ---- C, D ----
                const bool any_nf = nNF > 0;
                double t_comm = 0.0;
                if (!any_nf) {
                    t_comm = INF;
                    if (nF > 0) {
                        if (one_c) {
                            uint32_t best = 0u;
                            for (int k = 0; k < nF; ++k) {
                                const int4 f = flows.get(k);
                                const unsigned long long kd = kd_of(f);
                                if (((uint32_t)(kd >> csh) & cmask) == cmask) continue;               // no channel: ticks, never a winner
                                const uint32_t key = (uint32_t)kd & kmask;
                                const double rem = __hiloint2double(f.y, f.x);
                                if (key > best) { best = key; t_comm = rem; }
                                else if (key == best) t_comm = (rem < t_comm) ? rem : t_comm;
                            }
                        } else if (tab_c) {
                            for (int q = 0; q < C; ++q) ck_sm[q * 32 + lane] = 0u;
                            for (int k = 0; k < nF; ++k) {
                                const unsigned long long kd = kd_of(flows.get(k));
                                const uint32_t q = (uint32_t)(kd >> csh) & cmask;
                                if (q != cmask) {
                                    const uint32_t key = (uint32_t)kd & kmask;
                                    if (key > ck_sm[q * 32 + lane]) ck_sm[q * 32 + lane] = key;
                                }
                            }
                            for (int k = 0; k < nF; ++k) {
                                const int4 f = flows.get(k);
                                const unsigned long long kd = kd_of(f);
                                const uint32_t q = (uint32_t)(kd >> csh) & cmask;
                                if (q != cmask && ck_sm[q * 32 + lane] == ((uint32_t)kd & kmask)) {
                                    const double rem = __hiloint2double(f.y, f.x);
                                    t_comm = (rem < t_comm) ? rem : t_comm;
                                }
                            }
                        } else {
                            for (int k = 0; k < nF; ++k) {
                                const int4 f = flows.get(k);
                                const unsigned long long kd = kd_of(f);
                                const uint32_t q = (uint32_t)(kd >> csh) & cmask;
                                if (q == cmask) continue;
                                const uint32_t key = (uint32_t)kd & kmask;
                                bool win = true;
                                for (int j = 0; j < nF && win; ++j) {
                                    const unsigned long long kd2 = kd_of(flows.get(j));
                                    if (((uint32_t)(kd2 >> csh) & cmask) == q && ((uint32_t)kd2 & kmask) > key) win = false;
                                }
                                if (win) { const double rem = __hiloint2double(f.y, f.x); t_comm = (rem < t_comm) ? rem : t_comm; }
                            }
                        }
                    }
                }
                // ---- E, I, J ----
                const double tick = (t_comm < t_op) ? t_comm : t_op;
                {
                    const bool ticked_ops = n_active > 0;
                    const bool ticked_flows = (!any_nf) && (nF > 0);                                  // RCE:434-439
                    if (ticked_flows) comm = __dadd_rn(comm, tick);
                    if (ticked_ops) comp = __dadd_rn(comp, tick);
                    t = __dadd_rn(t, tick);
                    if (tick_no < a.trace_cap) { tr_n[(size_t)tick_no * 32 + lane] = n_active; tr_tick[(size_t)tick_no * 32 + lane] = tick; }
                    else status = RAMP_ST_TRACE_OVERFLOW;
                    ++tick_no;
                }
                // ---- H ----
                int tailO = nO;                       // ops readied in this tick are appended behind the current frontier
                auto complete_dep = [&](const unsigned long long kd) {                              // JOB:525-536
                    const int child = (int)(kd >> dsh);
                    const uint32_t inc = (uint32_t)(kd >> ish) & imask;
                    const uint32_t old = cnt_sm[child * 32 + lane];
                    const uint32_t thr = op_thr[child];
                    const uint32_t neu = old + inc;
                    cnt_sm[child * 32 + lane] = (uint16_t)neu;
                    if (old < thr && thr <= neu) { ops.put(tailO, op_rec[child], child); ++tailO; }   // JOB:531 for every member
                };
                if (any_nf) {
                    for (int k = 0; k < nNF; ++k) complete_dep(nfs.get(k));
                    deps_completed += nNF;
                    nNF = 0;
                } else {
                    int p = 0;
                    for (int k = 0; k < nF; ++k) {
                        int4 f = flows.get(k);
                        const double r2 = tick_down(__hiloint2double(f.y, f.x), tick);              // JOB:561
                        if (r2 == 0.0) { complete_dep(kd_of(f)); ++deps_completed; }                // JOB:562
                        else { f.x = __double2loint(r2); f.y = __double2hiint(r2); flows.put(p, f); ++p; }
                    }
                    nF = p;
                }
                // ---- G ----
                int p = 0;
                for (int k = 0; k < nO; ++k) {
                    int4 r = ops.rec(k);
                    const int op = ops.idx(k);
                    bool win;
                    if (one_w) win = (uint32_t)r.z == best_w;
                    else if (tab_w) win = wk_sm[(r.w & 0xffff) * 32 + lane] == (uint32_t)r.z;
                    else { win = r.z < 0; r.z &= 0x7fffffff; }
                    if (win) {                                                                      // this tick's winner
                        const double rem = tick_down(__hiloint2double(r.y, r.x), tick);             // JOB:555
                        if (rem == 0.0) {                                                           // JOB:556
                            ++ops_completed;
                            const int2 row = op_row[op];
                            for (int e = row.x; e < row.x + row.y; ++e) {                           // JOB:496-506
                                const unsigned long long kd = dep_kd[e];
                                if ((kd >> fsh) & 1ull) {
                                    const double rt = dep_rt[e];
                                    flows.put(nF, make_int4(__double2loint(rt), __double2hiint(rt), (int)(uint32_t)kd, (int)(uint32_t)(kd >> 32)));
                                    ++nF;
                                } else { nfs.put(nNF, kd); ++nNF; }
                            }
                            continue;
                        }
                        r.x = __double2loint(rem); r.y = __double2hiint(rem);
                    }
                    ops.put(p, r, op); ++p;
                }
                for (int k = nO; k < tailO; ++k, ++p) { if (p != k) ops.put(p, ops.rec(k), ops.idx(k)); }
                nO = p;
                // ---- K, L ----
                const bool finished = (ops_completed == N) && (deps_completed == E);               // JOB:549-551
                if (finished) break;
                if (isinf(tick)) { status = RAMP_ST_INFINITE_TICK; break; }                         // RCE:462
            }

            // ---- results (RCE:450-452) ----
            const int n_rec = tick_no < a.trace_cap ? tick_no : a.trace_cap;
            const double steps = (double)H.num_training_steps;
            const double jct = __dmul_rn(t, steps);
            const int nmw = item.n_mounted_workers > 0 ? item.n_mounted_workers : H.orig_workers;
            const bool can_util = (status == RAMP_ST_OK);
            long long off = -1;
            if (a.pool.top != nullptr) {
                const unsigned long long o = atomicAdd(a.pool.top, (unsigned long long)n_rec);
                if (o + (unsigned long long)n_rec <= a.pool.len) off = (long long)o;
                else if (status == RAMP_ST_OK) status = RAMP_ST_TRACE_OVERFLOW;
            }
            double util = 0.0;
            {                                                                                       // RCE:830-832, tick order
                const double dn = (double)nmw;
                int32_t* pn = (off >= 0) ? a.pool.n_active + off : nullptr;
                double* pt = (off >= 0) ? a.pool.tick + off : nullptr;
#pragma unroll 4
                for (int k = 0; k < n_rec; ++k) {
                    const int nk = tr_n[(size_t)k * 32 + lane];
                    const double tk = tr_tick[(size_t)k * 32 + lane];
                    if (pn) { pn[k] = nk; pt[k] = tk; }
                    if (can_util) util = __dadd_rn(util, __dmul_rn(__ddiv_rn((double)nk, dn), __ddiv_rn(tk, jct)));
                }
            }
            a.res.jct[item.slot] = jct;
            a.res.comm[item.slot] = __dmul_rn(comm, steps);
            a.res.comp[item.slot] = __dmul_rn(comp, steps);
            a.res.n_ticks[item.slot] = tick_no;
            a.res.util[item.slot] = can_util ? util : 0.0;
            a.res.util_nmw[item.slot] = can_util ? nmw : -1;
            a.res.trace_off[item.slot] = off;
            a.res.status[item.slot] = status;
            if (a.stats) {
                atomicAdd(&a.stats->lookaheads, 1ull);
                atomicAdd(&a.stats->alg_bytes, (unsigned long long)(TD.algorithmic_bytes_static + 12ull * (unsigned long long)tick_no));
                atomicAdd(&a.stats->quotient_bytes, (unsigned long long)(20ull * N + 19ull * E + 24ull + 12ull * (unsigned long long)tick_no));
            }
        }
        __syncwarp();
    }
}

}  // namespace ramp
