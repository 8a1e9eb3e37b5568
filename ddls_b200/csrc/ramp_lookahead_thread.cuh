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
#ifndef RAMP_T_FASTF
#define RAMP_T_FASTF 6      // frontiers of up to this many flow entries (and 2 op classes) run the register-resident path
#endif
#ifndef RAMP_T_NFCAP
#define RAMP_T_NFCAP 8      // ready non-flow entries kept in shared memory per lane
#endif
#define RAMP_T_WCAP 8       // worker groups / channel groups with a per-lane winner table (more: pairwise comparison)
#define RAMP_T_CCAP 32

// header of a resident template blob (the blob is what the bulk copy moves: 16-byte aligned, size a multiple of 16)
struct ResHeader {
    int32_t n_ops, n_deps, n_workers, n_channels;      // classes, entries, worker groups, channel groups
    int32_t n_src, num_training_steps, orig_workers, _pad0;
    uint32_t kmask, cmask, imask, _pad1;               // dep word lo: key | SET of channel groups << cshift (empty: no channel);
    int32_t cshift, fshift, ishift, dshift;            // dep word hi: flow | inc << 1 | child << dshift  (fshift = 0, ishift = 1)
    int32_t off_op_row, off_op_thr, off_dep_kd, off_dep_rt;   // byte offsets from the blob start (op records follow the header)
    int32_t off_src, total_bytes, _pad2, _pad3;
};
static_assert(sizeof(ResHeader) == 96, "resident header is 96 bytes");

struct ChunkDesc { int32_t template_id, count; };      // up to 32 work items of one template; items at [chunk * 32 + lane]

// recorded by the first lookahead of a template that completes (deterministic per template): lets later ones write their trace
// in place and keep every list in shared memory
struct __align__(16) TemplateHints { int32_t n_ticks, max_o, max_f, max_nf; };   // read / written as ONE 16-byte access


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
    TemplateHints* hints;           // [max_templates], zero = nothing recorded yet
    double* hint_jct;               // [max_templates] job completion time the first lookahead of the template found (0 = none):
                                    //   lets later ones accumulate the utilisation inside the tick loop (RCE:830-832 divides by it)
};

__host__ __device__ inline size_t thread_smem_bytes(int tmpl_cap, int n_cap) {
    size_t per_lane = (size_t)RAMP_T_FCAP * 16 + (size_t)RAMP_T_OCAP * 20 + (size_t)RAMP_T_NFCAP * 4
                      + (size_t)(RAMP_T_WCAP + RAMP_T_CCAP) * 4 + (size_t)n_cap * 2;
    return (size_t)tmpl_cap + 32 * per_lane + 64;
}
__host__ __device__ inline uint64_t thread_scratch_bytes(int spill_ops, int spill_deps, int trace_cap) {
    // per lane: ops spill (record 16 + index 4), flows spill (16), non-flow spill (4), temp trace (tick 8 + n 4)
    const uint64_t per_lane = (uint64_t)spill_ops * 20 + (uint64_t)spill_deps * 20 + (uint64_t)trace_cap * 12;
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

// Lane-interleaved per-lane lists (element k of this lane at base[k * 32 + lane]).  SPILL: entries past the shared-memory
// capacity live in the CTA's HBM slab; !SPILL: the template's recorded frontier sizes (TemplateHints) fit the capacity.
//   ready op class   {remaining.lo, remaining.hi, key, worker group | class size << 16} + its class index
//   ready flow entry {remaining.lo, remaining.hi, dep word lo (key | set of channel groups << cshift), dep word hi (flow | inc << 1 | child << dshift)}
//   ready non-flow   dep word hi
// nothing the tick loop needs about a ready item is behind a second load
template <bool SPILL>
struct LaneOps {
    int4* a_sm; int32_t* i_sm; int4* a_gl; int32_t* i_gl; int lane;
    __device__ __forceinline__ int4 rec(int k) const {
        if (!SPILL || k < RAMP_T_OCAP) return a_sm[k * 32 + lane];
        return a_gl[(size_t)(k - RAMP_T_OCAP) * 32 + lane];
    }
    __device__ __forceinline__ int idx(int k) const {
        if (!SPILL || k < RAMP_T_OCAP) return i_sm[k * 32 + lane];
        return i_gl[(size_t)(k - RAMP_T_OCAP) * 32 + lane];
    }
    __device__ __forceinline__ void put(int k, const int4 r, int i) const {
        if (!SPILL || k < RAMP_T_OCAP) { a_sm[k * 32 + lane] = r; i_sm[k * 32 + lane] = i; }
        else { a_gl[(size_t)(k - RAMP_T_OCAP) * 32 + lane] = r; i_gl[(size_t)(k - RAMP_T_OCAP) * 32 + lane] = i; }
    }
};
template <bool SPILL>
struct LaneFlows {
    int4* sm; int4* gl; int lane;
    __device__ __forceinline__ int4 get(int k) const {
        if (!SPILL || k < RAMP_T_FCAP) return sm[k * 32 + lane];
        return gl[(size_t)(k - RAMP_T_FCAP) * 32 + lane];
    }
    __device__ __forceinline__ void put(int k, const int4 v) const {
        if (!SPILL || k < RAMP_T_FCAP) sm[k * 32 + lane] = v; else gl[(size_t)(k - RAMP_T_FCAP) * 32 + lane] = v;
    }
};
template <bool SPILL>
struct LaneNF {
    uint32_t* sm; uint32_t* gl; int lane;
    __device__ __forceinline__ uint32_t get(int k) const {
        if (!SPILL || k < RAMP_T_NFCAP) return sm[k * 32 + lane];
        return gl[(size_t)(k - RAMP_T_NFCAP) * 32 + lane];
    }
    __device__ __forceinline__ void put(int k, uint32_t v) const {
        if (!SPILL || k < RAMP_T_NFCAP) sm[k * 32 + lane] = v; else gl[(size_t)(k - RAMP_T_NFCAP) * 32 + lane] = v;
    }
};

// Remaining times are non-negative doubles (validated and canonicalised to +0 at registration): their u64 bit patterns order like
// the values, so every comparison of the tick loop -- winners' minimum, tick = min(t_comm, t_op), "did it complete" -- is a 64-bit
// INTEGER comparison (two ISETP) instead of an FP64 one.  On sm_100a DSETP / DADD results come back through the long scoreboard
// (profiles/r2_ncu_thread_source.md: 20 % of all stall samples sat on four DSETP after tick_down); only the subtraction of the
// survivors and the three accumulators stay FP64, off the control path.
//   x -= min(tick, x) == 0  (JOB:555-556, 561-562)  <=>  x <= tick   (x > tick >= 0 gives x - tick > 0: no underflow to zero)
typedef unsigned long long u64_t;
__device__ __forceinline__ u64_t rem_bits(const int4& r) { return ((u64_t)(uint32_t)r.y << 32) | (u64_t)(uint32_t)r.x; }

struct LaneCtx {                      // what one lane's lookahead works on
    const unsigned char* tm;         // template blob in shared memory
    int4* f_sm; int4* o_sm; uint32_t* nf_sm; int32_t* oi_sm; uint32_t* wk_sm; uint32_t* ck_sm; uint16_t* cnt_sm;
    int4* f_gl; int4* o_gl; uint32_t* nf_gl; int32_t* oi_gl;
    int32_t* tr_n; double* tr_tick;  // trace destination: element k at [k * tr_stride]
    int tr_stride, tr_cap;
    int lane, n_cap;
    double util_jct, util_dn;        // != 0: accumulate sum (n_active / util_dn) * (tick / util_jct) in tick order (RCE:830-832)
};

struct LaneResult { double t, comm, comp, util; int tick_no, status, max_o, max_f, max_nf; };

// _run_lookahead for one lane.  SPILL = false: every frontier fits its shared-memory capacity (TemplateHints);
// SIMPLE = true: one worker group and at most one channel group (the usual quotient of a partitioned job): the winner is the
// largest key, no tables.
template <bool SPILL, bool SIMPLE>
__device__ __forceinline__ LaneResult thread_lookahead(const LaneCtx& x) {
    const int lane = x.lane;
    const ResHeader& H = *reinterpret_cast<const ResHeader*>(x.tm);
    const int4* op_rec = reinterpret_cast<const int4*>(x.tm + sizeof(ResHeader));                  // {cost.lo, cost.hi, key, worker | weight << 16}
    const int2* op_row = reinterpret_cast<const int2*>(x.tm + H.off_op_row);
    const uint32_t* op_thr = reinterpret_cast<const uint32_t*>(x.tm + H.off_op_thr);
    const uint2* dep_kd = reinterpret_cast<const uint2*>(x.tm + H.off_dep_kd);
    const double* dep_rt = reinterpret_cast<const double*>(x.tm + H.off_dep_rt);
    const int32_t* src_ops = reinterpret_cast<const int32_t*>(x.tm + H.off_src);
    const int N = H.n_ops, E = H.n_deps, W = H.n_workers, C = H.n_channels;
    const uint32_t kmask = H.kmask, imask = H.imask;
    const int csh = H.cshift, dsh = H.dshift;
    const bool tab_w = (W <= RAMP_T_WCAP), tab_c = (C <= RAMP_T_CCAP);
    const LaneOps<SPILL> ops{x.o_sm, x.oi_sm, x.o_gl, x.oi_gl, lane};
    const LaneFlows<SPILL> flows{x.f_sm, x.f_gl, lane};
    const LaneNF<SPILL> nfs{x.nf_sm, x.nf_gl, lane};
    uint16_t* cnt = x.cnt_sm + lane;
    uint32_t* wk = x.wk_sm + lane;
    uint32_t* ck = x.ck_sm + lane;
    const double INF = __longlong_as_double(RAMP_INF_BITS);

    for (int i = 0; i < N && i < x.n_cap; ++i) cnt[i * 32] = 0;
    int nO = H.n_src, nF = 0, nNF = 0;
    for (int k = 0; k < nO; ++k) { const int op = src_ops[k]; ops.put(k, op_rec[op], op); }       // RCE:1334
    LaneResult R;
    R.t = 0.0; R.comm = 0.0; R.comp = 0.0; R.util = 0.0; R.tick_no = 0; R.max_o = nO; R.max_f = 0; R.max_nf = 0;
    R.status = (N <= x.n_cap) ? RAMP_ST_OK : RAMP_ST_TABLE_FULL;                                    // cannot happen (eligibility)
    int to_complete = N + E;              // ops and deps still to complete (JOB:549-551)
    const bool do_util = x.util_jct != 0.0;
    int util_last_n = -1;
    double util_last_q = 0.0;
    // the term of one tick (RCE:830-832); a tick with no active worker or no length adds +0.0 and is skipped (see the epilogue)
    auto add_util = [&](const int n_active, const double tick) {
        if (n_active != util_last_n) { util_last_n = n_active; util_last_q = __ddiv_rn((double)n_active, x.util_dn); }
        R.util = __dadd_rn(R.util, __dmul_rn(util_last_q, __ddiv_rn(tick, x.util_jct)));
    };
    int tr_idx = 0;                       // trace element k lives at [k * tr_stride]

    if (R.status == RAMP_ST_OK) for (;;) {       // left through ONE combined exit test per tick
        if (nF <= RAMP_T_FASTF && nO <= 2) {
            // ======== small frontiers (the usual case on a quotient): every ready item is loaded ONCE into registers and each
            // phase runs code specialised for the exact number of ready ops (0-2) and flows (0-4): winners by pairwise
            // comparison (no tables, any number of worker / channel groups), no loop or predication overhead ========
            static_assert(RAMP_T_FASTF == 6, "the dispatch below has cases for up to 6 ready flow entries");
            int4 fr[RAMP_T_FASTF], orr[2];
            int oi[2];
            bool ow0 = false, ow1 = false;
            u64_t t_op = RAMP_INF_BITS;
            int n_active = 0;
            // ---- A, B ----
            if (nO >= 1) {
                orr[0] = x.o_sm[lane]; oi[0] = x.oi_sm[lane];
                ow0 = true;
                if (nO == 2) {
                    orr[1] = x.o_sm[32 + lane]; oi[1] = x.oi_sm[32 + lane];
                    ow1 = true;
                    if (((orr[0].w ^ orr[1].w) & 0xffff) == 0) { if ((uint32_t)orr[0].z > (uint32_t)orr[1].z) ow1 = false; else ow0 = false; }
                    if (ow1) { t_op = rem_bits(orr[1]); n_active = (int)((uint32_t)orr[1].w >> 16); }
                }
                if (ow0) { const u64_t r0 = rem_bits(orr[0]); t_op = (r0 < t_op) ? r0 : t_op; n_active += (int)((uint32_t)orr[0].w >> 16); }
            }
            // ---- C, D, E, I, J, H: dispatched ONCE on the number of ready flow entries; each case is straight-line code ----
            const bool any_nf = nNF > 0;
            u64_t tick_b = 0ull;
            double tick = 0.0;
            int tailO = nO;
            auto complete_dep = [&](const uint32_t hi) {                                            // JOB:525-536
                const int child = (int)(hi >> dsh);
                const uint32_t inc = (hi >> 1) & imask;
                const uint32_t old = cnt[child * 32];
                const uint32_t thr = op_thr[child];
                const uint32_t neu = old + inc;
                cnt[child * 32] = (uint16_t)neu;
                if (old < thr && thr <= neu) { ops.put(tailO, op_rec[child], child); ++tailO; }     // JOB:531 for every member
            };
            // E, I, J: the tick, the three accumulators, the utilisation term and the trace entry
            auto take_tick = [&](const u64_t t_comm, const bool ticked_flows) {
                tick_b = (t_comm < t_op) ? t_comm : t_op;
                tick = __longlong_as_double((long long)tick_b);
                if (ticked_flows) R.comm = __dadd_rn(R.comm, tick);                                  // RCE:434-439, 777-791
                if (n_active > 0) { R.comp = __dadd_rn(R.comp, tick); if (do_util && tick_b != 0ull) add_util(n_active, tick); }
                R.t = __dadd_rn(R.t, tick);
                if (R.tick_no < x.tr_cap) { x.tr_n[tr_idx] = n_active; x.tr_tick[tr_idx] = tick; tr_idx += x.tr_stride; }
                else R.status = RAMP_ST_TRACE_OVERFLOW;
                ++R.tick_no;
            };
            auto flow_tick = [&](auto nf_tag) {
                constexpr int NF = decltype(nf_tag)::value;
                uint32_t gm[NF], key[NF];
#pragma unroll
                for (int k = 0; k < NF; ++k) fr[k] = x.f_sm[k * 32 + lane];
#pragma unroll
                for (int k = 0; k < NF; ++k) { gm[k] = (uint32_t)fr[k].z >> csh; key[k] = (uint32_t)fr[k].z & kmask; }
                u64_t t_comm = RAMP_INF_BITS;
#pragma unroll
                for (int k = 0; k < NF; ++k) {
                    // the entry wins on a channel group of its set unless a ready entry with a larger key lies on that group too
                    // (an empty set -- no channel -- never wins, it only ticks)
                    uint32_t open_groups = gm[k];
#pragma unroll
                    for (int j = 0; j < NF; ++j) if (j != k && key[j] > key[k]) open_groups &= ~gm[j];
                    if (open_groups) { const u64_t rem = rem_bits(fr[k]); t_comm = (rem < t_comm) ? rem : t_comm; }
                }
                take_tick(t_comm, true);
                int p = 0;
#pragma unroll
                for (int k = 0; k < NF; ++k) {
                    const u64_t rb = rem_bits(fr[k]);
                    if (rb <= tick_b) { complete_dep((uint32_t)fr[k].w); --to_complete; }        // JOB:561-562
                    else {
                        const double r2 = __dsub_rn(__longlong_as_double((long long)rb), tick);
                        fr[k].x = __double2loint(r2); fr[k].y = __double2hiint(r2); x.f_sm[p * 32 + lane] = fr[k]; ++p;
                    }
                }
                nF = p;
            };
            // by frequency on the quotient of a partitioned job: one ready flow entry, a non-flow tick, none, two, ...
            if (any_nf) {                                           // zero-length tick that completes the ready non-flow deps
                take_tick(0ull, false);
                _Pragma("unroll 1")
                for (int k = 0; k < nNF; ++k) complete_dep(nfs.get(k));
                to_complete -= nNF;
                nNF = 0;
            } else {
                // a chain of compare-and-branch, most frequent count first; the empty asm keeps the compiler from folding the chain
                // back into a jump table (constant-bank load + indirect branch on the critical path of every tick)
                auto opaque = [](int v) { asm volatile("" : "+r"(v)); return v; };
                if (opaque(nF) == 1) flow_tick(std::integral_constant<int, 1>{});
                else if (opaque(nF) == 0) take_tick(RAMP_INF_BITS, false);
                else if (opaque(nF) == 2) flow_tick(std::integral_constant<int, 2>{});
                else if (opaque(nF) == 3) flow_tick(std::integral_constant<int, 3>{});
                else if (opaque(nF) == 4) flow_tick(std::integral_constant<int, 4>{});
                else if (opaque(nF) == 5) flow_tick(std::integral_constant<int, 5>{});
                else flow_tick(std::integral_constant<int, 6>{});
            }
            // ---- G ----
            int p = 0;
            auto tick_op = [&](int4 r, const int op, const bool win) {
                if (win) {                                                                          // this tick's winner
                    const u64_t rb = rem_bits(r);
                    if (rb <= tick_b) {                                                             // JOB:555-556
                        --to_complete;
                        const int2 row = op_row[op];
                        _Pragma("unroll 1")
                        for (int e = row.x; e < row.x + row.y; ++e) {                               // JOB:496-506
                            const uint2 kd = dep_kd[e];
                            if (kd.y & 1u) {
                                const double rt = dep_rt[e];
                                flows.put(nF, make_int4(__double2loint(rt), __double2hiint(rt), (int)kd.x, (int)kd.y));
                                ++nF;
                            } else { nfs.put(nNF, kd.y); ++nNF; }
                        }
                        return;
                    }
                    const double rem = __dsub_rn(__longlong_as_double((long long)rb), tick);
                    r.x = __double2loint(rem); r.y = __double2hiint(rem);
                }
                x.o_sm[p * 32 + lane] = r; x.oi_sm[p * 32 + lane] = op; ++p;
            };
            if (nO >= 1) {
                tick_op(orr[0], oi[0], ow0);
                if (nO == 2) tick_op(orr[1], oi[1], ow1);
            }
            if (tailO != nO) {
                _Pragma("unroll 1")
                for (int k = nO; k < tailO; ++k, ++p) { if (p != k) ops.put(p, ops.rec(k), ops.idx(k)); }
            }
            nO = p;
            if (SPILL) {
                R.max_o = (tailO > R.max_o) ? tailO : R.max_o;
                R.max_f = (nF > R.max_f) ? nF : R.max_f;
                R.max_nf = (nNF > R.max_nf) ? nNF : R.max_nf;
            }
            if ((to_complete == 0) | ((uint32_t)(tick_b >> 32) == 0x7FF00000u) | (R.status != RAMP_ST_OK)) {
                if (to_complete != 0 && (uint32_t)(tick_b >> 32) == 0x7FF00000u) R.status = RAMP_ST_INFINITE_TICK;   // JOB:549-551 first, then RCE:462
                break;
            }
            continue;
        }
        // ---- A, B: winners per worker group: largest key; t_op = min of their remaining times ----
        u64_t t_op = RAMP_INF_BITS;
        int n_active = 0;
        uint32_t best_w = 0u;                 // SIMPLE: the winner's key
        if (nO > 0) {
            if (SIMPLE) {
                _Pragma("unroll 1")
                for (int k = 0; k < nO; ++k) {
                    const int4 r = ops.rec(k);
                    if ((uint32_t)r.z > best_w) { best_w = (uint32_t)r.z; t_op = rem_bits(r); n_active = (int)((uint32_t)r.w >> 16); }
                }
            } else if (tab_w) {
                _Pragma("unroll 1")
                for (int w = 0; w < W; ++w) wk[w * 32] = 0u;
                _Pragma("unroll 1")
                for (int k = 0; k < nO; ++k) {
                    const int4 r = ops.rec(k);
                    const int w = r.w & 0xffff;
                    if ((uint32_t)r.z > wk[w * 32]) wk[w * 32] = (uint32_t)r.z;
                }
                _Pragma("unroll 1")
                for (int k = 0; k < nO; ++k) {
                    const int4 r = ops.rec(k);
                    if (wk[(r.w & 0xffff) * 32] == (uint32_t)r.z) {
                        const u64_t rem = rem_bits(r);
                        t_op = (rem < t_op) ? rem : t_op;
                        n_active += (int)((uint32_t)r.w >> 16);
                    }
                }
            } else {
                // more worker groups than table slots: pairwise comparison; the winners are marked in bit 31 of the key
                // (keys are ranks <= N < 2^31) for phase G, which clears the mark
                _Pragma("unroll 1")
                for (int k = 0; k < nO; ++k) {
                    int4 r = ops.rec(k);
                    bool win = true;
                    _Pragma("unroll 1")
                    for (int j = 0; j < nO && win; ++j) {
                        const int4 r2 = ops.rec(j);
                        if ((r2.w & 0xffff) == (r.w & 0xffff) && ((uint32_t)r2.z & 0x7fffffffu) > (uint32_t)r.z) win = false;
                    }
                    if (win) {
                        const u64_t rem = rem_bits(r);
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
        u64_t t_comm = 0ull;
        if (!any_nf) {
            t_comm = RAMP_INF_BITS;
            if (nF > 0) {
                if (SIMPLE) {
                    uint32_t best = 0u;
                    _Pragma("unroll 1")
                    for (int k = 0; k < nF; ++k) {
                        const int4 f = flows.get(k);
                        if (((uint32_t)f.z >> csh) == 0u) continue;                                  // no channel: ticks, never a winner
                        const uint32_t key = (uint32_t)f.z & kmask;
                        const u64_t rem = rem_bits(f);
                        if (key > best) { best = key; t_comm = rem; }
                        else if (key == best) t_comm = (rem < t_comm) ? rem : t_comm;
                    }
                } else if (tab_c) {
                    // per-group table: largest key among the ready entries whose set contains the group
                    _Pragma("unroll 1")
                    for (int q = 0; q < C; ++q) ck[q * 32] = 0u;
                    _Pragma("unroll 1")
                    for (int k = 0; k < nF; ++k) {
                        const uint32_t lo = (uint32_t)flows.get(k).z;
                        const uint32_t key = lo & kmask;
                        uint32_t m = lo >> csh;
                        while (m) { const int q = __ffs((int)m) - 1; m &= m - 1u; if (key > ck[q * 32]) ck[q * 32] = key; }
                    }
                    _Pragma("unroll 1")
                    for (int k = 0; k < nF; ++k) {
                        const int4 f = flows.get(k);
                        const uint32_t key = (uint32_t)f.z & kmask;
                        uint32_t m = (uint32_t)f.z >> csh;
                        bool win = false;
                        while (m && !win) { const int q = __ffs((int)m) - 1; m &= m - 1u; win = ck[q * 32] == key; }
                        if (win) { const u64_t rem = rem_bits(f); t_comm = (rem < t_comm) ? rem : t_comm; }
                    }
                } else {
                    _Pragma("unroll 1")
                    for (int k = 0; k < nF; ++k) {
                        const int4 f = flows.get(k);
                        uint32_t open_groups = (uint32_t)f.z >> csh;
                        if (open_groups == 0u) continue;
                        const uint32_t key = (uint32_t)f.z & kmask;
                        _Pragma("unroll 1")
                        for (int j = 0; j < nF && open_groups; ++j) {
                            const uint32_t lo2 = (uint32_t)flows.get(j).z;
                            if ((lo2 & kmask) > key) open_groups &= ~(lo2 >> csh);
                        }
                        if (open_groups) { const u64_t rem = rem_bits(f); t_comm = (rem < t_comm) ? rem : t_comm; }
                    }
                }
            }
        }
        // ---- E, I, J ----
        const u64_t tick_b = (t_comm < t_op) ? t_comm : t_op;
        const double tick = __longlong_as_double((long long)tick_b);
        {
            if ((!any_nf) && (nF > 0)) R.comm = __dadd_rn(R.comm, tick);                             // RCE:434-439, 777-791
            if (n_active > 0) { R.comp = __dadd_rn(R.comp, tick); if (do_util && tick_b != 0ull) add_util(n_active, tick); }
            R.t = __dadd_rn(R.t, tick);
            if (R.tick_no < x.tr_cap) { x.tr_n[tr_idx] = n_active; x.tr_tick[tr_idx] = tick; tr_idx += x.tr_stride; }
            else R.status = RAMP_ST_TRACE_OVERFLOW;
            ++R.tick_no;
        }
        // ---- H ----
        int tailO = nO;                       // ops readied in this tick are appended behind the current frontier
        auto complete_dep = [&](const uint32_t hi) {                                                // JOB:525-536
            const int child = (int)(hi >> dsh);
            const uint32_t inc = (hi >> 1) & imask;
            const uint32_t old = cnt[child * 32];
            const uint32_t thr = op_thr[child];
            const uint32_t neu = old + inc;
            cnt[child * 32] = (uint16_t)neu;
            if (old < thr && thr <= neu) { ops.put(tailO, op_rec[child], child); ++tailO; }         // JOB:531 for every member
        };
        if (any_nf) {
            _Pragma("unroll 1")
            for (int k = 0; k < nNF; ++k) complete_dep(nfs.get(k));
            to_complete -= nNF;
            nNF = 0;
        } else {
            int p = 0;
            _Pragma("unroll 1")
            for (int k = 0; k < nF; ++k) {
                int4 f = flows.get(k);
                const u64_t rb = rem_bits(f);
                if (rb <= tick_b) { complete_dep((uint32_t)f.w); --to_complete; }                // JOB:561-562
                else {
                    const double r2 = __dsub_rn(__longlong_as_double((long long)rb), tick);
                    f.x = __double2loint(r2); f.y = __double2hiint(r2); flows.put(p, f); ++p;
                }
            }
            nF = p;
        }
        // ---- G ----
        int p = 0;
        _Pragma("unroll 1")
        for (int k = 0; k < nO; ++k) {
            int4 r = ops.rec(k);
            const int op = ops.idx(k);
            bool win;
            if (SIMPLE) win = (uint32_t)r.z == best_w;
            else if (tab_w) win = wk[(r.w & 0xffff) * 32] == (uint32_t)r.z;
            else { win = r.z < 0; r.z &= 0x7fffffff; }
            if (win) {                                                                              // this tick's winner
                const u64_t rb = rem_bits(r);
                if (rb <= tick_b) {                                                                 // JOB:555-556
                    --to_complete;
                    const int2 row = op_row[op];
                    _Pragma("unroll 1")
                    for (int e = row.x; e < row.x + row.y; ++e) {                                   // JOB:496-506
                        const uint2 kd = dep_kd[e];
                        if (kd.y & 1u) {
                            const double rt = dep_rt[e];
                            flows.put(nF, make_int4(__double2loint(rt), __double2hiint(rt), (int)kd.x, (int)kd.y));
                            ++nF;
                        } else { nfs.put(nNF, kd.y); ++nNF; }
                    }
                    continue;
                }
                const double rem = __dsub_rn(__longlong_as_double((long long)rb), tick);
                r.x = __double2loint(rem); r.y = __double2hiint(rem);
            }
            ops.put(p, r, op); ++p;
        }
        _Pragma("unroll 1")
        for (int k = nO; k < tailO; ++k, ++p) { if (p != k) ops.put(p, ops.rec(k), ops.idx(k)); }
        nO = p;
        if (SPILL) {                          // the sizes the fast path relies on next time (TemplateHints)
            R.max_o = (tailO > R.max_o) ? tailO : R.max_o;
            R.max_f = (nF > R.max_f) ? nF : R.max_f;
            R.max_nf = (nNF > R.max_nf) ? nNF : R.max_nf;
        }
        // ---- K, L ----
        if ((to_complete == 0) | ((uint32_t)(tick_b >> 32) == 0x7FF00000u) | (R.status != RAMP_ST_OK)) {
            if (to_complete != 0 && (uint32_t)(tick_b >> 32) == 0x7FF00000u) R.status = RAMP_ST_INFINITE_TICK;       // JOB:549-551 first, then RCE:462
            break;
        }
    }
    return R;
}

__global__ void __launch_bounds__(32) ramp_lookahead_thread_kernel(const ThreadArgs a) {
    extern __shared__ __align__(128) unsigned char smem_thr[];
    __shared__ __align__(8) unsigned long long mbar;
    const int lane = threadIdx.x;
    unsigned char* st = smem_thr + a.tmpl_cap;                      // per-lane state, by decreasing alignment
    LaneCtx x;
    x.tm = smem_thr;                                                // template blob
    x.f_sm = reinterpret_cast<int4*>(st);                                                // [FCAP][32]
    x.o_sm = x.f_sm + RAMP_T_FCAP * 32;                                                  // [OCAP][32]
    x.nf_sm = reinterpret_cast<uint32_t*>(x.o_sm + RAMP_T_OCAP * 32);                    // [NFCAP][32]
    x.oi_sm = reinterpret_cast<int32_t*>(x.nf_sm + RAMP_T_NFCAP * 32);                   // [OCAP][32]
    x.wk_sm = reinterpret_cast<uint32_t*>(x.oi_sm + RAMP_T_OCAP * 32);                   // [WCAP][32]
    x.ck_sm = x.wk_sm + RAMP_T_WCAP * 32;                                                // [CCAP][32]
    x.cnt_sm = reinterpret_cast<uint16_t*>(x.ck_sm + RAMP_T_CCAP * 32);                  // [n_cap][32]
    unsigned char* slab = a.scratch + (uint64_t)blockIdx.x * a.scratch_stride;
    x.f_gl = reinterpret_cast<int4*>(slab);                                              // [spill_deps][32]
    x.o_gl = x.f_gl + (size_t)a.spill_deps * 32;                                         // [spill_ops][32]
    double* tmp_tick = reinterpret_cast<double*>(x.o_gl + (size_t)a.spill_ops * 32);     // [trace_cap][32]
    x.nf_gl = reinterpret_cast<uint32_t*>(tmp_tick + (size_t)a.trace_cap * 32);          // [spill_deps][32]
    x.oi_gl = reinterpret_cast<int32_t*>(x.nf_gl + (size_t)a.spill_deps * 32);           // [spill_ops][32]
    int32_t* tmp_n = x.oi_gl + (size_t)a.spill_ops * 32;                                 // [trace_cap][32]
    x.lane = lane; x.n_cap = a.n_cap;

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
                             ::"r"(smem_u32(smem_thr)), "l"(TD.res_blob), "r"(bytes), "r"(smem_u32(&mbar)) : "memory");
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
        // what an earlier lookahead of this template recorded: number of ticks and the largest frontiers (deterministic per
        // template).  With it the trace goes straight to an exactly-sized pool allocation and no list can leave shared memory.
        TemplateHints hint;
        { const int4 hv = __ldcg(reinterpret_cast<const int4*>(&a.hints[ch.template_id])); hint.n_ticks = hv.x; hint.max_o = hv.y; hint.max_f = hv.z; hint.max_nf = hv.w; }
        const bool fast = hint.n_ticks > 0 && hint.n_ticks <= a.trace_cap && hint.max_o <= RAMP_T_OCAP && hint.max_f <= RAMP_T_FCAP
                          && hint.max_nf <= RAMP_T_NFCAP;
        if (lane < ch.count) {
            const WorkItem item = a.items[(size_t)c * 32 + lane];
            const ResHeader& H = *reinterpret_cast<const ResHeader*>(x.tm);
            const bool simple = (H.n_workers == 1) && (H.n_channels <= 1);
            long long off = -1;
            int status0 = RAMP_ST_OK;
            const bool direct = fast && a.pool.top != nullptr;          // trace written in place
            if (direct) {
                const unsigned long long o = atomicAdd(a.pool.top, (unsigned long long)hint.n_ticks);
                if (o + (unsigned long long)hint.n_ticks <= a.pool.len) off = (long long)o; else status0 = RAMP_ST_TRACE_OVERFLOW;
            }
            if (off >= 0) { x.tr_n = a.pool.n_active + off; x.tr_tick = a.pool.tick + off; x.tr_stride = 1; x.tr_cap = hint.n_ticks; }
            else { x.tr_n = tmp_n + lane; x.tr_tick = tmp_tick + lane; x.tr_stride = 32; x.tr_cap = a.trace_cap; }
            // utilisation inside the tick loop when an earlier lookahead of the template left its completion time
            const int nmw = item.n_mounted_workers > 0 ? item.n_mounted_workers : H.orig_workers;
            double hj = 0.0;
            if (fast) hj = __ldcg(&a.hint_jct[ch.template_id]);
            x.util_jct = (hj > 0.0 && !isinf(hj)) ? hj : 0.0;
            {   // keep the conversion out of the tick loop (the compiler re-materialised it there: one I2F.F64 per tick)
                double dn = (double)nmw;
                asm volatile("" : "+d"(dn));
                x.util_dn = dn;
            }
            LaneResult R;
            if (fast) R = simple ? thread_lookahead<false, true>(x) : thread_lookahead<false, false>(x);
            else R = simple ? thread_lookahead<true, true>(x) : thread_lookahead<true, false>(x);
            int status = (R.status == RAMP_ST_OK) ? status0 : R.status;
            if (direct && status == RAMP_ST_OK && R.tick_no != hint.n_ticks) status = RAMP_ST_TRACE_OVERFLOW;   // cannot happen
            if (!fast && status == RAMP_ST_OK) {                      // every lane of the chunk writes the same values
                *reinterpret_cast<int4*>(&a.hints[ch.template_id]) = make_int4(R.tick_no, R.max_o, R.max_f, R.max_nf);
                a.hint_jct[ch.template_id] = __dmul_rn(R.t, (double)H.num_training_steps);
            }

            // ---- results (RCE:450-452) ----
            const int n_rec = R.tick_no < x.tr_cap ? R.tick_no : x.tr_cap;
            const double steps = (double)H.num_training_steps;
            const double jct = __dmul_rn(R.t, steps);
            const bool can_util = (status == RAMP_ST_OK);
            // the in-loop sum divided by the recorded completion time: valid when this lookahead found the very same one
            const bool util_done = can_util && x.util_jct != 0.0 && jct == x.util_jct;
            if (!direct && a.pool.top != nullptr) {
                const unsigned long long o = atomicAdd(a.pool.top, (unsigned long long)n_rec);
                if (o + (unsigned long long)n_rec <= a.pool.len) off = (long long)o;
                else if (status == RAMP_ST_OK) status = RAMP_ST_TRACE_OVERFLOW;
            }
            double util = util_done ? R.util : 0.0;
            if (!(util_done && (direct || off < 0))) {
                if (util_done) util = 0.0;                 // the loop below recomputes it while it copies the trace
                // RCE:830-832: util = sum over ticks, in tick order, of (n_active / n_mounted_workers) * (tick / jct).  A term
                // with n_active == 0 or tick == 0 is +0.0 (jct > 0 finite) and adding +0.0 leaves the non-negative sum as it
                // is, so those ticks are skipped (about two thirds of them); n_active / n_mounted_workers is re-used while
                // n_active repeats.  Same f64 operations on the same values for every term that can change the sum.
                const double dn = (double)nmw;
                const bool copy = (!direct) && off >= 0;
                int32_t* pn = copy ? a.pool.n_active + off : nullptr;
                double* pt = copy ? a.pool.tick + off : nullptr;
                const int32_t* sn = x.tr_n; const double* stt = x.tr_tick; const int sstr = x.tr_stride;
                const bool skip_zero = can_util && (jct > 0.0) && !isinf(jct);
                int last_n = -1;
                double last_q = 0.0;
#pragma unroll 4
                for (int k = 0; k < n_rec; ++k) {
                    const int nk = sn[(size_t)k * sstr];
                    const double tk = stt[(size_t)k * sstr];
                    if (copy) { pn[k] = nk; pt[k] = tk; }
                    if (!can_util) continue;
                    if (skip_zero && (nk == 0 || tk == 0.0)) continue;
                    if (nk != last_n) { last_n = nk; last_q = __ddiv_rn((double)nk, dn); }
                    util = __dadd_rn(util, __dmul_rn(last_q, __ddiv_rn(tk, jct)));
                }
            }
            a.res.jct[item.slot] = jct;
            a.res.comm[item.slot] = __dmul_rn(R.comm, steps);
            a.res.comp[item.slot] = __dmul_rn(R.comp, steps);
            a.res.n_ticks[item.slot] = R.tick_no;
            a.res.util[item.slot] = can_util ? util : 0.0;
            a.res.util_nmw[item.slot] = can_util ? nmw : -1;
            a.res.trace_off[item.slot] = off;
            a.res.status[item.slot] = status;
            if (a.stats) {
                atomicAdd(&a.stats->lookaheads, 1ull);
                atomicAdd(&a.stats->alg_bytes, (unsigned long long)(TD.algorithmic_bytes_static + 12ull * (unsigned long long)R.tick_no));
                atomicAdd(&a.stats->quotient_bytes, (unsigned long long)(20ull * H.n_ops + 19ull * H.n_deps + 24ull + 12ull * (unsigned long long)R.tick_no));
            }
        }
        __syncwarp();
    }
}

}  // namespace ramp
