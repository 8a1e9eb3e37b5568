// ramp_kernels.cuh -- device-side data layout and kernels of the B200-native RAMP simulator hot path.
//
// Reference semantics (cwfparsonson/ddls @ 9e0b5ba; RCE = ddls/environments/ramp_cluster/
// ramp_cluster_environment.py, JOB = ddls/demands/jobs/job.py):
//   ramp_plan_kernel       memo lookup/insert of _perform_lookahead_job_completion_time   RCE:469-518, 269-275
//   ramp_lookahead_kernel  _run_lookahead tick loop                                        RCE:379-467
//   ramp_step_kernel       step(): block/mount/register lookahead/outer event loop/stats    RCE:894-1167
//
// All simulation arithmetic is IEEE f64 without FMA contraction (compile with -fmad=false), in the
// reference's accumulation order, so results are bit-identical to CPython floats.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ramp_b200.h"

namespace ramp {

// ---------------------------------------------------------------------------------------------------
// HBM layout

// One registered lowered job.  All arrays live in one device allocation (256 B aligned segments) and are
// read-only for the kernels (read through the non-coherent path).
struct TemplateDev {
    int32_t n_ops, n_deps, n_workers, n_channels;
    int32_t num_training_steps, model_id, degree, n_src;
    int32_t canon_id;           // id of the first registered byte-identical template (exact memo key)
    int32_t trace_need;         // min(N + E + 1, trace_cap): upper bound on ticks (>= 1 op or dep completes per tick)
    const double*   op_cost;      // [N]
    const uint32_t* op_key;       // [N] unique rank key: larger wins; == argmax priority, lowest index on ties (RCE:56-66)
    const uint16_t* op_worker;    // [N]
    const uint16_t* op_n_parents; // [N]
    const int32_t*  row_ptr;      // [N+1]
    const int32_t*  dep_dst;      // [E]
    const double*   dep_run_time; // [E]
    const uint32_t* dep_key;      // [E] unique rank key (RCE:672-685)
    const uint16_t* dep_channel;  // [E]
    const uint8_t*  dep_is_flow;  // [E]
    const int32_t*  src_ops;      // [n_src] ops with in-degree 0: the initial ops_ready (JOB:474-481)
    uint64_t scratch_bytes;       // dynamic state one running lookahead of this template needs
    uint64_t algorithmic_bytes_static; // 20 N + 19 E + 24, + 12 T added per run
};

struct WorkItem {
    int32_t template_id;
    int32_t slot;               // result slot
    int32_t episode;            // -1 for standalone runs
    int32_t _pad;
};

// lookahead result slots (SoA); slot == memo hash-table position (+ B extra slots for RAMP_MEMO_OFF)
struct ResultSlots {
    double*   jct;
    double*   comm;
    double*   comp;
    int32_t*  n_ticks;
    int32_t*  status;
    int64_t*  trace_off;        // offset into the trace pool, -1 if none
};

struct TracePool {
    int32_t* n_active;          // [pool_len]
    double*  tick;              // [pool_len]
    unsigned long long* top;    // bump allocator (entries)
    uint64_t len;
};

struct Counters {               // zeroed at the start of every step
    int32_t n_work;
    int32_t work_cursor;
    int32_t err_episode;        // first episode that recorded an error (+1), 0 if none
    int32_t err_status;
};

struct MemoStats { unsigned long long lookups, hits, lookaheads, alg_bytes; };

// running-job table fields (SoA: [field][row][episode])
enum { RF_JCT = 0, RF_STARTED, RF_COMM, RF_COMP, RF_UTIL, RF_PART_OP_MEM, RF_PART_DEP, RF_FLOW, RF_ORIG_OP_MEM,
       RF_ORIG_DEP, RF_COUNT };
enum { RI_JOB_IDX = 0, RI_N_WORKERS, RI_N_CHANNELS, RI_COUNT };

// per-episode scalars (SoA: [field][episode])
enum { EF_NOW = 0, EF_NEXT_ARRIVAL, EF_LAST_ARRIVAL, EF_LOAD_SUM, EF_COUNT };
enum { EI_NUM_ARRIVED = 0, EI_NUM_COMPLETED, EI_NUM_BLOCKED, EI_QUEUED, EI_N_RUNNING, EI_STEP_COUNTER, EI_EVENT_SEQ,
       EI_LOAD_N, EI_STATUS, EI_DONE, EI_LAST_SLOT, EI_PLAN_SLOT, EI_PLAN_RAN, EI_COUNT };

struct EpisodeState {
    int32_t B, max_running, max_jobs, n_jobs;
    int32_t n_cluster_workers, queue_capacity;
    double eps, max_sim_time;
    double*  ef;                // [EF_COUNT][B]
    int32_t* ei;                // [EI_COUNT][B]
    double*  rf;                // [RF_COUNT][max_running][B]
    int32_t* ri;                // [RI_COUNT][max_running][B]
    ramp_job_record_t* rec;     // [B][max_jobs]
    const ramp_arrival_t* arr;  // [B][max_jobs]
};

struct MemoTable {
    unsigned long long* keys;   // [cap] 0 = empty
    uint32_t mask;              // cap - 1
    int32_t mode;
};

struct LookaheadArgs {
    const TemplateDev* templates;
    const WorkItem* items;
    const int32_t* n_work;      // device-side count
    int32_t* cursor;            // device-side work cursor (persistent CTAs pull items)
    unsigned char* scratch;     // [gridDim.x][scratch_stride]
    uint64_t scratch_stride;
    ResultSlots res;
    TracePool pool;
    int32_t trace_cap;          // per-CTA temp trace capacity
    int32_t w_cap, c_cap;       // shared-memory key array capacities
    MemoStats* stats;
};

__host__ __device__ inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// dynamic state of one running lookahead, carved out of the CTA's scratch slab
struct ScratchView {
    double*   op_rem;           // [N] remaining_run_time of ready/ticking ops (JOB:555), written when an op becomes ready
    double*   dep_rem;          // [E] remaining_run_time of ready deps (JOB:561), written when a dep becomes ready
    uint32_t* par_done;         // [N] len(parent_deps_completed) JOB:530
    int32_t*  ops_list[2];      // ops_ready frontier (ping-pong)
    int32_t*  deps_list[2];     // deps_ready frontier (ping-pong)
    int32_t*  tr_n;             // [trace_cap] temp trace
    double*   tr_tick;          // [trace_cap]
};

__host__ __device__ inline uint64_t scratch_bytes_for(int32_t N, int32_t E) {
    uint64_t b = 0;
    b += align_up((uint64_t)N * 8, 16);
    b += align_up((uint64_t)E * 8, 16);
    b += align_up((uint64_t)N * 4, 16);
    b += 2 * align_up((uint64_t)N * 4, 16);
    b += 2 * align_up((uint64_t)E * 4, 16);
    return b;
}

__device__ inline ScratchView carve(unsigned char* base, int32_t N, int32_t E, uint64_t trace_region_off) {
    ScratchView v;
    uint64_t o = 0;
    v.op_rem = (double*)(base + o);        o += align_up((uint64_t)N * 8, 16);
    v.dep_rem = (double*)(base + o);       o += align_up((uint64_t)E * 8, 16);
    v.par_done = (uint32_t*)(base + o);    o += align_up((uint64_t)N * 4, 16);
    v.ops_list[0] = (int32_t*)(base + o);  o += align_up((uint64_t)N * 4, 16);
    v.ops_list[1] = (int32_t*)(base + o);  o += align_up((uint64_t)N * 4, 16);
    v.deps_list[0] = (int32_t*)(base + o); o += align_up((uint64_t)E * 4, 16);
    v.deps_list[1] = (int32_t*)(base + o); o += align_up((uint64_t)E * 4, 16);
    v.tr_tick = (double*)(base + trace_region_off);
    return v;
}

// ---------------------------------------------------------------------------------------------------
// small device helpers

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// Python: x -= min(tick, x) (JOB:555, JOB:561); min(a, b) returns a unless b < a.
__device__ __forceinline__ double tick_down(double rem, double tick) {
    const double m = (rem < tick) ? rem : tick;
    return __dsub_rn(rem, m);
}

// Warp-aggregated append to a frontier list: one shared-memory atomic per warp.  Must be called by all
// 32 lanes of the warp convergently.
__device__ __forceinline__ void warp_push(int32_t* list, int* counter, bool pred, int32_t val) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = val;
}

__device__ __forceinline__ double warp_min_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, v, o);
        v = (other < v) ? other : v;
    }
    return v;
}

__device__ __forceinline__ int warp_sum_i32(int v) {
    return __reduce_add_sync(0xffffffffu, v);
}

#define RAMP_INF_BITS 0x7FF0000000000000ull

// ---------------------------------------------------------------------------------------------------
// _run_lookahead (RCE:379-467): one CTA per lookahead, persistent CTAs pull work items.
//
// Per tick (letters as in SURVEY.md 3.3):
//   A  per-worker arg-max over ready ops         -> atomicMax of unique rank keys into smem wkey[]
//   C  any ready non-flow dep?                    -> __syncthreads_or
//   D  per-channel arg-max over ready deps        -> atomicMax into smem ckey[]   (skipped if C)
//   B/D min remaining over the winners           -> warp shuffles + one smem atomicMin per warp (u64 bit pattern
//                                                   of non-negative doubles is order preserving)
//   E  tick = min(t_op, t_comm)
//   G  winners: rem -= min(tick, rem); == 0 -> completed, out-edges appended to the next dep frontier
//   H  deps of the pre-tick snapshot: same; completed -> atomicAdd on the child's parent counter, == n_parents
//      -> child appended to the next op frontier
//   I,J thread 0 accumulates t / comm / comp and the trace in tick order
template <int NT>
__global__ void __launch_bounds__(NT) ramp_lookahead_kernel(const LookaheadArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* wkey = reinterpret_cast<uint32_t*>(smem_raw);        // [w_cap] best rank key per worker this tick
    uint32_t* ckey = wkey + a.w_cap;                               // [c_cap] best rank key per channel this tick
    int32_t* doneq = reinterpret_cast<int32_t*>(ckey + a.c_cap);   // [w_cap] ops completed this tick (<= 1 per worker)

    __shared__ int s_work;
    __shared__ int s_n_ops[2], s_n_deps[2];
    __shared__ int s_doneq_n;
    __shared__ int s_ops_completed, s_deps_completed;
    __shared__ unsigned long long s_min_op[2], s_min_dep[2];
    __shared__ int s_n_active[2];
    __shared__ int s_stop;                                         // 0 continue, 1 finished, 2 error

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    constexpr int NW = NT / 32;

    unsigned char* slab = a.scratch + (uint64_t)blockIdx.x * a.scratch_stride;
    const uint64_t trace_region = a.scratch_stride - align_up((uint64_t)a.trace_cap * 12, 16);

    for (;;) {
        if (tid == 0) s_work = atomicAdd(a.cursor, 1);
        __syncthreads();
        const int wi = s_work;
        if (wi >= *a.n_work) break;
        const WorkItem item = a.items[wi];
        const TemplateDev& T = a.templates[item.template_id];
        const int N = T.n_ops, E = T.n_deps, W = T.n_workers, C = T.n_channels;
        ScratchView sv = carve(slab, N, E, trace_region);
        sv.tr_n = reinterpret_cast<int32_t*>(sv.tr_tick + a.trace_cap);

        const double* __restrict__ op_cost = T.op_cost;
        const uint32_t* __restrict__ op_key = T.op_key;
        const uint16_t* __restrict__ op_worker = T.op_worker;
        const uint16_t* __restrict__ op_n_parents = T.op_n_parents;
        const int32_t* __restrict__ row_ptr = T.row_ptr;
        const int32_t* __restrict__ dep_dst = T.dep_dst;
        const double* __restrict__ dep_run_time = T.dep_run_time;
        const uint32_t* __restrict__ dep_key = T.dep_key;
        const uint16_t* __restrict__ dep_channel = T.dep_channel;
        const uint8_t* __restrict__ dep_is_flow = T.dep_is_flow;

        // ---- init (JOB:432-484) ----
        for (int i = tid; i < N; i += NT) sv.par_done[i] = 0u;
        for (int i = tid; i < W; i += NT) wkey[i] = 0u;
        for (int i = tid; i < C; i += NT) ckey[i] = 0u;
        for (int k = tid; k < T.n_src; k += NT) {
            const int op = __ldg(&T.src_ops[k]);
            sv.ops_list[0][k] = op;
            sv.op_rem[op] = __ldg(&op_cost[op]);                  // RCE:1334
        }
        if (tid == 0) {
            s_n_ops[0] = T.n_src; s_n_ops[1] = 0; s_n_deps[0] = 0; s_n_deps[1] = 0;
            s_doneq_n = 0; s_ops_completed = 0; s_deps_completed = 0;
            s_min_op[0] = s_min_op[1] = RAMP_INF_BITS; s_min_dep[0] = s_min_dep[1] = RAMP_INF_BITS;
            s_n_active[0] = s_n_active[1] = 0;
            s_stop = 0;
        }
        __syncthreads();

        // thread-0 private accumulators (Stopwatch UT:485-496, JOB:170-171)
        double t = 0.0, comm = 0.0, comp = 0.0;
        int tick_no = 0;
        int status = RAMP_ST_OK;
        int cur = 0;

        for (;;) {
            const int nxt = cur ^ 1;
            const int nO = s_n_ops[cur], nD = s_n_deps[cur];
            const int32_t* ops = sv.ops_list[cur];      // written in other ticks: coherent loads
            const int32_t* deps = sv.deps_list[cur];
            int32_t* ops_n = sv.ops_list[nxt];
            int32_t* deps_n = sv.deps_list[nxt];

            // ---- A: highest-priority ready op per worker (RCE:562-590, 44-67) ----
            for (int k = tid; k < nO; k += NT) {
                const int i = ops[k];
                atomicMax(&wkey[__ldg(&op_worker[i])], __ldg(&op_key[i]));
            }
            // ---- C: any ready non-flow dep? (RCE:520-540) ----
            int nf = 0;
            for (int k = tid; k < nD; k += NT) nf |= (__ldg(&dep_is_flow[deps[k]]) == 0);
            const int any_nf = __syncthreads_or(nf);

            // ---- D: highest-priority ready dep per channel (RCE:608-629, 665-689) ----
            if (!any_nf) {
                for (int k = tid; k < nD; k += NT) {
                    const int e = deps[k];
                    const uint32_t c = __ldg(&dep_channel[e]);
                    if (c != RAMP_NO_CHANNEL) atomicMax(&ckey[c], __ldg(&dep_key[e]));
                }
                __syncthreads();
            }

            // ---- B / D.iii: shortest remaining time over the winners (RCE:592-606, 653-663) ----
            {
                double mo = __longlong_as_double(RAMP_INF_BITS), md = mo;
                int na = 0;
                for (int k = tid; k < nO; k += NT) {
                    const int i = ops[k];
                    if (wkey[__ldg(&op_worker[i])] == __ldg(&op_key[i])) {
                        ++na;
                        const double r = sv.op_rem[i];
                        mo = (r < mo) ? r : mo;
                    }
                }
                if (!any_nf) {
                    for (int k = tid; k < nD; k += NT) {
                        const int e = deps[k];
                        const uint32_t c = __ldg(&dep_channel[e]);
                        if (c != RAMP_NO_CHANNEL && ckey[c] == __ldg(&dep_key[e])) {
                            const double r = sv.dep_rem[e];
                            md = (r < md) ? r : md;
                        }
                    }
                }
                mo = warp_min_f64(mo);
                md = warp_min_f64(md);
                na = warp_sum_i32(na);
                if (lane == 0) {
                    const unsigned long long bo = (unsigned long long)__double_as_longlong(mo);
                    const unsigned long long bd = (unsigned long long)__double_as_longlong(md);
                    if (bo != RAMP_INF_BITS) atomicMin(&s_min_op[cur], bo);
                    if (bd != RAMP_INF_BITS) atomicMin(&s_min_dep[cur], bd);
                    if (na) atomicAdd(&s_n_active[cur], na);
                }
            }
            __syncthreads();

            // ---- E: tick (RCE:426) ----
            const double t_op = __longlong_as_double((long long)s_min_op[cur]);
            const double t_comm = any_nf ? 0.0 : __longlong_as_double((long long)s_min_dep[cur]);
            const double tick = (t_comm < t_op) ? t_comm : t_op;
            const int n_active = s_n_active[cur];

            // ---- G: tick the winners (RCE:691-716, JOB:553-557, 492-501) ----
            int done_local = 0;
            for (int kb = 0; kb < nO; kb += NT) {
                const int k = kb + tid;
                const bool valid = k < nO;
                int i = 0;
                bool done = false;
                if (valid) {
                    i = ops[k];
                    if (wkey[__ldg(&op_worker[i])] == __ldg(&op_key[i])) {
                        const double r = tick_down(sv.op_rem[i], tick);
                        if (r == 0.0) done = true; else sv.op_rem[i] = r;
                    }
                }
                warp_push(ops_n, &s_n_ops[nxt], valid && !done, i);    // still ready next tick
                warp_push(doneq, &s_doneq_n, done, i);                  // completed: expand out-edges below
                done_local += done ? 1 : 0;
            }
            // ---- H: tick the deps of the pre-tick snapshot (RCE:718-775, JOB:559-563, 525-536) ----
            int ddone_local = 0;
            for (int kb = 0; kb < nD; kb += NT) {
                const int k = kb + tid;
                const bool valid = k < nD;
                int e = 0, child = 0;
                bool done = false, readied = false;
                if (valid) {
                    e = deps[k];
                    const bool ticked = !(any_nf && __ldg(&dep_is_flow[e]));     // RCE:434-439
                    if (!any_nf) {                                               // release this tick's channel winner slot
                        const uint32_t c = __ldg(&dep_channel[e]);
                        if (c != RAMP_NO_CHANNEL) ckey[c] = 0u;
                    }
                    if (ticked) {
                        const double r = tick_down(sv.dep_rem[e], tick);
                        if (r == 0.0) {
                            done = true;
                            child = __ldg(&dep_dst[e]);
                            const uint32_t cnt = atomicAdd(&sv.par_done[child], 1u) + 1u;   // JOB:530
                            readied = (cnt == (uint32_t)__ldg(&op_n_parents[child]));        // JOB:531 (fires once)
                        } else {
                            sv.dep_rem[e] = r;
                        }
                    }
                }
                warp_push(deps_n, &s_n_deps[nxt], valid && !done, e);
                if (readied) sv.op_rem[child] = __ldg(&op_cost[child]);
                warp_push(ops_n, &s_n_ops[nxt], readied, child);
                ddone_local += done ? 1 : 0;
            }
            done_local = warp_sum_i32(done_local);
            ddone_local = warp_sum_i32(ddone_local);
            if (lane == 0) {
                if (done_local) atomicAdd(&s_ops_completed, done_local);
                if (ddone_local) atomicAdd(&s_deps_completed, ddone_local);
            }
            __syncthreads();

            // ---- G (cont.): out-edges of completed ops become ready next tick (JOB:496-506), one warp per op,
            //      contiguous dep indices -> coalesced writes ----
            {
                const int nq = s_doneq_n;
                for (int q = warp; q < nq; q += NW) {
                    const int i = doneq[q];
                    const int start = __ldg(&row_ptr[i]), deg = __ldg(&row_ptr[i + 1]) - start;
                    int base = 0;
                    if (lane == 0 && deg > 0) base = atomicAdd(&s_n_deps[nxt], deg);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    for (int j = lane; j < deg; j += 32) {
                        deps_n[base + j] = start + j;
                        sv.dep_rem[start + j] = __ldg(&dep_run_time[start + j]);     // RCE:542-560
                    }
                }
                for (int i = tid; i < W; i += NT) wkey[i] = 0u;
            }
            // ---- I, J, K, L: serial bookkeeping in tick order (RCE:442-465, 777-791) ----
            if (tid == 0) {
                const bool ticked_ops = n_active > 0;
                const bool ticked_flows = (!any_nf) && (nD > 0);
                if (ticked_ops && ticked_flows) { comm = __dadd_rn(comm, tick); comp = __dadd_rn(comp, tick); }
                else if (ticked_flows) comm = __dadd_rn(comm, tick);
                else if (ticked_ops) comp = __dadd_rn(comp, tick);
                t = __dadd_rn(t, tick);
                if (tick_no < a.trace_cap) { sv.tr_n[tick_no] = n_active; sv.tr_tick[tick_no] = tick; }
                else status = RAMP_ST_TRACE_OVERFLOW;
                ++tick_no;
                // reset this tick's reduction cells and frontier counters for their next use
                s_min_op[nxt] = RAMP_INF_BITS; s_min_dep[nxt] = RAMP_INF_BITS; s_n_active[nxt] = 0;
                s_min_op[cur] = RAMP_INF_BITS; s_min_dep[cur] = RAMP_INF_BITS; s_n_active[cur] = 0;
                s_n_ops[cur] = 0; s_n_deps[cur] = 0;
                if (s_ops_completed == N && s_deps_completed == E) s_stop = 1;                    // JOB:549-551
                else if (isinf(tick)) { s_stop = 2; status = RAMP_ST_INFINITE_TICK; }              // RCE:462
            }
            __syncthreads();
            if (tid == 0) s_doneq_n = 0;
            cur = nxt;
            if (s_stop) break;
        }

        // ---- results (RCE:450-452): copy the trace to an exactly-sized pool allocation ----
        __shared__ long long s_trace_off;
        if (tid == 0) {
            const double steps = (double)T.num_training_steps;
            a.res.jct[item.slot] = __dmul_rn(t, steps);
            a.res.comm[item.slot] = __dmul_rn(comm, steps);
            a.res.comp[item.slot] = __dmul_rn(comp, steps);
            a.res.n_ticks[item.slot] = tick_no;
            const int n_rec = tick_no < a.trace_cap ? tick_no : a.trace_cap;
            long long off = -1;
            if (a.pool.top != nullptr) {
                const unsigned long long o = atomicAdd(a.pool.top, (unsigned long long)n_rec);
                if (o + (unsigned long long)n_rec <= a.pool.len) off = (long long)o;
                else if (status == RAMP_ST_OK) status = RAMP_ST_TRACE_OVERFLOW;
            }
            a.res.trace_off[item.slot] = off;
            a.res.status[item.slot] = status;
            s_trace_off = off;
            s_n_ops[0] = n_rec;
            if (a.stats) {
                atomicAdd(&a.stats->lookaheads, 1ull);
                atomicAdd(&a.stats->alg_bytes, (unsigned long long)(T.algorithmic_bytes_static + 12ull * (unsigned long long)tick_no));
            }
        }
        __syncthreads();
        if (s_trace_off >= 0) {
            const int n_rec = s_n_ops[0];
            for (int k = tid; k < n_rec; k += NT) {
                a.pool.n_active[s_trace_off + k] = sv.tr_n[k];
                a.pool.tick[s_trace_off + k] = sv.tr_tick[k];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// memo lookup / insert + work-list construction (RCE:469-518)

struct PlanArgs {
    const ramp_action_t* actions;   // [B]
    const TemplateDev* templates;
    int32_t n_templates;
    EpisodeState ep;
    MemoTable memo;
    WorkItem* items;                // [B]
    Counters* counters;
    MemoStats* stats;
};

__global__ void ramp_plan_kernel(const PlanArgs p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int B = p.ep.B;
    if (b >= B) return;
    int32_t* ei = p.ep.ei;
    ei[EI_PLAN_SLOT * B + b] = -1;
    ei[EI_PLAN_RAN * B + b] = 0;
    const ramp_action_t act = p.actions[b];
    if ((act.flags & RAMP_ACT_SKIP) || ei[EI_DONE * B + b]) return;
    if (act.template_id < 0 || act.template_id >= p.n_templates) return;
    if (ei[EI_QUEUED * B + b] < 0) return;                              // reported by the step kernel
    const TemplateDev& T = p.templates[act.template_id];
    const uint32_t cap_mask = p.memo.mask;
    int slot = -1;
    bool ran = false;
    if (p.memo.mode == RAMP_MEMO_OFF) {
        slot = (int)(cap_mask + 1u) + b;
        ran = true;
    } else {
        unsigned long long key;
        if (p.memo.mode == RAMP_MEMO_REFERENCE)                           // [model][max_num_partitions] per env instance RCE:491-492
            key = ((unsigned long long)(b + 1) << 32) | ((unsigned long long)(T.model_id & 0xFFFF) << 16)
                  | (unsigned long long)(T.degree & 0xFFFF);
        else
            key = 0x8000000000000000ull | (unsigned long long)(T.canon_id + 1);
        uint32_t pos = (uint32_t)splitmix64(key) & cap_mask;
        for (uint32_t probe = 0; probe <= cap_mask; ++probe) {
            const unsigned long long old = atomicCAS(&p.memo.keys[pos], 0ull, key);
            if (old == 0ull) { slot = (int)pos; ran = true; break; }      // miss: this episode runs the lookahead RCE:502-506
            if (old == key) { slot = (int)pos; break; }                   // hit RCE:495-498
            pos = (pos + 1u) & cap_mask;
        }
        atomicAdd(&p.stats->lookups, 1ull);
        if (slot >= 0 && !ran) atomicAdd(&p.stats->hits, 1ull);
    }
    if (slot < 0) {                                                       // table full
        atomicCAS(&p.counters->err_episode, 0, b + 1);
        ei[EI_STATUS * B + b] = RAMP_ST_TABLE_FULL;
        return;
    }
    ei[EI_PLAN_SLOT * B + b] = slot;
    ei[EI_PLAN_RAN * B + b] = ran ? 1 : 0;
    if (ran) {
        const int w = atomicAdd(&p.counters->n_work, 1);
        WorkItem it; it.template_id = act.template_id; it.slot = slot; it.episode = b; it._pad = 0;
        p.items[w] = it;
    }
}

// ---------------------------------------------------------------------------------------------------
// RampClusterEnvironment.step (RCE:894-1179): one thread per episode; serial f64 in the reference's order.

struct StepArgs {
    const ramp_action_t* actions;   // [B]
    EpisodeState ep;
    ResultSlots res;
    TracePool pool;
    Counters* counters;
    double* stats_out;              // [B][RAMP_STEP_STATS_LEN] or null
    int32_t* n_cluster_steps_out;   // [B] or null
    int32_t fuse_empty_steps;
};

#define EF(f) ef[(f) * B + b]
#define EI(f) ei[(f) * B + b]
#define RF(f, row) rf[((f) * R + (row)) * B + b]
#define RI(f, row) ri[((f) * R + (row)) * B + b]

__device__ inline void step_register_blocked(const EpisodeState& ep, int b, int job_idx, double* st) {   // RCE:1504-1540
    const int B = ep.B;
    int32_t* ei = ep.ei;
    ramp_job_record_t& r = ep.rec[(size_t)b * ep.max_jobs + job_idx];
    if (EI(EI_QUEUED) == job_idx) EI(EI_QUEUED) = -1;
    if (r.status == RAMP_JS_BLOCKED) return;
    r.status = RAMP_JS_BLOCKED;
    r.event_seq = EI(EI_EVENT_SEQ)++;
    EI(EI_NUM_BLOCKED)++;
    st[RAMP_SS_NUM_JOBS_BLOCKED] += 1.0;
}

__device__ inline void step_remove_running(const EpisodeState& ep, int b, int pos) {   // keeps dict (insertion) order
    const int B = ep.B, R = ep.max_running;
    double* rf = ep.rf; int32_t* ri = ep.ri; int32_t* ei = ep.ei;
    const int n = EI(EI_N_RUNNING);
    for (int k = pos; k + 1 < n; ++k) {
        for (int f = 0; f < RF_COUNT; ++f) RF(f, k) = RF(f, k + 1);
        for (int f = 0; f < RI_COUNT; ++f) RI(f, k) = RI(f, k + 1);
    }
    EI(EI_N_RUNNING) = n - 1;
}

__device__ inline bool step_is_done(const EpisodeState& ep, int b) {   // RCE:1542-1557
    const int B = ep.B;
    const double* ef = ep.ef; const int32_t* ei = ep.ei;
    if (EF(EF_NOW) >= ep.max_sim_time) return true;
    return (ep.n_jobs - EI(EI_NUM_ARRIVED)) == 0 && EI(EI_N_RUNNING) == 0 && EI(EI_QUEUED) < 0;
}

__device__ inline void step_get_next_job(const EpisodeState& ep, int b) {   // RCE:351-377
    const int B = ep.B;
    double* ef = ep.ef; int32_t* ei = ep.ei;
    const int k = EI(EI_NUM_ARRIVED);
    ramp_job_record_t& r = ep.rec[(size_t)b * ep.max_jobs + k];
    r.status = RAMP_JS_QUEUED; r.event_seq = 0;
    r.time_arrived = EF(EF_NOW); r.time_started = 0.0; r.time_completed = 0.0;
    r.jct = r.comm = r.comp = r.util = 0.0;
    const ramp_arrival_t a = ep.arr[(size_t)b * ep.max_jobs + k];
    EF(EF_LAST_ARRIVAL) = EF(EF_NOW);                                        // RCE:362
    EF(EF_NEXT_ARRIVAL) = __dadd_rn(EF(EF_NEXT_ARRIVAL), a.interarrival);    // RCE:363
    EF(EF_LOAD_SUM) = __dadd_rn(EF(EF_LOAD_SUM),
                                __ddiv_rn(__dadd_rn(a.orig_op_mem, a.orig_dep_size),
                                          __dsub_rn(EF(EF_NEXT_ARRIVAL), EF(EF_LAST_ARRIVAL))));   // RCE:364
    EI(EI_LOAD_N)++;
    EI(EI_NUM_ARRIVED) = k + 1;
}

__global__ void ramp_step_kernel(const StepArgs s) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const EpisodeState& ep = s.ep;
    const int B = ep.B, R = ep.max_running;
    if (b >= B) return;
    double* ef = ep.ef; int32_t* ei = ep.ei; double* rf = ep.rf; int32_t* ri = ep.ri;
    const ramp_action_t act = s.actions[b];

    double st[RAMP_STEP_STATS_LEN];
    double st0[RAMP_STEP_STATS_LEN];
#pragma unroll
    for (int k = 0; k < RAMP_STEP_STATS_LEN; ++k) { st[k] = 0.0; st0[k] = 0.0; }
    int n_cluster_steps = 0;

    if (!((act.flags & RAMP_ACT_SKIP) || EI(EI_DONE))) {
        for (int cs = 0;; ++cs) {
#pragma unroll
            for (int k = 0; k < RAMP_STEP_STATS_LEN; ++k) st[k] = 0.0;
            st[RAMP_SS_STEP_COUNTER] = (double)EI(EI_STEP_COUNTER);            // RCE:309
            st[RAMP_SS_STEP_START_TIME] = EF(EF_NOW);                          // RCE:310
            bool has_action = (cs == 0) && act.template_id >= 0;
            int handled = -1;
            if (has_action) {
                handled = EI(EI_QUEUED);
                if (handled < 0 || EI(EI_PLAN_SLOT) < 0) {
                    if (handled < 0) { EI(EI_STATUS) = RAMP_ST_NO_QUEUED_JOB; atomicCAS(&s.counters->err_episode, 0, b + 1); }
                    has_action = false;
                }
            }
            // RCE:914-919: queued jobs not handled by the action are blocked
            if (!has_action && EI(EI_QUEUED) >= 0) step_register_blocked(ep, b, EI(EI_QUEUED), st);

            if (has_action) {
                const int slot = EI(EI_PLAN_SLOT);
                EI(EI_LAST_SLOT) = slot;
                st[RAMP_SS_LOOKAHEAD_RAN] = (double)EI(EI_PLAN_RAN);
                ramp_job_record_t& r = ep.rec[(size_t)b * ep.max_jobs + handled];
                r.status = RAMP_JS_RUNNING;
                r.time_started = EF(EF_NOW);                                    // RCE:1418
                EI(EI_QUEUED) = -1;                                             // RCE:1420
                const int lst = s.res.status[slot];
                if (lst != RAMP_ST_OK) {                                        // the reference raises (RCE:462)
                    EI(EI_STATUS) = lst; atomicCAS(&s.counters->err_episode, 0, b + 1);
                    step_register_blocked(ep, b, handled, st);
                } else {
                    const double jct = s.res.jct[slot];
                    if (jct > act.max_acceptable_jct) {                         // RCE:815 (strict)
                        step_register_blocked(ep, b, handled, st);              // RCE:821-824
                    } else if (EI(EI_N_RUNNING) >= R) {
                        EI(EI_STATUS) = RAMP_ST_TABLE_FULL; atomicCAS(&s.counters->err_episode, 0, b + 1);
                        step_register_blocked(ep, b, handled, st);
                    } else {
                        // RCE:830-832 serial sum in tick order
                        double util = 0.0;
                        const long long off = s.res.trace_off[slot];
                        const int T = s.res.n_ticks[slot];
                        const double nmw = (double)act.n_mounted_workers;
                        if (off >= 0) {
                            for (int k = 0; k < T; ++k)
                                util = __dadd_rn(util, __dmul_rn(__ddiv_rn((double)s.pool.n_active[off + k], nmw),
                                                                 __ddiv_rn(s.pool.tick[off + k], jct)));
                        }
                        const int row = EI(EI_N_RUNNING)++;
                        const ramp_arrival_t arr = ep.arr[(size_t)b * ep.max_jobs + handled];
                        RF(RF_JCT, row) = jct; RF(RF_STARTED, row) = EF(EF_NOW);
                        RF(RF_COMM, row) = s.res.comm[slot]; RF(RF_COMP, row) = s.res.comp[slot]; RF(RF_UTIL, row) = util;
                        RF(RF_PART_OP_MEM, row) = act.part_op_mem; RF(RF_PART_DEP, row) = act.part_dep_size;
                        RF(RF_FLOW, row) = act.flow_size;
                        RF(RF_ORIG_OP_MEM, row) = arr.orig_op_mem; RF(RF_ORIG_DEP, row) = arr.orig_dep_size;
                        RI(RI_JOB_IDX, row) = handled; RI(RI_N_WORKERS, row) = act.n_mounted_workers;
                        RI(RI_N_CHANNELS, row) = act.n_mounted_channels;
                        r.jct = jct; r.comm = s.res.comm[slot]; r.comp = s.res.comp[slot]; r.util = util;
                    }
                }
            }

            // ---- outer event loop RCE:942-1044 ----
            double util_mounted_sum = 0.0, util_cluster_sum = 0.0;
            double sum_jobs = 0.0, sum_workers = 0.0, sum_channels = 0.0, sum_comp_frac = 0.0, sum_comm_frac = 0.0;
            int n_frac = 0, n_iter = 0;
            bool step_done = false;
            while (!step_done) {
                const double now = EF(EF_NOW);
                double tick = __dsub_rn(EF(EF_NEXT_ARRIVAL), now);                           // RCE:950
                { const double b2 = __dsub_rn(ep.max_sim_time, now); if (b2 < tick) tick = b2; }
                const int nr = EI(EI_N_RUNNING);
                for (int k = 0; k < nr; ++k) {                                               // RCE:951-954
                    const double remaining = __dsub_rn(RF(RF_JCT, k), __dsub_rn(now, RF(RF_STARTED, k)));
                    if (remaining < tick) tick = remaining;
                }
                int mounted_workers = 0, mounted_channels = 0;
                double util_sum = 0.0;
                for (int k = 0; k < nr; ++k) {                                               // RCE:962-982
                    const double jct = RF(RF_JCT, k);
                    const double frac = __ddiv_rn(tick, jct);
                    const double pom = RF(RF_PART_OP_MEM, k), pds = RF(RF_PART_DEP, k);
                    const double oom = RF(RF_ORIG_OP_MEM, k), ods = RF(RF_ORIG_DEP, k);
                    st[RAMP_SS_COMPUTE_INFO_PROCESSED] = __dadd_rn(st[RAMP_SS_COMPUTE_INFO_PROCESSED], __dmul_rn(pom, frac));
                    st[RAMP_SS_DEP_INFO_PROCESSED] = __dadd_rn(st[RAMP_SS_DEP_INFO_PROCESSED], __dmul_rn(pds, frac));
                    st[RAMP_SS_FLOW_INFO_PROCESSED] = __dadd_rn(st[RAMP_SS_FLOW_INFO_PROCESSED], __dmul_rn(RF(RF_FLOW, k), frac));
                    st[RAMP_SS_CLUSTER_INFO_PROCESSED] = __dadd_rn(st[RAMP_SS_CLUSTER_INFO_PROCESSED], __dmul_rn(__dadd_rn(pom, pds), frac));
                    st[RAMP_SS_DEMAND_COMPUTE_INFO_PROCESSED] = __dadd_rn(st[RAMP_SS_DEMAND_COMPUTE_INFO_PROCESSED], __dmul_rn(oom, frac));
                    st[RAMP_SS_DEMAND_DEP_INFO_PROCESSED] = __dadd_rn(st[RAMP_SS_DEMAND_DEP_INFO_PROCESSED], __dmul_rn(ods, frac));
                    st[RAMP_SS_DEMAND_TOTAL_INFO_PROCESSED] = __dadd_rn(st[RAMP_SS_DEMAND_TOTAL_INFO_PROCESSED], __dmul_rn(__dadd_rn(oom, ods), frac));
                    sum_comp_frac = __dadd_rn(sum_comp_frac, __ddiv_rn(RF(RF_COMP, k), jct));
                    sum_comm_frac = __dadd_rn(sum_comm_frac, __ddiv_rn(RF(RF_COMM, k), jct));
                    ++n_frac;
                    mounted_workers += RI(RI_N_WORKERS, k);      // workers / channels of distinct jobs are disjoint (ramp_rules.py:1-40)
                    mounted_channels += RI(RI_N_CHANNELS, k);
                    util_sum = __dadd_rn(util_sum, RF(RF_UTIL, k));
                }
                sum_jobs = __dadd_rn(sum_jobs, (double)nr);                                   // RCE:984
                sum_workers = __dadd_rn(sum_workers, (double)mounted_workers);               // RCE:986
                sum_channels = __dadd_rn(sum_channels, (double)mounted_channels);            // RCE:987
                if (nr > 0) {                                                                // RCE:989-994
                    const double mean_util = __ddiv_rn(util_sum, (double)nr);
                    util_mounted_sum = __dadd_rn(util_mounted_sum, mean_util);
                    util_cluster_sum = __dadd_rn(util_cluster_sum,
                                                 __dmul_rn(__ddiv_rn((double)mounted_workers, (double)ep.n_cluster_workers), mean_util));
                }
                ++n_iter;
                EF(EF_NOW) = __dadd_rn(now, tick);                                            // RCE:998
                const double now2 = EF(EF_NOW);

                // RCE:1004-1017, 1466-1502
                int k = 0;
                while (k < EI(EI_N_RUNNING)) {
                    const double remaining = __dsub_rn(__dsub_rn(RF(RF_JCT, k), __dsub_rn(now2, RF(RF_STARTED, k))), ep.eps);
                    if (remaining <= 0.0) {
                        ramp_job_record_t& r = ep.rec[(size_t)b * ep.max_jobs + RI(RI_JOB_IDX, k)];
                        r.status = RAMP_JS_COMPLETED; r.time_completed = now2;
                        r.event_seq = EI(EI_EVENT_SEQ)++;
                        EI(EI_NUM_COMPLETED)++;
                        st[RAMP_SS_NUM_JOBS_COMPLETED] += 1.0;
                        step_remove_running(ep, b, k);
                        step_done = true;
                    } else {
                        ++k;
                    }
                }
                // RCE:1019-1040
                if ((ep.n_jobs - EI(EI_NUM_ARRIVED)) > 0) {
                    if (__dadd_rn(now2, ep.eps) >= EF(EF_NEXT_ARRIVAL)) {
                        const int idx = EI(EI_NUM_ARRIVED);
                        step_get_next_job(ep, b);
                        st[RAMP_SS_NUM_JOBS_ARRIVED] += 1.0;
                        if (EI(EI_QUEUED) < 0 && ep.queue_capacity >= 1) EI(EI_QUEUED) = idx;   // RCE:1030-1031
                        else step_register_blocked(ep, b, idx, st);                             // RCE:1034
                        step_done = true;
                    }
                } else {
                    EF(EF_NEXT_ARRIVAL) = __longlong_as_double(RAMP_INF_BITS);                  // RCE:1040
                }
                if (step_is_done(ep, b)) step_done = true;                                       // RCE:1043
            }

            // ---- RCE:1046-1084 ----
            st[RAMP_SS_STEP_END_TIME] = EF(EF_NOW);
            st[RAMP_SS_STEP_TIME] = __dsub_rn(st[RAMP_SS_STEP_END_TIME], st[RAMP_SS_STEP_START_TIME]);
            st[RAMP_SS_MEAN_NUM_JOBS_RUNNING] = __ddiv_rn(sum_jobs, (double)n_iter);
            st[RAMP_SS_MEAN_NUM_MOUNTED_WORKERS] = __ddiv_rn(sum_workers, (double)n_iter);
            st[RAMP_SS_MEAN_NUM_MOUNTED_CHANNELS] = __ddiv_rn(sum_channels, (double)n_iter);
            st[RAMP_SS_MEAN_COMPUTE_OVERHEAD_FRAC] = n_frac > 0 ? __ddiv_rn(sum_comp_frac, (double)n_frac) : 0.0;
            st[RAMP_SS_MEAN_COMMUNICATION_OVERHEAD_FRAC] = n_frac > 0 ? __ddiv_rn(sum_comm_frac, (double)n_frac) : 0.0;
            {
                const double dt = st[RAMP_SS_STEP_TIME];
#define RAMP_TP(dst, src) st[dst] = (st[src] != 0.0 && dt != 0.0) ? __ddiv_rn(st[src], dt) : 0.0   /* RCE:1064-1077 */
                RAMP_TP(RAMP_SS_MEAN_COMPUTE_THROUGHPUT, RAMP_SS_COMPUTE_INFO_PROCESSED);
                RAMP_TP(RAMP_SS_MEAN_DEP_THROUGHPUT, RAMP_SS_DEP_INFO_PROCESSED);
                RAMP_TP(RAMP_SS_MEAN_FLOW_THROUGHPUT, RAMP_SS_FLOW_INFO_PROCESSED);
                RAMP_TP(RAMP_SS_MEAN_CLUSTER_THROUGHPUT, RAMP_SS_CLUSTER_INFO_PROCESSED);
                RAMP_TP(RAMP_SS_MEAN_DEMAND_COMPUTE_THROUGHPUT, RAMP_SS_DEMAND_COMPUTE_INFO_PROCESSED);
                RAMP_TP(RAMP_SS_MEAN_DEMAND_DEP_THROUGHPUT, RAMP_SS_DEMAND_DEP_INFO_PROCESSED);
                RAMP_TP(RAMP_SS_MEAN_DEMAND_TOTAL_THROUGHPUT, RAMP_SS_DEMAND_TOTAL_INFO_PROCESSED);
#undef RAMP_TP
            }
            st[RAMP_SS_UTIL_MOUNTED_SUM] = util_mounted_sum;
            st[RAMP_SS_UTIL_CLUSTER_SUM] = util_cluster_sum;
            st[RAMP_SS_NUM_TICKS] = (double)n_iter;
            st[RAMP_SS_JOB_QUEUE_LENGTH] = EI(EI_QUEUED) >= 0 ? 1.0 : 0.0;                       // RCE:1082
            EI(EI_STEP_COUNTER)++;                                                                // RCE:1109
            const bool done = step_is_done(ep, b);
            if (done) {                                                                           // RCE:1111-1121
                while (EI(EI_N_RUNNING) > 0) {
                    step_register_blocked(ep, b, RI(RI_JOB_IDX, 0), st);
                    step_remove_running(ep, b, 0);
                }
                EI(EI_DONE) = 1;
            }
            st[RAMP_SS_DONE] = done ? 1.0 : 0.0;
            ++n_cluster_steps;
            if (cs == 0) {
#pragma unroll
                for (int k = 0; k < RAMP_STEP_STATS_LEN; ++k) st0[k] = st[k];
            }
            // RJPE:394-395: while len(job_queue) == 0 and not is_done(): step(Action())
            if (!s.fuse_empty_steps || done || EI(EI_QUEUED) >= 0) break;
        }
        if (s.fuse_empty_steps) st0[RAMP_SS_DONE] = EI(EI_DONE) ? 1.0 : 0.0;
    } else {
        st0[RAMP_SS_DONE] = EI(EI_DONE) ? 1.0 : 0.0;
        st0[RAMP_SS_STEP_COUNTER] = (double)EI(EI_STEP_COUNTER);
        st0[RAMP_SS_JOB_QUEUE_LENGTH] = EI(EI_QUEUED) >= 0 ? 1.0 : 0.0;
    }
    if (s.stats_out) {
        double* o = s.stats_out + (size_t)b * RAMP_STEP_STATS_LEN;
#pragma unroll
        for (int k = 0; k < RAMP_STEP_STATS_LEN; ++k) o[k] = st0[k];
    }
    if (s.n_cluster_steps_out) s.n_cluster_steps_out[b] = n_cluster_steps;
}

// RCE:202-295 for every episode
__global__ void ramp_reset_kernel(const EpisodeState ep) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int B = ep.B;
    if (b >= B) return;
    double* ef = ep.ef; int32_t* ei = ep.ei;
    for (int f = 0; f < EF_COUNT; ++f) EF(f) = 0.0;
    for (int f = 0; f < EI_COUNT; ++f) EI(f) = 0;
    EI(EI_QUEUED) = -1; EI(EI_LAST_SLOT) = -1; EI(EI_PLAN_SLOT) = -1;
    for (int k = 0; k < ep.max_jobs; ++k) {
        ramp_job_record_t& r = ep.rec[(size_t)b * ep.max_jobs + k];
        r.status = RAMP_JS_NOT_ARRIVED; r.event_seq = 0;
        r.time_arrived = r.time_started = r.time_completed = 0.0; r.jct = r.comm = r.comp = r.util = 0.0;
    }
    EF(EF_NEXT_ARRIVAL) = 0.0;                 // RCE:280
    step_get_next_job(ep, b);                  // RCE:281
    EI(EI_QUEUED) = 0;
}

// packs the per-episode scalars for read-back / NCCL all-gather: [B][RAMP_EP_LEN]
__global__ void ramp_export_episode_state_kernel(const EpisodeState ep, double* out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int B = ep.B;
    if (b >= B) return;
    const double* ef = ep.ef; const int32_t* ei = ep.ei;
    double* o = out + (size_t)b * RAMP_EP_LEN;
    o[RAMP_EP_TIME] = EF(EF_NOW); o[RAMP_EP_NEXT_ARRIVAL] = EF(EF_NEXT_ARRIVAL);
    o[RAMP_EP_NUM_ARRIVED] = EI(EI_NUM_ARRIVED); o[RAMP_EP_NUM_COMPLETED] = EI(EI_NUM_COMPLETED);
    o[RAMP_EP_NUM_BLOCKED] = EI(EI_NUM_BLOCKED); o[RAMP_EP_QUEUED_JOB] = EI(EI_QUEUED);
    o[RAMP_EP_NUM_RUNNING] = EI(EI_N_RUNNING); o[RAMP_EP_STEP_COUNTER] = EI(EI_STEP_COUNTER);
    o[RAMP_EP_LOAD_RATE_SUM] = EF(EF_LOAD_SUM); o[RAMP_EP_LOAD_RATE_N] = EI(EI_LOAD_N);
    o[RAMP_EP_DONE] = EI(EI_DONE); o[RAMP_EP_STATUS] = EI(EI_STATUS);
}

#undef EF
#undef EI
#undef RF
#undef RI

}  // namespace ramp
