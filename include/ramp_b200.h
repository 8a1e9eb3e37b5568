/*
 * ramp_b200 -- C ABI of the B200-native RAMP cluster simulator hot path.
 *
 * The reference (cwfparsonson/ddls @ 9e0b5ba) is pure Python and has no FFI; the interface this
 * library sits behind is the Python class surface of
 *   ddls/environments/ramp_cluster/ramp_cluster_environment.py  ("RCE")
 *     RampClusterEnvironment.__init__  RCE:75-83    -> ramp_engine_create
 *     RampClusterEnvironment.reset     RCE:202-295  -> ramp_reset
 *     RampClusterEnvironment.step      RCE:894-1179 -> ramp_step_host / ramp_step_device
 *       _place_ops/_schedule_ops/_place_deps/_schedule_deps RCE:1305-1415 -> ramp_register_template (+ per-step mount rows)
 *       _perform_lookahead_job_completion_time + memo dicts RCE:469-518, RCE:269-275 -> memo hash table inside the step
 *       _run_lookahead                  RCE:379-467  -> ramp_lookahead_kernel (also callable alone: ramp_run_lookaheads)
 *       _register_completed_lookahead   RCE:793-888  -> ramp_register_kernel
 *       outer event loop + stats        RCE:942-1167 -> ramp_advance_kernel
 *     RampClusterEnvironment.is_done   RCE:1542-1557 -> stats[RAMP_SS_DONE]
 * batched over `n_episodes` independent cluster instances (one per RL rollout) that advance in lock step.
 * ddls_b200/engine.py is the ctypes binding; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: every function returns RAMP_OK (0) or a negative error code; ramp_last_error() returns a
 * message for the calling thread (the Python wrapper re-raises it as `Exception`, the type the reference
 * raises, e.g. RCE:462, RCE:1328).  All pointers are plain host or device pointers as documented; the
 * library owns only what ramp_engine_create / ramp_register_template allocate.  One host thread per engine.
 */
#ifndef RAMP_B200_H
#define RAMP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAMP_OK 0
#define RAMP_ERR_CUDA (-1)
#define RAMP_ERR_BAD_ARG (-2)
#define RAMP_ERR_CAPACITY (-3)       /* a table / pool configured at create time is full             */
#define RAMP_ERR_SIM (-4)            /* the simulation itself raised (see per-episode status codes)   */

#define RAMP_NO_CHANNEL 0xFFFFu

/* per-lookahead / per-episode status codes written by the kernels */
#define RAMP_ST_OK 0
#define RAMP_ST_INFINITE_TICK 1      /* RCE:462 "Last tick was infinite" (deadlocked job graph)       */
#define RAMP_ST_TRACE_OVERFLOW 2     /* more ticks than ramp_config_t.trace_cap                       */
#define RAMP_ST_TABLE_FULL 3         /* running-job table full (ramp_config_t.max_running)            */
#define RAMP_ST_NO_QUEUED_JOB 4      /* an action was given for an episode with an empty job queue    */
#define RAMP_ST_JOBS_EXHAUSTED 5     /* more arrivals than ramp_config_t.max_jobs                     */
#define RAMP_ST_BAD_TEMPLATE 6       /* an action names a template id that was never registered       */

/* memo modes (RCE:269-277, RCE:488-506) */
#define RAMP_MEMO_REFERENCE 0        /* key = (episode, model, max partition degree): first seen wins, per episode */
#define RAMP_MEMO_EXACT 1            /* key = 64-bit fingerprint of the lowered job: shared across episodes        */
#define RAMP_MEMO_OFF 2              /* every mount runs its lookahead                                              */
#define RAMP_MEMO_SHARED 3           /* reference semantics (per-episode first-seen-wins on (model, degree)) on top of a batch-wide
                                        result cache keyed by the byte-identical lowered job: the lookahead an episode would run is
                                        executed once per batch and survives ramp_reset -- same results, far fewer lookaheads */

typedef struct ramp_engine ramp_engine_t;

typedef struct {
    int32_t device;              /* CUDA device ordinal                                                   */
    int32_t n_episodes;          /* B: independent cluster instances                                      */
    int32_t n_cluster_workers;   /* topology.graph.graph['num_workers'] RCE:991                           */
    int32_t max_jobs;            /* arrivals per episode the arrival stream can hold                      */
    int32_t max_running;         /* rows of the per-episode running-job table (<= n_cluster_workers suffices) */
    int32_t max_templates;       /* lowered jobs that can be registered                                    */
    int32_t memo_mode;           /* RAMP_MEMO_*                                                            */
    int32_t memo_capacity_log2;  /* hash table slots = 1 << this (0 -> sized from n_episodes)              */
    int32_t trace_cap;           /* max ticks recorded per lookahead (0 -> 16384)                          */
    int32_t job_queue_capacity;  /* RCE:205 (default 10; the queue never holds more than one job)          */
    double  machine_epsilon;     /* RCE:83 (1e-7)                                                          */
    double  max_simulation_run_time; /* RCE:204                                                            */
} ramp_config_t;

/* One lowered (partitioned + placed + scheduled) job == the complete input of _run_lookahead.
 * Op index = rank of the op id in sorted() order; dep index = rank of (u, v, k) in sorted() order
 * (== CSR-by-source position).  HOST pointers; copied to the device by ramp_register_template. */
typedef struct {
    int32_t n_ops, n_deps, n_workers, n_channels;
    int32_t num_training_steps;   /* JOB:82, RCE:450-452                                   */
    int32_t model_id;             /* dense id of job.details['model'] RCE:489              */
    int32_t degree;               /* max partition degree RCE:488                          */
    int32_t _pad;
    const double*   op_cost;      /* [N] compute_cost[device_type] RCE:1334                */
    const int64_t*  op_prio;      /* [N] worker.mounted_job_op_to_priority RCE:1397        */
    const uint16_t* op_worker;    /* [N] job-local worker id RCE:1336                      */
    const uint16_t* op_n_parents; /* [N] predecessors that are not also successors JOB:508-523 */
    const int32_t*  row_ptr;      /* [N+1] CSR out-edges                                   */
    const int32_t*  dep_dst;      /* [E]                                                   */
    const double*   dep_run_time; /* [E] init_run_time after RCE:542-560                   */
    const int64_t*  dep_prio;     /* [E] channel.mounted_job_dep_to_priority RCE:1412      */
    const uint16_t* dep_channel;  /* [E] job-local channel id or RAMP_NO_CHANNEL           */
    const uint8_t*  dep_is_flow;  /* [E] RCE:531-536                                       */
} ramp_lowered_job_t;

/* Result of one lookahead (RCE:467). */
typedef struct {
    double  jct, comm, comp;      /* x num_training_steps RCE:450-452 */
    int32_t n_ticks;
    int32_t status;               /* RAMP_ST_*                        */
} ramp_lookahead_result_t;

/* Per-arrival job description (JobsGenerator stays host-side Python; RCE:351-377 reads only these). */
typedef struct {
    double interarrival;          /* gap added to time_next_job_to_arrive when THIS job arrives RCE:363 (inf after the last) */
    double orig_op_mem;           /* original_job.details['job_total_op_memory_cost'] RCE:364, RCE:971 */
    double orig_dep_size;         /* original_job.details['job_total_dep_size']                        */
} ramp_arrival_t;

/* Per-episode, per-step action row.  template_id < 0 is Action() (RJPE:395): no job handled, a queued
 * job is blocked (RCE:914-919). */
typedef struct {
    double  max_acceptable_jct;   /* details['max_acceptable_job_completion_time'][device] RCE:815 */
    double  part_op_mem;          /* partitioned job details['job_total_op_memory_cost'] RCE:966    */
    double  part_dep_size;        /* partitioned job details['job_total_dep_size'] RCE:967          */
    double  flow_size;            /* details['job_total_flow_size'] RCE:882-888                     */
    int32_t n_mounted_workers;    /* len(details['mounted_workers'])  RCE:832                       */
    int32_t n_mounted_channels;   /* len(details['mounted_channels']) RCE:979                       */
    int32_t template_id;          /* from ramp_register_template, or -1                             */
    int32_t flags;                /* RAMP_ACT_*                                                     */
} ramp_action_t;

#define RAMP_ACT_SKIP 1           /* leave this episode untouched this call (e.g. it is done)          */

/* step statistics: double[RAMP_STEP_STATS_LEN] per episode; same names as step_stats RCE:306-338, 1046-1084 */
enum {
    RAMP_SS_STEP_COUNTER = 0, RAMP_SS_STEP_START_TIME, RAMP_SS_STEP_END_TIME, RAMP_SS_STEP_TIME,
    RAMP_SS_NUM_JOBS_COMPLETED, RAMP_SS_NUM_JOBS_ARRIVED, RAMP_SS_NUM_JOBS_BLOCKED, RAMP_SS_JOB_QUEUE_LENGTH,
    RAMP_SS_MEAN_NUM_JOBS_RUNNING, RAMP_SS_MEAN_NUM_MOUNTED_WORKERS, RAMP_SS_MEAN_NUM_MOUNTED_CHANNELS,
    RAMP_SS_MEAN_COMPUTE_OVERHEAD_FRAC, RAMP_SS_MEAN_COMMUNICATION_OVERHEAD_FRAC,
    RAMP_SS_COMPUTE_INFO_PROCESSED, RAMP_SS_DEP_INFO_PROCESSED, RAMP_SS_FLOW_INFO_PROCESSED,
    RAMP_SS_CLUSTER_INFO_PROCESSED, RAMP_SS_DEMAND_COMPUTE_INFO_PROCESSED, RAMP_SS_DEMAND_DEP_INFO_PROCESSED,
    RAMP_SS_DEMAND_TOTAL_INFO_PROCESSED,
    RAMP_SS_MEAN_COMPUTE_THROUGHPUT, RAMP_SS_MEAN_DEP_THROUGHPUT, RAMP_SS_MEAN_FLOW_THROUGHPUT,
    RAMP_SS_MEAN_CLUSTER_THROUGHPUT, RAMP_SS_MEAN_DEMAND_COMPUTE_THROUGHPUT, RAMP_SS_MEAN_DEMAND_DEP_THROUGHPUT,
    RAMP_SS_MEAN_DEMAND_TOTAL_THROUGHPUT,
    RAMP_SS_UTIL_MOUNTED_SUM,     /* sum of the step's 'mean_mounted_worker_utilisation_frac' list RCE:990 */
    RAMP_SS_UTIL_CLUSTER_SUM,     /* sum of the step's 'mean_cluster_worker_utilisation_frac' list RCE:991 */
    RAMP_SS_NUM_TICKS,            /* outer-loop iterations (= length of the two lists above)               */
    RAMP_SS_DONE,                 /* is_done() after the step RCE:1176                                     */
    RAMP_SS_LOOKAHEAD_RAN,        /* 1 if this step executed _run_lookahead (memo miss)                    */
    RAMP_STEP_STATS_LEN
};

/* job record table: one row per (episode, job idx) */
enum { RAMP_JS_NOT_ARRIVED = 0, RAMP_JS_QUEUED = 1, RAMP_JS_RUNNING = 2, RAMP_JS_COMPLETED = 3, RAMP_JS_BLOCKED = 4 };
typedef struct {
    int32_t status;               /* RAMP_JS_*                                        */
    int32_t event_seq;            /* order of the completion / blocking event          */
    double  time_arrived, time_started, time_completed;
    double  jct, comm, comp, util; /* lookahead results + mean_mounted_worker_utilisation_frac RCE:830-832 */
} ramp_job_record_t;

/* ---- lifecycle ---------------------------------------------------------------------------------- */
const char* ramp_last_error(void);
int ramp_engine_create(const ramp_config_t* cfg, ramp_engine_t** out);
int ramp_engine_destroy(ramp_engine_t* eng);
/* the CUDA stream all engine work is issued on (cudaStream_t as void*) */
void* ramp_engine_stream(ramp_engine_t* eng);

/* Copies a lowered job to HBM, derives the priority-rank keys and the initial ready set, and returns its id. */
int ramp_register_template(ramp_engine_t* eng, const ramp_lowered_job_t* job, int32_t* template_id_out);
int ramp_template_count(ramp_engine_t* eng);

/* ---- batched RampClusterEnvironment.reset / step ------------------------------------------------ */
/* RCE:202-295 for every episode.  arrivals: HOST [n_episodes][n_jobs]; clears the memo (RCE:269-275). */
int ramp_reset(ramp_engine_t* eng, const ramp_arrival_t* arrivals, int32_t n_jobs);

/* Overwrites arrival rows [first_job, first_job + n) of one episode (HOST rows).  Lets a host-driven caller (the
 * drop-in RampClusterEnvironment, whose JobsGenerator samples the next job only when needed, RCE:351-377) stream
 * the arrival process instead of fixing it at reset. */
int ramp_set_arrivals(ramp_engine_t* eng, int32_t episode, int32_t first_job, const ramp_arrival_t* rows, int32_t n);
/* How many jobs one episode's arrival stream holds so far (<= max_jobs).  The engine treats `n_jobs - arrived > 0` as the
 * reference's `len(self.jobs_generator) > 0` (RCE:1019-1040, RCE:1542-1557): a host that draws jobs lazily from a generator
 * that never runs dry ('remove_and_repeat' sampling) keeps it one ahead of the arrivals instead of fixing it at reset. */
int ramp_set_job_count(ramp_engine_t* eng, int32_t episode, int32_t n_jobs);
/* max_simulation_run_time / job_queue_capacity of the NEXT ramp_reset (RCE:202-205), so that one engine serves every
 * reset() of a drop-in environment */
int ramp_set_limits(ramp_engine_t* eng, double max_simulation_run_time, int32_t job_queue_capacity);

/* One RampClusterEnvironment.step for every episode.  HOST buffers; the host<->device copies are issued
 * on the engine stream inside the call:  actions [n_episodes] in,  stats [n_episodes][RAMP_STEP_STATS_LEN]
 * out (may be NULL).  fuse_empty_steps != 0 additionally runs, per episode, the RJPE:394-395 loop
 * `while len(job_queue) == 0 and not done: step(Action())` on the device; stats then describe the action
 * step, n_cluster_steps_out[b] (may be NULL) how many cluster steps episode b took in total. */
int ramp_step_host(ramp_engine_t* eng, const ramp_action_t* actions, int32_t fuse_empty_steps,
                   double* stats_out, int32_t* n_cluster_steps_out);
/* Same with DEVICE pointers (inputs already resident in HBM, outputs left there); asynchronous on the
 * engine stream -- call ramp_sync() before reading. */
int ramp_step_device(ramp_engine_t* eng, const ramp_action_t* d_actions, int32_t fuse_empty_steps,
                     double* d_stats_out, int32_t* d_n_cluster_steps_out);
int ramp_sync(ramp_engine_t* eng);
/* Raises (returns RAMP_ERR_SIM) if any episode recorded a RAMP_ST_* error since the last check. */
int ramp_check_status(ramp_engine_t* eng, int32_t* first_bad_episode_out, int32_t* status_out);

/* ---- state read-back (HOST destinations) -------------------------------------------------------- */
int ramp_get_job_records(ramp_engine_t* eng, ramp_job_record_t* out /* [n_episodes][max_jobs] */);
/* per-episode scalars: double[n_episodes][RAMP_EP_LEN] */
enum { RAMP_EP_TIME = 0, RAMP_EP_NEXT_ARRIVAL, RAMP_EP_NUM_ARRIVED, RAMP_EP_NUM_COMPLETED, RAMP_EP_NUM_BLOCKED,
       RAMP_EP_QUEUED_JOB, RAMP_EP_NUM_RUNNING, RAMP_EP_STEP_COUNTER, RAMP_EP_LOAD_RATE_SUM, RAMP_EP_LOAD_RATE_N,
       RAMP_EP_DONE, RAMP_EP_STATUS, RAMP_EP_LEN };
int ramp_get_episode_state(ramp_engine_t* eng, double* out);
/* device pointer of the same table (for NCCL all-gather of episode metrics without a host bounce) */
int ramp_episode_state_device(ramp_engine_t* eng, double** d_out);
/* writes the table into a caller-owned DEVICE buffer [n_episodes][RAMP_EP_LEN] (asynchronous on the engine stream),
 * e.g. a torch tensor that is then all-gathered over NCCL */
int ramp_export_episode_state_to(ramp_engine_t* eng, double* d_dst);
/* memo statistics since the last reset: lookups, hits, lookaheads executed */
int ramp_get_memo_stats(ramp_engine_t* eng, int64_t* lookups, int64_t* hits, int64_t* lookaheads);
/* {lookups, per-episode hits, batch-wide (shared) hits, lookaheads executed} since the last reset */
int ramp_get_memo_stats_ex(ramp_engine_t* eng, int64_t out[4]);
/* the lookahead (memoised or fresh) used by episode `episode`'s most recent mount: result + trace
 * (tick_counter_to_active_workers_tick_size RCE:467); trace buffers are HOST, capacity trace_cap. */
int ramp_get_last_lookahead(ramp_engine_t* eng, int32_t episode, ramp_lookahead_result_t* res,
                            int32_t* trace_n_active, double* trace_tick, int32_t trace_cap);

/* ---- the lookahead kernel on its own ------------------------------------------------------------ */
/* Runs _run_lookahead (RCE:379-467) for n work items; item k uses template template_ids[k] (HOST array).
 * results: HOST [n].  trace_n_active / trace_tick: HOST [n][trace_cap] or NULL.  kernel_ms_out (may be
 * NULL) receives the CUDA-event duration of the kernel alone. */
int ramp_run_lookaheads(ramp_engine_t* eng, const int32_t* template_ids, int32_t n,
                        ramp_lookahead_result_t* results, int32_t* trace_n_active, double* trace_tick,
                        int32_t trace_cap, float* kernel_ms_out);

/* kernel launch counter (gpu_launches in bench.py) and device time spent in the lookahead kernel inside
 * ramp_step_* since the last call (CUDA events on the engine stream) */
int64_t ramp_launch_count(ramp_engine_t* eng);
int ramp_get_lookahead_kernel_time(ramp_engine_t* eng, double* total_ms, int64_t* launches, int64_t* work_items,
                                   int64_t* algorithmic_bytes, int32_t reset);
/* the same 20 N + 19 E + 12 T + 24 accounting on the sizes of the symmetry quotients the thread-per-lookahead kernel really
 * simulated (since the last reset of the counters above; read it BEFORE resetting them) */
int ramp_get_quotient_bytes(ramp_engine_t* eng, int64_t* quotient_bytes);

/* ---- native template expansion: what OpPartition / update_dep_run_times / the SRPT schedulers /
 * FirstFitDepPlacer compute for ONE job placed on a block of servers (sub-op k of every split op on server k), without
 * the reference's Python objects.  Replaces RJPE:320-360 + agents/partitioners/utils.py:42-110 + actions/utils.py:13-393 +
 * srpt_*_scheduler.py for that case; the Python twin is ddls_b200/template_builder.py.  Host-only: needs no GPU. ---- */
typedef struct {
    int32_t n_fwd;                /* forward ops 1..n of the un-mirrored job graph (ddls/utils.py:278-340)          */
    int32_t n_edges;
    const double*  fwd_cost;      /* [n] forward_compute_time                                                       */
    const double*  bwd_cost;      /* [n] backward_compute_time                                                      */
    const double*  act_size;      /* [n] activation_size                                                            */
    const double*  par_size;      /* [n] parameter_size                                                             */
    const int32_t* edge_src;      /* [n_edges] 1-based forward op ids                                               */
    const int32_t* edge_dst;
} ramp_forward_graph_t;

typedef struct {
    int32_t n_servers;            /* servers of the block, in sorted server-id order                                */
    int32_t num_communication_groups;  /* of the whole topology (x in actions/utils.py:40)                          */
    const int32_t* coords;        /* [n_servers][3] (communication group, rack, server)                             */
    double channel_bandwidth, latency, io_latency;   /* topologies/ramp.py:27, heuristic_config.yaml:73-82          */
} ramp_block_t;

enum { RAMP_RUN_TIMES_ONE_TO_ONE = 0, RAMP_RUN_TIMES_REFERENCE = 1 };

/* What the mount scalars (ramp_action_t) are summed from: edge sizes, op memory costs, and the op indices in the job
 * graph's node order (the order the reference's Python sums run in, JOB:224-248). */
typedef struct {
    double*  dep_size;            /* [n_deps] */
    double*  op_mem;              /* [n_ops]  */
    int32_t* node_order;          /* [n_ops]  */
} ramp_expanded_aux_t;

/* Fills `out` (and `aux` if not NULL) with malloc'ed arrays; release with ramp_free_expanded_job / ramp_free_expanded_aux. */
int ramp_expand_template(const ramp_forward_graph_t* graph, int32_t degree, double min_op_run_time_quantum,
                         const ramp_block_t* block, int32_t run_time_mode, int32_t num_training_steps,
                         ramp_lowered_job_t* out, ramp_expanded_aux_t* aux);
void ramp_free_expanded_aux(ramp_expanded_aux_t* aux);

/* RampFirstFitOpPlacer.get (agents/placers/ramp_first_fit_op_placer.py:27-113, agents/placers/utils.py:68-582) for one job:
 * which server every (sub-)op goes to on a possibly busy cluster.  Server index = (cg * racks + rack) * servers + server. */
typedef struct {
    int32_t shape[3];             /* communication groups, racks per group, servers per rack (ramp.py:36-41)        */
    int32_t _pad;
    const double*  free_mem;      /* [servers] memory_capacity - memory_occupied of the server's worker (utils.py:235) */
    const uint8_t* busy;          /* [servers] a job is mounted there (one job per worker, ramp_rules.py:6-39)       */
} ramp_cluster_state_t;

/* splits[n_fwd]: sub-ops per forward op (1 = unsplit).  server_out: servers of op 1's sub-ops 0.., then op 2's ... (the
 * backward op shares them); offset_out[n_fwd + 1]: where each op's servers start.  Returns RAMP_OK, 1 if the job cannot be
 * placed (the reference then leaves it out of the Action and RCE:914-919 blocks it), or a negative RAMP_ERR_*. */
int ramp_first_fit_place(const ramp_forward_graph_t* graph, const int32_t* splits, const ramp_cluster_state_t* state,
                         int32_t* server_out, int32_t* offset_out);
void ramp_free_expanded_job(ramp_lowered_job_t* job);
/* ramp_first_fit_place for many cluster states in one call: busy_words[k] / server_mask_out[k] are bit sets over the servers
 * (n_words x 64 bits each); ok_out[k] = 0 when the job cannot be placed on state k.  Free servers have memory_capacity bytes free
 * (one job per worker, ramp_rules.py:6-39). */
int ramp_first_fit_place_many(const ramp_forward_graph_t* graph, const int32_t* splits, const int32_t shape[3], double memory_capacity,
                              int32_t n_states, int32_t n_words, const uint64_t* busy_words, uint64_t* server_mask_out, uint8_t* ok_out);

/* ---- symmetry quotient of a lowered job (host-only; ddls_b200/csrc/ramp_quotient.cpp).  ramp_register_template applies it
 * by itself; it is exported so that tests can check it without a GPU.  The quotient job is what _run_lookahead
 * (RCE:379-467) is simulated on: one op per class of ops that provably tick in lock step (e.g. the n sub-ops of a
 * partitioned op, agents/partitioners/utils.py:42-110), one dep entry per (class of deps, group of identical channels).
 *   op_weight     class size: what a winning class adds to the trace's active-worker count (RCE:709-715)
 *   op_threshold  n_parents x class size: the class is readied when its counter passes through it (JOB:525-536)
 *   dep_inc       members of the entry: what its completion adds to the child class's counter
 *   op_class[N], dep_entry[E]: where every original op / dep went.  Keys are unique ranks (larger wins). ---- */
typedef struct {
    int32_t n_ops, n_deps, n_workers, n_channels;     /* classes, entries, worker groups, channel groups */
    double*   op_cost;
    uint32_t* op_key;
    uint32_t* op_worker;
    uint32_t* op_weight;
    uint32_t* op_threshold;
    int32_t*  row_ptr;
    int32_t*  dep_dst;
    double*   dep_run_time;
    uint32_t* dep_key;
    uint32_t* dep_channel;       /* split entries: the channel group; 0xFFFFFFFF = none, or merged (see dep_group_mask) */
    uint64_t* dep_group_mask;    /* bit g set: members of the entry lie on channels of group g (valid when masks_valid)   */
    uint8_t*  dep_is_flow;
    uint32_t* dep_inc;
    int32_t*  op_class;
    int32_t*  dep_entry;
    int32_t   merged;            /* 1: one entry per dep class with a group set; 0: one entry per (dep class, group)        */
    int32_t   masks_valid;       /* 0 when there are more than 64 channel groups                                            */
} ramp_quotient_t;
int ramp_quotient_template(const ramp_lowered_job_t* job, ramp_quotient_t* out);
void ramp_free_quotient(ramp_quotient_t* q);


/* ---- device-resident rollouts (SURVEY.md 8f-2 / 8f-4 on the device, 8g): one RampJobPartitioningEnvironment.step for every
 * episode without the host in the loop.  Per episode the action is the maximum partition degree of the queued job
 * (RJPE:300-343).  ramp_env_decide places the job with the reference's first-fit rule (agents/placers/utils.py:394-443, 532-582:
 * the first free block in the (block shape, origin) order the host enumerated into `cand_*`), looks the lowered job up by
 * (model, degree, block geometry) and writes the engine's action rows; ramp_env_advance runs the batched cluster step with the
 * RJPE:394-395 loop fused and then, per episode, the reward (rewards/job_acceptance.py), the occupancy of the cluster, and the
 * dynamic observation features + action mask of the next queued job (observations/...observation.py:80-131, 358-498).
 * Episodes the tables cannot decide (ops of one job split different numbers of times, a block geometry whose template is not
 * registered yet) are listed for the host, which patches their rows before ramp_env_advance. ---- */
typedef struct {
    int32_t shape[3];             /* communication groups, racks per group, servers per rack                       */
    int32_t n_models, max_degree, n_geoms, jobs_per_episode, n_words;   /* n_words = ceil(servers / 64)             */
    int32_t apply_action_mask;    /* 1: an invalid action is an error (RJPE:317-319); 0: it becomes action 0       */
    int32_t num_training_steps;
    double  fail_reward, success_reward;
    const int32_t*  cand_ptr;     /* [max_degree + 2] candidates of degree d are [cand_ptr[d], cand_ptr[d + 1])    */
    const uint64_t* cand_mask;    /* [n_cand][n_words] servers of the block (bit set)                              */
    const int32_t*  cand_geom;    /* [n_cand] geometry index of the block (what the lowered job depends on)        */
    const uint8_t*  uniform;      /* [n_models][max_degree + 1] 1: every op of the model takes `degree` sub-ops and the job fits
                                     the block's memory, so placing it is one first-fit search                    */
    const uint8_t*  shape_ok;     /* [max_degree + 1] a RAMP-symmetric block shape exists (action mask)            */
    const double*   model_params; /* [n_models][5] sequential completion time, #ops, #deps, op memory, dep size    */
    const double*   jobs_params;  /* [8][2] (min, max) of JobsGenerator.jobs_params in observation.PARAM_KEYS order */
} ramp_env_config_t;

/* device buffers of the environment (valid until the engine is destroyed): what a device-resident policy reads and writes */
typedef struct {
    int32_t*  actions;            /* [B] in: max partition degree chosen for the queued job (0 = do not place)      */
    double*   reward;             /* [B] out                                                                        */
    uint8_t*  done;               /* [B] out                                                                        */
    int32_t*  queued_model;       /* [B] out: model of the queued job (-1 none)                                     */
    float*    obs_dynamic;        /* [B][11] out: graph features that change per job / cluster state                */
    uint8_t*  action_mask;        /* [B][max_degree + 1] out                                                        */
    uint64_t* busy;               /* [B][n_words] occupancy of the cluster                                          */
    int32_t*  template_id;        /* [B] template the decision mounted (-1 none)                                    */
    int32_t   n_episodes, n_actions, n_models;   /* B, max_degree + 1, job types                                    */
} ramp_env_buffers_t;

/* page-locked HOST arrays of the same shapes (valid until the engine is destroyed): give them to ramp_env_decide / ramp_env_read
 * so that the per-step copies are true asynchronous DMA transfers (busy / template_id are not mirrored: NULL) */
int ramp_env_host_mirror(ramp_engine_t* eng, ramp_env_buffers_t* out);

int ramp_env_create(ramp_engine_t* eng, const ramp_env_config_t* cfg);
int ramp_env_set_template(ramp_engine_t* eng, int32_t model, int32_t degree, int32_t geom, int32_t template_id, const double mount[6]);
/* model_of / frac / max_acceptable_jct (NaN = frac x sequential time): HOST [n_episodes][jobs_per_episode]; arrivals as ramp_reset */
int ramp_env_reset(ramp_engine_t* eng, const int32_t* model_of, const double* frac, const double* max_acceptable_jct,
                   const ramp_arrival_t* arrivals);
int ramp_env_buffers(ramp_engine_t* eng, ramp_env_buffers_t* out);
/* actions: HOST [n_episodes] (copied) or NULL (already in ramp_env_buffers_t.actions).  need_host_out: HOST [n_episodes] list of
 * episodes the tables could not decide, n_need_host_out their number (pass NULL for both when `uniform` covers every model). */
int ramp_env_decide(ramp_engine_t* eng, const int32_t* actions, int32_t* n_need_host_out, int32_t* need_host_out);
int ramp_env_patch(ramp_engine_t* eng, int32_t episode, int32_t template_id, const uint64_t* server_mask, const double mount[6]);
int ramp_env_advance(ramp_engine_t* eng);
/* HOST copies of the outputs (any may be NULL) */
/* The reference leaves step_stats['mean_mounted_worker_utilisation_frac'] / ['mean_cluster_worker_utilisation_frac'] as LISTS with
 * one entry per outer-loop iteration of the step (RCE:989-994); RAMP_SS_UTIL_*_SUM / RAMP_SS_NUM_TICKS carry their sum and length.
 * ramp_enable_tick_lists makes the step kernel also keep the entries of the last cluster step (up to `cap` per episode);
 * ramp_get_tick_lists copies one episode's to HOST arrays [cap] and returns the length in n_out (RAMP_ERR_CAPACITY when the
 * step had more iterations than `cap` given to ramp_enable_tick_lists). */
int ramp_enable_tick_lists(ramp_engine_t* eng, int32_t cap);
int ramp_get_tick_lists(ramp_engine_t* eng, int32_t episode, double* mounted_out, double* cluster_out, int32_t cap, int32_t* n_out);
/* step statistics / cluster-step counts of the last ramp_env_advance (the engine's own buffers): HOST [n_episodes][RAMP_STEP_STATS_LEN], [n_episodes] */
int ramp_get_last_step_stats(ramp_engine_t* eng, double* stats_out, int32_t* n_cluster_steps_out);
/* HOST copies of the occupancy [n_episodes][n_words], of the actions the device holds, and of the number of decisions every episode
 * has taken since ramp_env_reset (= its env-steps; a finished episode takes none) -- any may be NULL */
int ramp_env_read_state(ramp_engine_t* eng, uint64_t* busy_out, int32_t* actions_out, int32_t* n_decided_out);
/* also raises what ramp_check_status would (RAMP_ERR_SIM) -- one synchronisation per step for a host-side policy */
int ramp_env_read(ramp_engine_t* eng, double* reward, uint8_t* done, int32_t* queued_model, float* obs_dynamic, uint8_t* action_mask);


/* ---- the GNN policy forward on the device: GNNPolicy.forward (ml_models/policies/gnn_policy.py:137-296) =
 * num_rounds MeanPool message-passing rounds over the queued job's graph (ml_models/models/mean_pool.py:107-150, gnn.py:84-92),
 * the mean of the node embeddings, a graph module over [graph features | action mask] (gnn_policy.py:96-109), and an RLlib
 * FullyConnectedNetwork read-out (fcnet_hiddens, separate value branch) with the log-mask added to the logits
 * (gnn_policy.py:283-290).  What the policy sees of a job's ops and deps is fixed per job TYPE (node / edge features:
 * observation.py:503-567), so the message passing is run once per model and weight set (ramp_policy_embed: one CTA per model);
 * per decision only the graph module + read-out run (one warp per episode, weights staged in shared memory), reading the
 * environment's device buffers and writing its `actions` -- a rollout step never leaves the device.  fp32 throughout.
 * DGL semantics restated: a node without incoming edges keeps a zero embedding after a round (dgl update_all fills
 * zero-in-degree nodes with zeros); every other node averages reduce_module over [its own (node | zeros) state, messages]. ---- */
typedef struct {
    int32_t in_features_node, in_features_edge, in_features_graph;   /* 5, 2, 17 (gnn.yaml)                          */
    int32_t n_actions;                                               /* action_space.n = max_partitions_per_op + 1    */
    int32_t out_features_msg, out_features_hidden, out_features_node, out_features_graph;   /* 32, 64, 16, 8          */
    int32_t num_rounds;                                              /* >= 2 (gnn.py:40-41)                           */
    int32_t fcnet_hidden;                                            /* one hidden layer of the read-out (256)        */
    int32_t aggregator_activation;                                   /* 0 relu, 1 leaky_relu(0.01)                    */
    int32_t fcnet_activation;                                        /* 0 relu, 2 tanh                                */
    int32_t apply_action_mask;
    int32_t n_models;
} ramp_policy_config_t;
typedef struct ramp_policy ramp_policy_t;
/* number of fp32 words of the weight blob for `cfg`, in torch state_dict order of GNNPolicy: per round [node LN w,b | node
 * Linear W,b | edge LN w,b | edge Linear W,b | reduce LN w,b | reduce Linear W,b]; graph LN w,b | graph Linear W,b; read-out hidden
 * W,b | logits W,b | value hidden W,b | value W,b.  Every W is [out][in] row-major as torch.nn.Linear keeps it. */
int64_t ramp_policy_weight_count(const ramp_policy_config_t* cfg);
int ramp_policy_create(int device, const ramp_policy_config_t* cfg, ramp_policy_t** out);
void ramp_policy_destroy(ramp_policy_t* p);
/* weights: HOST [ramp_policy_weight_count]; the per-model embeddings become stale until the next ramp_policy_embed */
int ramp_policy_set_weights(ramp_policy_t* p, const float* weights, int64_t n);
/* one job type: HOST node features [n_nodes][in_node], edge features [n_edges][in_edge], edge endpoints (node indices), and the
 * per-graph statistics that sit between the dynamic graph features ([6]: observation.py:425-469) */
int ramp_policy_set_model(ramp_policy_t* p, int32_t model, int32_t n_nodes, int32_t n_edges, const float* node_features,
                          const float* edge_features, const int32_t* edges_src, const int32_t* edges_dst, const float* graph_static);
/* message passing + node mean of every registered model; embeddings_out: HOST [n_models][out_features_node] or NULL */
int ramp_policy_embed(ramp_policy_t* p, float* embeddings_out);
/* read-out on HOST inputs (tests, host-side policies): model [n], graph_features [n][in_features_graph], action_mask
 * [n][n_actions] -> logits [n][n_actions], value [n] (either may be NULL) */
int ramp_policy_forward(ramp_policy_t* p, int32_t n, const int32_t* model, const float* graph_features, const uint8_t* action_mask,
                        float* logits_out, float* value_out);
/* one decision for every episode of `eng`'s environment on the engine's stream: reads queued_model / obs_dynamic / action_mask,
 * writes ramp_env_buffers_t.actions (greedy: the first maximal logit; sample: categorical over softmax(logits) from a counter-based
 * generator keyed by (seed, episode)); finished episodes get action 0.  No host transfer. */
int ramp_policy_act(ramp_policy_t* p, ramp_engine_t* eng, int32_t sample, uint64_t seed);
/* HOST copies of the last ramp_policy_act: logits [B][n_actions], value [B], log-probability of the chosen action [B], actions [B] (any may be NULL) */
int ramp_policy_read(ramp_policy_t* p, ramp_engine_t* eng, float* logits_out, float* value_out, float* logp_out, int32_t* actions_out);

/* A rollout segment recorded on the device, for a trainer: ramp_policy_trajectory_begin sizes [horizon][n_episodes] slots;
 * ramp_policy_trajectory_record(t, 0) after ramp_policy_act stores what the policy saw and decided in slot t (dynamic graph features,
 * model of the queued job, action mask, action, its log-probability, the value estimate), (t, 1) after ramp_env_advance stores the
 * reward and done flag that came back -- device-to-device copies on the engine's stream, no synchronisation;
 * ramp_policy_trajectory_read copies the first n_steps slots to HOST arrays (any may be NULL): ONE transfer per segment instead of
 * one per step.  (The static part of the observation is a function of `model`: ramp_policy_set_model.) */
/* page-locked host memory for the read-back targets (any HOST pointer works; page-locked ones make the copies asynchronous DMA) */
void* ramp_pinned_alloc(size_t bytes);
void ramp_pinned_free(void* ptr);
int ramp_policy_trajectory_begin(ramp_policy_t* p, ramp_engine_t* eng, int32_t horizon);
int ramp_policy_trajectory_record(ramp_policy_t* p, ramp_engine_t* eng, int32_t t, int32_t phase);
int ramp_policy_trajectory_read(ramp_policy_t* p, ramp_engine_t* eng, int32_t n_steps, float* obs_dynamic_out, int32_t* model_out,
                                uint8_t* action_mask_out, int32_t* action_out, float* logp_out, float* value_out, double* reward_out,
                                uint8_t* done_out);

#ifdef __cplusplus
}
#endif
#endif
