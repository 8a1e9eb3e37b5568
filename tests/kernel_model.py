"""Pure-Python model of the tick algorithm the CUDA lookahead kernels implement (ddls_b200/csrc/ramp_kernels.cuh).

NOT the reference's algorithm restated (that is oracle/): this follows the KERNEL's data structures -- rank keys,
per-worker arg-max slots, the double-buffered per-channel winner table that survivors and arrivals vote into, the
one-tick list of ready non-flow deps whose presence makes a zero-length tick that skips the flow pass -- so that the
design can be checked against the reference's recorded runs on CPU, where the kernels cannot run.
"""
import math

import numpy as np


def rank_keys(prio):
    """Larger key wins; ties go to the lower index (ramp_engine.cu make_rank_keys; RCE:56-66, 672-685)."""
    n = len(prio)
    order = sorted(range(n), key=lambda i: (-int(prio[i]), i))
    key = [0] * n
    for r, i in enumerate(order):
        key[i] = n - r
    return key


def run_lookahead_model(job):
    N, E = job.n_ops, job.n_deps
    op_key, dep_key = rank_keys(job.op_prio), rank_keys(job.dep_prio)
    row = job.row_ptr
    in_deg = [0] * N
    for e in range(E):
        in_deg[int(job.dep_dst[e])] += 1
    n_par = [int(x) for x in job.op_n_parents]
    par_done = [0] * N
    ops = [(i, float(job.op_cost[i]) + 0.0) for i in range(N) if in_deg[i] == 0]    # ready ops: (op, remaining)
    flows = []                 # ready flows: [dep, remaining]
    nf = []                    # ready non-flow deps (zero run time: each lives for exactly one tick)
    ck_cur, ck_nxt = {}, {}    # channel -> best key among the ready flows
    t = comm = comp = 0.0
    trace_n, trace_tick = [], []
    ops_completed = deps_completed = 0
    INF = math.inf
    while True:
        # A, B
        wkey = {}
        for op, rem in ops:
            w = int(job.op_worker[op])
            wkey[w] = max(wkey.get(w, 0), op_key[op])
        winners = [(op, rem) for op, rem in ops if wkey[int(job.op_worker[op])] == op_key[op]]
        t_op = min([rem for _, rem in winners], default=INF)
        n_active = len(winners)
        # C, D
        any_nf = len(nf) > 0
        if any_nf:
            t_comm = 0.0
        else:
            t_comm = min([rem for e, rem in flows
                          if int(job.dep_channel[e]) != 0xFFFF and ck_cur.get(int(job.dep_channel[e]), 0) == dep_key[e]], default=INF)
            ck_cur = {}        # cleared: becomes the next tick's "next" table
        vote = ck_cur if any_nf else ck_nxt
        tick = t_comm if t_comm < t_op else t_op
        ticked_ops, ticked_flows = n_active > 0, (not any_nf) and len(flows) > 0
        if ticked_flows:
            comm += tick
        if ticked_ops:
            comp += tick
        t += tick
        trace_n.append(n_active)
        trace_tick.append(tick)
        ops_next = []

        def complete_dep(e):
            child = int(job.dep_dst[e])
            par_done[child] += 1
            if par_done[child] == n_par[child]:
                ops_next.append((child, float(job.op_cost[child]) + 0.0))
        # H
        if any_nf:
            for e in nf:
                complete_dep(e)
            deps_completed += len(nf)
            nf = []
            survivors = flows
        else:
            survivors = []
            for e, rem in flows:
                r2 = rem - (rem if rem < tick else tick)
                if r2 == 0.0:
                    complete_dep(e)
                    deps_completed += 1
                else:
                    survivors.append([e, r2])
                    c = int(job.dep_channel[e])
                    if c != 0xFFFF:
                        ck_nxt[c] = max(ck_nxt.get(c, 0), dep_key[e])
        # G
        win_set = {op for op, _ in winners}
        arrivals = []
        for op, rem in ops:
            if op in win_set:
                r2 = rem - (rem if rem < tick else tick)
                if r2 == 0.0:
                    ops_completed += 1
                    arrivals.extend(range(int(row[op]), int(row[op + 1])))
                    continue
                rem = r2
            ops_next.append((op, rem))
        for e in arrivals:
            if int(job.dep_is_flow[e]):
                survivors.append([e, float(job.dep_run_time[e]) + 0.0])
                c = int(job.dep_channel[e])
                if c != 0xFFFF:
                    vote[c] = max(vote.get(c, 0), dep_key[e])
            else:
                nf.append(e)
        flows = survivors
        ops = ops_next
        finished = ops_completed == N and deps_completed == E
        if finished or math.isinf(tick):
            break
        if not any_nf:
            ck_cur, ck_nxt = ck_nxt, ck_cur
    steps = float(job.num_training_steps)
    return dict(jct=t * steps, comm=comm * steps, comp=comp * steps, n_ticks=len(trace_tick),
                trace_n_active=np.array(trace_n, dtype=np.int32), trace_tick=np.array(trace_tick, dtype=np.float64),
                finished=finished)
