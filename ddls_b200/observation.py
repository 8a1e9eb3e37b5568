"""Observation encoder of RampJobPartitioningEnvironment on flat arrays (host side).

Restates RampJobPartitioningObservation._encode_obs and the feature functions it calls
(ramp_job_partitioning/observations/ramp_job_partitioning_observation.py:80-131 action mask, :204-241 padding, :243-306
encode, :358-621 features) on the job's arrays instead of networkx / dict objects: per-op and per-dep features are
vectorised divisions, graph features are a dozen min-max normalisations, the action mask needs only the number of free
workers and the RAMP shape.  Pinned against 81 observations recorded from the unmodified reference
(tests/test_observation.py, fixture tests/fixtures/obs_cases.npz from oracle/gen_obs_cases.py)."""
from __future__ import annotations

import math

import numpy as np

PARAM_KEYS = ('job_total_num_ops', 'job_total_num_deps', 'job_sequential_completion_times', 'max_acceptable_job_completion_times',
              'max_acceptable_job_completion_time_fracs', 'job_total_op_memory_costs', 'job_total_dep_sizes', 'job_num_training_steps')


def _block_shapes_exist(action: int, shape) -> bool:
    """get_factor_pairs + get_block_shapes (agents/placers/utils.py:445-530): is there any RAMP-symmetric block of `action` servers?"""
    for i in range(1, action + 1):
        if action % i:
            continue
        p0, p1 = action // i, i
        var = math.sqrt(p0)
        if (var % 1 == 0) and (var <= shape[0] and var <= shape[1] and p1 <= shape[2]):
            return True
        if not (p0 > shape[0] or p0 > shape[1] or p1 > shape[2]):
            return True
    return False


def action_set_and_mask(max_partitions_per_op: int, n_workers: int, n_mounted_workers: int, shape):
    """get_action_set_and_action_mask (observation.py:80-131): action 0 always valid, odd actions > 1 never, an even action (or 1)
    needs that many free workers and (if > 1) a symmetric block shape."""
    action_set, mask = [0], [True]
    for a in range(1, max_partitions_per_op + 1):
        action_set.append(a)
        ok = False
        if (a > 1 and a % 2 == 0) or a == 1:
            if a <= n_workers - n_mounted_workers:
                ok = True if a == 1 else _block_shapes_exist(a, shape)
        mask.append(ok)
    return action_set, mask


def _norm(x, lo, hi):
    return (x - lo) / (hi - lo) if hi - lo != 0 else 1


def encode_observation(op_compute, op_memory, op_depth, edge_src, edge_dst, edge_size, params, *, max_compute_cost, max_compute_op,
                       max_memory_cost, max_memory_op, max_dep_size, max_dep_index, max_depth, sequential_completion_time,
                       max_acceptable_jct, max_acceptable_frac, total_op_memory, total_dep_size, num_training_steps,
                       n_mounted_workers, n_jobs_running, n_workers, shape, max_partitions_per_op, max_nodes, machine_epsilon=1e-7):
    """Arrays are in the job graph's node / edge iteration order (the mirrored, un-partitioned job of the queue head);
    params[k] = (min, max) of jobs_params for PARAM_KEYS[k]; max_compute_op / max_memory_op: op index or -1 (the reference
    compares the op id with job.details['max_compute_node'], a per-device dict, so that feature is never set)."""
    op_compute, op_memory, op_depth = (np.asarray(a, dtype=np.float64) for a in (op_compute, op_memory, op_depth))
    edge_size = np.asarray(edge_size, dtype=np.float64)
    N, E = len(op_compute), len(edge_size)
    eps = machine_epsilon
    # ---- node features (observation.py:522-567): compute / max, is-max, memory / max, is-max, depth / max depth ----
    nf = np.zeros((N, 5), dtype=np.float64)
    nf[:, 0] = op_compute / max_compute_cost if max_compute_cost != 0 else 0.0
    if max_compute_op >= 0:
        nf[max_compute_op, 1] = 1.0
    nf[:, 2] = op_memory / max_memory_cost if max_memory_cost != 0 else 0.0
    if max_memory_op >= 0:
        nf[max_memory_op, 3] = 1.0
    nf[:, 4] = op_depth / max_depth
    nf = np.where(nf < 0, nf + eps, nf)
    # ---- edge features (observation.py:503-520) ----
    ef = np.zeros((E, 2), dtype=np.float64)
    ef[:, 0] = edge_size / max_dep_size
    if max_dep_index >= 0:
        ef[max_dep_index, 1] = 1.0
    ef = np.where(ef < 0, ef + eps, ef)
    # ---- graph features (observation.py:358-498) ----
    P = {k: (float(params[i][0]), float(params[i][1])) for i, k in enumerate(PARAM_KEYS)}
    cc, mc = op_compute / max_compute_cost, op_memory / max_memory_cost
    gf = [
        _norm(N, *P['job_total_num_ops']), _norm(E, *P['job_total_num_deps']),
        _norm(sequential_completion_time, *P['job_sequential_completion_times']),
        _norm(max_acceptable_jct, *P['max_acceptable_job_completion_times']),
        _norm(max_acceptable_frac, *P['max_acceptable_job_completion_time_fracs']), max_acceptable_frac,
        _norm(total_op_memory, *P['job_total_op_memory_costs']), _norm(total_dep_size, *P['job_total_dep_sizes']),
        _norm(num_training_steps, *P['job_num_training_steps']),
        np.mean(cc), np.median(cc), np.mean(mc), np.median(mc),
        np.mean(edge_size) / max_dep_size, np.median(edge_size) / max_dep_size,
    ]
    gf = [g + eps if g < 0 else g for g in gf]
    net = [n_mounted_workers / n_workers, n_jobs_running / n_workers]
    net = [g + eps if g < 0 else g for g in net]
    action_set, mask = action_set_and_mask(max_partitions_per_op, n_workers, n_mounted_workers, shape)
    graph_features = np.concatenate((np.array(gf + net, dtype=np.float32), np.array(mask, dtype=np.int16)))
    # ---- padding (observation.py:204-241) ----
    max_edges = int(max_nodes * (max_nodes - 1) / 2)
    if N > max_edges or E > max_edges:
        raise Exception(f'ERROR: Trying to encode job with {N} nodes / {E} edges but max nodes set to {max_nodes}. Increase max nodes or use smaller computation graphs.')
    node_features = np.zeros((max_nodes, 5), dtype=np.float32); node_features[:N] = nf.astype(np.float32)
    edge_features = np.zeros((max_edges, 2), dtype=np.float32); edge_features[:E] = ef.astype(np.float32)
    src = np.zeros(max_edges, dtype=np.float32); src[:E] = np.asarray(edge_src, dtype=np.float32)
    dst = np.zeros(max_edges, dtype=np.float32); dst[:E] = np.asarray(edge_dst, dtype=np.float32)
    return {'action_set': np.array(action_set, dtype=np.int16), 'action_mask': np.array(mask, dtype=np.int16),
            'node_features': node_features, 'edge_features': edge_features, 'graph_features': graph_features,
            'edges_src': src, 'edges_dst': dst, 'node_split': np.array([N], dtype=np.float32),
            'edge_split': np.array([E], dtype=np.float32)}


def job_arrays(g, num_training_steps: int = 1):
    """The flat arrays ``encode_observation`` reads, derived from a forward-pass profile alone (``synth.ForwardGraph``) in the
    node / edge iteration order of the job graph the reference builds from it:

      * nodes (utils.py:342-398 mirror_graph / combine_graphs): forward 1..n in file order, then the backward mirrors in the
        order their forward ops were visited, i.e. 2n, 2n-1, ..., n+1;
      * edges: by source node in that order, then insertion order -- the forward edges by file order, the backward edges
        (2n-(v-1), 2n-(u-1)) in forward-edge iteration order, and the join n -> n+1 last in n's adjacency; an edge carries
        its source op's activation size (utils.py:394-396);
      * depth (job.py:23-29): number of nodes on the shortest path from the first source node, 0 if unreachable;
      * max memory op / max dep (job.py:306-325): the FIRST strict maximum in iteration order.

    Returns a dict of the positional and keyword arguments of ``encode_observation`` that depend on the job type only."""
    n = g.n
    order = list(range(1, n + 1)) + [2 * n - (i - 1) for i in range(1, n + 1)]
    idx = {v: k for k, v in enumerate(order)}
    N = 2 * n
    comp = np.zeros(N); mem = np.zeros(N)
    for i in range(1, n + 1):
        comp[idx[i]] = g.fwd[i - 1]; comp[idx[2 * n - (i - 1)]] = g.bwd[i - 1]
        mem[idx[i]] = mem[idx[2 * n - (i - 1)]] = g.act[i - 1] + g.par[i - 1]
    adj = {v: [] for v in order}                                   # insertion-ordered adjacency, as networkx keeps it
    for (u, v) in g.edges:
        if v not in adj[u]:
            adj[u].append(v)
    fwd_iter = [(u, v) for u in range(1, n + 1) for v in adj[u]]   # forward_graph.edges()
    for (u, v) in fwd_iter:
        bu, bv = 2 * n - (v - 1), 2 * n - (u - 1)
        if bv not in adj[bu]:
            adj[bu].append(bv)
    if (n + 1) not in adj[n]:
        adj[n].append(n + 1)
    act = {}
    for i in range(1, n + 1):
        act[i] = act[2 * n - (i - 1)] = g.act[i - 1]
    edges = [(u, v) for u in order for v in adj[u]]
    src = np.array([idx[u] for u, _ in edges], dtype=np.int64)
    dst = np.array([idx[v] for _, v in edges], dtype=np.int64)
    size = np.array([act[u] for u, _ in edges], dtype=np.float64)
    # breadth-first depth from the source (the first node without parents)
    has_parent = set(v for _, v in edges)
    source = next(v for v in order if v not in has_parent)
    depth = {source: 1}
    frontier = [source]
    while frontier:
        nxt = []
        for u in frontier:
            for v in adj[u]:
                if v not in depth:
                    depth[v] = depth[u] + 1
                    nxt.append(v)
        frontier = nxt
    dep = np.array([depth.get(v, 0) for v in order], dtype=np.float64)
    max_mem, max_mem_op = 0.0, -1
    for k in range(N):
        if mem[k] > max_mem:
            max_mem, max_mem_op = mem[k], k
    max_dep, max_dep_idx = 0.0, -1
    for k in range(len(edges)):
        if size[k] > max_dep:
            max_dep, max_dep_idx = size[k], k
    total_mem = 0.0
    for k in range(N):
        total_mem += mem[k]
    total_dep = 0.0
    for k in range(len(edges)):
        total_dep += size[k]
    seq = 0.0
    for k in range(N):
        seq += comp[k]
    return {'op_compute': comp, 'op_memory': mem, 'op_depth': dep, 'edge_src': src, 'edge_dst': dst, 'edge_size': size,
            'max_compute_cost': float(comp.max()), 'max_compute_op': -1, 'max_memory_cost': float(max_mem), 'max_memory_op': max_mem_op,
            'max_dep_size': float(max_dep), 'max_dep_index': max_dep_idx, 'max_depth': float(dep.max()),
            'sequential_completion_time': seq * num_training_steps, 'total_op_memory': total_mem, 'total_dep_size': total_dep}


def static_observation(g, machine_epsilon: float = 1e-7):
    """What the policy observes of a job type whatever the cluster holds: node features [N, 5], edge features [E, 2], edge
    endpoints and the six per-graph statistics (mean / median of the normalised op compute, op memory and dep size:
    observation.py:425-469) -- ``encode_observation``'s arrays without padding."""
    a = job_arrays(g)
    cc, mc = a['op_compute'] / a['max_compute_cost'], a['op_memory'] / a['max_memory_cost']
    nf = np.zeros((len(cc), 5)); nf[:, 0] = cc; nf[:, 2] = mc; nf[:, 4] = a['op_depth'] / a['max_depth']
    if a['max_memory_op'] >= 0:
        nf[a['max_memory_op'], 3] = 1.0
    ef = np.zeros((len(a['edge_size']), 2)); ef[:, 0] = a['edge_size'] / a['max_dep_size']
    if a['max_dep_index'] >= 0:
        ef[a['max_dep_index'], 1] = 1.0
    stats = [np.mean(cc), np.median(cc), np.mean(mc), np.median(mc), np.mean(a['edge_size']) / a['max_dep_size'],
             np.median(a['edge_size']) / a['max_dep_size']]
    return {'node_features': nf.astype(np.float32), 'edge_features': ef.astype(np.float32),
            'edges_src': a['edge_src'].astype(np.int32), 'edges_dst': a['edge_dst'].astype(np.int32),
            'graph_static': np.array(stats, dtype=np.float32)}
