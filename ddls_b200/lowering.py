"""The drop-in boundary: lower a reference-style ``Action`` into flat arrays.

Reads, by attribute / key name only (duck-typed, so it works on the unmodified
reference classes as well as on ``ddls_b200.host`` mirrors):

  * ``action.actions['op_partition'].partitioned_jobs[job_id]``  (actions/op_partition.py:8-78)
  * ``action.actions['op_placement'].action[job_id][op_id] -> worker_id``  (actions/op_placement.py:7-30)
  * ``action.actions['op_schedule'].action[worker_id][job_id][op_id] -> priority``  (actions/op_schedule.py:3-15)
  * ``action.actions['dep_placement'].action[job_id][dep_id] -> set(channel_id | None)``  (actions/dep_placement.py:6-34)
  * ``action.actions['dep_schedule'].action[channel_id][job_id][dep_id] -> priority``  (actions/dep_schedule.py:3-15)
  * ``cluster.topology.graph.graph['worker_to_node' | 'worker_to_type']``  (ramp_cluster_environment.py:169-198)

and reproduces what ``RampClusterEnvironment._place_ops/_schedule_ops/_place_deps/_schedule_deps/
_register_running_job`` (ramp_cluster_environment.py:1305-1423) write into the per-device dicts before
``_run_lookahead`` (ramp_cluster_environment.py:379) consumes them.
"""
from __future__ import annotations

import numpy as np

from .lowered import LoweredJob, MountScalars, NO_CHANNEL


class ModelRegistry:
    """job.details['model'] (str) -> dense int id; the memo key (RCE:488-489) uses the id."""

    def __init__(self):
        self._ids = {}

    def get(self, model) -> int:
        if model not in self._ids:
            self._ids[model] = len(self._ids)
        return self._ids[model]

    def __len__(self):
        return len(self._ids)


def lower_job(cluster, action, job_id, model_registry: ModelRegistry = None) -> LoweredJob:
    """Lower the job ``job_id`` of ``action`` (must be handled by all five action parts)."""
    parts = action.actions
    op_partition, op_placement, op_schedule = parts['op_partition'], parts['op_placement'], parts['op_schedule']
    dep_placement, dep_schedule = parts['dep_placement'], parts['dep_schedule']
    if any(p is None for p in (op_partition, op_placement, op_schedule, dep_placement, dep_schedule)):
        raise Exception(f'Action for job_id {job_id} is missing one of the five action parts.')

    job = op_partition.partitioned_jobs[job_id]
    g = job.computation_graph
    worker_to_node = cluster.topology.graph.graph['worker_to_node']
    worker_to_type = cluster.topology.graph.graph['worker_to_type']

    # ---- ops ---------------------------------------------------------------------------
    op_ids = sorted(g.nodes)                       # RCE:56 sorts op ids; rank = index
    op_index = {op: i for i, op in enumerate(op_ids)}
    N = len(op_ids)
    placement = op_placement.action[job_id]
    try:
        op_worker_global = [placement[op] for op in op_ids]
    except KeyError as e:
        raise Exception(f'Op {e} of job_id {job_id} has no worker in the OpPlacement (RCE:527 would KeyError).')
    worker_ids = sorted(set(op_worker_global))
    worker_index = {w: i for i, w in enumerate(worker_ids)}
    if len(worker_ids) > 0xFFFF:
        raise Exception('More than 65535 mounted workers per job is not supported.')
    op_worker = np.fromiter((worker_index[w] for w in op_worker_global), dtype=np.uint16, count=N)
    op_cost = np.empty(N, dtype=np.float64)
    op_prio = np.empty(N, dtype=np.int64)
    sched = op_schedule.action
    for i, op in enumerate(op_ids):
        w = op_worker_global[i]
        op_cost[i] = g.nodes[op]['compute_cost'][worker_to_type[w]]    # RCE:1334 / JOB:32-39
        p = sched[w][job_id][op]                                        # RCE:1397
        if int(p) != p:
            raise Exception(f'Non-integer op priority {p!r} is not supported by the lowering.')
        op_prio[i] = int(p)

    # ---- deps --------------------------------------------------------------------------
    dep_ids = sorted(g.edges(keys=True)) if g.is_multigraph() else sorted((u, v, 0) for u, v in g.edges)
    E = len(dep_ids)
    if any(k != 0 for (_, _, k) in dep_ids):
        raise Exception('Multigraph edge keys != 0 never become ready in the reference (JOB:503-506); not supported.')
    row_ptr = np.zeros(N + 1, dtype=np.int32)
    dep_dst = np.empty(E, dtype=np.int32)
    dep_run_time = np.empty(E, dtype=np.float64)
    dep_is_flow = np.empty(E, dtype=np.uint8)
    dep_channel_global = [None] * E
    flow_size = 0
    dplace = dep_placement.action[job_id] if job_id in dep_placement.action else {}
    succ = {op: set(g.successors(op)) for op in op_ids}
    for e, (u, v, k) in enumerate(dep_ids):
        iu, iv = op_index[u], op_index[v]
        row_ptr[iu + 1] += 1
        dep_dst[e] = iv
        attrs = g[u][v][k]
        same_server = worker_to_node[op_worker_global[iu]] == worker_to_node[op_worker_global[iv]]
        size = attrs['size']
        if same_server or size == 0:                                    # RCE:551-556, RCE:531-536
            dep_run_time[e] = 0.0
            dep_is_flow[e] = 0
        else:
            rt = attrs['init_run_time']                                 # set by OpPlacement -> AU:13
            if rt is None:
                raise Exception(f'Dep {(u, v, k)} has no init_run_time; was an OpPlacement built for this job?')
            dep_run_time[e] = rt
            dep_is_flow[e] = 1
            # RCE:882-888 runs after job.reset_job() (RCE:861) has set every init_run_time to None
            # (JOB:445-452), so set_dep_init_run_time returns None != 0 for EVERY flow: all flows count.
            flow_size += size
        chans = [c for c in dplace.get((u, v, k), ()) if c is not None] # RCE:1362-1364
        if len(chans) > 1:
            raise Exception(f'Dep {(u, v, k)} is placed on {len(chans)} channels; RAMP is a complete graph '
                            f'(topologies/ramp.py:43-46) so every flow is one hop -- multi-channel deps are not supported.')
        if chans:
            dep_channel_global[e] = chans[0]
    np.cumsum(row_ptr, out=row_ptr)
    channel_ids = sorted({c for c in dep_channel_global if c is not None})
    if len(channel_ids) >= NO_CHANNEL:
        raise Exception('More than 65534 mounted channels per job is not supported.')
    channel_index = {c: i for i, c in enumerate(channel_ids)}
    dep_channel = np.full(E, NO_CHANNEL, dtype=np.uint16)
    dep_prio = np.zeros(E, dtype=np.int64)
    dsched = dep_schedule.action
    for e, c in enumerate(dep_channel_global):
        if c is not None:
            dep_channel[e] = channel_index[c]
            p = dsched[c][job_id][dep_ids[e]]                           # RCE:1412
            if int(p) != p:
                raise Exception(f'Non-integer dep priority {p!r} is not supported by the lowering.')
            dep_prio[e] = int(p)

    # parents: predecessors that are not also successors (JOB:508-523)
    op_n_parents = np.zeros(N, dtype=np.int64)
    for op in op_ids:
        n = 0
        for p in g.predecessors(op):
            if p not in succ[op]:
                n += 1
        op_n_parents[op_index[op]] = n
    if N and op_n_parents.max() > 0xFFFF:
        raise Exception('An op with more than 65535 parents is not supported.')

    device_type = worker_to_type[worker_ids[0]] if worker_ids else None
    model = job.details['model'] if 'model' in job.details else ''
    degree = op_partition.job_id_to_max_partition_degree[job_id]        # RCE:488
    mount = MountScalars(
        max_acceptable_jct=float(job.details['max_acceptable_job_completion_time'][device_type]),  # RCE:815
        part_op_mem=float(job.details['job_total_op_memory_cost']),    # RCE:966
        part_dep_size=float(job.details['job_total_dep_size']),        # RCE:967
        flow_size=float(flow_size),
        n_mounted_workers=len(worker_ids),
        n_mounted_channels=len(channel_ids))
    lj = LoweredJob(n_ops=N, n_deps=E, n_workers=len(worker_ids), n_channels=len(channel_ids),
                    num_training_steps=int(job.num_training_steps),
                    model_id=(model_registry.get(model) if model_registry is not None else 0),
                    degree=int(degree),
                    op_cost=op_cost, op_prio=op_prio, op_worker=op_worker,
                    op_n_parents=op_n_parents.astype(np.uint16), row_ptr=row_ptr, dep_dst=dep_dst,
                    dep_run_time=dep_run_time, dep_prio=dep_prio, dep_channel=dep_channel, dep_is_flow=dep_is_flow,
                    mount=mount, model=model, op_ids=op_ids, dep_ids=dep_ids,
                    worker_ids=worker_ids, channel_ids=channel_ids)
    return lj.canonicalise()
