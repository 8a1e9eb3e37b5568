"""Reference-shaped objects built from a LoweredJob: a Job-like object (networkx MultiDiGraph + details) and the
five Action parts with the reference's attribute names (actions/*.py).  Used by the tests and examples on boxes
where the reference package is not installed: ``lower_job(cluster, action, job_id)`` applied to these objects
reproduces the LoweredJob they were built from, so the drop-in ``RampClusterEnvironment`` can be driven end to end."""
from __future__ import annotations

from collections import defaultdict

import networkx as nx
import numpy as np

from ..lowered import LoweredJob, NO_CHANNEL


class SyntheticJob:
    """The attributes of ddls.demands.jobs.job.Job the hot path and its boundary read."""

    def __init__(self, graph, job_id, num_training_steps, frac, details, original_job=None):
        self.computation_graph = graph
        self.job_id = job_id
        self.num_training_steps = num_training_steps
        self.max_acceptable_job_completion_time_frac = frac
        self.details = details
        self.original_job = original_job if original_job is not None else self
        self.job_total_operation_memory_cost = details['job_total_op_memory_cost']
        self.job_total_dependency_size = details['job_total_dep_size']

    def register_job_arrived(self, time_arrived, job_idx):       # JOB:394-410
        self.details['time_arrived'] = time_arrived
        self.details['time_started'] = None
        self.details['time_completed'] = None
        self.details['job_idx'] = job_idx
        self.original_job.details['job_idx'] = job_idx

    def register_job_running(self, time_started):                # JOB:412-419
        self.details['time_started'] = time_started

    def register_job_completed(self, time_completed):            # JOB:421-430
        self.details['time_completed'] = time_completed


class _Part:
    def __init__(self, action, job_ids):
        self.action = action
        self.job_ids = set(job_ids)


class SyntheticAction:
    """action.py:3-69: ``actions`` (defaultdict -> None), ``job_ids`` = intersection over the parts."""

    def __init__(self, **parts):
        self.actions = defaultdict(lambda: None)
        for k, v in parts.items():
            if v is not None:
                self.actions[k] = v
        if len(self.actions) > 0:
            self.job_ids = set(set.intersection(*[p.job_ids for p in self.actions.values()]))
        else:
            self.job_ids = {}
        self.job_idxs = set()


def op_ids_of(lj: LoweredJob):
    return lj.op_ids if lj.op_ids is not None else [f'{i:07d}' for i in range(lj.n_ops)]


def build_original_job(job_id, model, orig_op_mem, orig_dep_size, frac, seq_time, num_training_steps):
    g = nx.MultiDiGraph()
    g.add_node('0', compute_cost={'A100': seq_time / max(num_training_steps, 1)}, memory_cost=orig_op_mem)
    details = {'model': model, 'job_total_op_memory_cost': orig_op_mem, 'job_total_dep_size': orig_dep_size,
               'job_sequential_completion_time': {'A100': seq_time},
               'max_acceptable_job_completion_time': {'A100': frac * seq_time},
               'mounted_workers': set(), 'mounted_channels': set()}
    return SyntheticJob(g, job_id, num_training_steps, frac, details)


def build_action(lj: LoweredJob, original_job: SyntheticJob, cluster, worker_ids=None, channel_ids=None):
    """Job-like partitioned job + the five action parts whose lowering is exactly ``lj``."""
    job_id = original_job.job_id
    ops = op_ids_of(lj)
    worker_ids = worker_ids or lj.worker_ids
    if worker_ids is None:
        worker_ids = sorted(cluster.topology.graph.graph['worker_to_node'].keys())[:lj.n_workers]
    w2n = cluster.topology.graph.graph['worker_to_node']
    g = nx.MultiDiGraph()
    mem_each = lj.mount.part_op_mem / max(lj.n_ops, 1)
    for i, op in enumerate(ops):
        g.add_node(op, compute_cost={'A100': float(lj.op_cost[i])}, memory_cost=0.0, pass_type='forward_pass')
    # spread the op memory so that the per-worker capacity check (A100.py:43) sees the real total
    for op in ops:
        g.nodes[op]['memory_cost'] = mem_each
    src = np.repeat(np.arange(lj.n_ops), np.diff(lj.row_ptr))
    n_flow = int(lj.dep_is_flow.sum())
    flow_size_each = lj.mount.flow_size / n_flow if n_flow else 0.0
    n_nonflow_same = 0
    placement = {op: worker_ids[int(lj.op_worker[i])] for i, op in enumerate(ops)}
    if channel_ids is None:
        channel_ids = lj.channel_ids
    chan_of_local = {}
    dep_place = defaultdict(set)
    dep_sched = defaultdict(lambda: defaultdict(dict))
    for e in range(lj.n_deps):
        u, v = ops[int(src[e])], ops[int(lj.dep_dst[e])]
        same = w2n[placement[u]] == w2n[placement[v]]
        if lj.dep_is_flow[e]:
            if same:
                raise Exception('a flow between ops on the same server cannot be expressed')
            size = flow_size_each
        else:
            size = 0.0 if not same else 1.0
            n_nonflow_same += same
        g.add_edge(u, v, key=0, size=size, init_run_time=float(lj.dep_run_time[e]) if lj.dep_is_flow[e] else 0.0)
        c = int(lj.dep_channel[e])
        if c != NO_CHANNEL:
            if c not in chan_of_local:
                if channel_ids is not None:
                    chan_of_local[c] = channel_ids[c]
                else:
                    chan_of_local[c] = f'src_{w2n[placement[u]]}_dst_{w2n[placement[v]]}_channel_0'
            cid = chan_of_local[c]
            dep_place[(u, v, 0)].add(cid)
            dep_sched[cid][job_id][(u, v, 0)] = int(lj.dep_prio[e])
        else:
            dep_place[(u, v, 0)].add(None)
    details = {'model': original_job.details['model'], 'job_idx': original_job.details.get('job_idx'),
               'job_total_op_memory_cost': lj.mount.part_op_mem, 'job_total_dep_size': lj.mount.part_dep_size,
               'job_sequential_completion_time': dict(original_job.details['job_sequential_completion_time']),
               'max_acceptable_job_completion_time': {'A100': lj.mount.max_acceptable_jct},
               'max_partitions_per_op': lj.degree, 'mounted_workers': set(), 'mounted_channels': set(),
               'time_arrived': original_job.details.get('time_arrived')}
    pjob = SyntheticJob(g, job_id, lj.num_training_steps, original_job.max_acceptable_job_completion_time_frac, details,
                        original_job=original_job)
    op_partition = _Part({job_id: {op: 1 for op in ops}}, [job_id])
    op_partition.partitioned_jobs = {job_id: pjob}
    op_partition.original_jobs = {job_id: original_job}
    op_partition.job_id_to_max_partition_degree = defaultdict(lambda: 1, {job_id: lj.degree})
    op_placement = _Part({job_id: placement}, [job_id])
    sched = defaultdict(lambda: defaultdict(dict))
    for i, op in enumerate(ops):
        sched[placement[op]][job_id][op] = int(lj.op_prio[i])
    op_schedule = _Part(sched, [job_id])
    dep_placement = _Part({job_id: dep_place}, [job_id])
    dep_schedule = _Part(dep_sched, [job_id])
    return SyntheticAction(op_partition=op_partition, op_placement=op_placement, op_schedule=op_schedule,
                           dep_placement=dep_placement, dep_schedule=dep_schedule), pjob


class SyntheticJobsGenerator:
    """JobsGenerator stand-in (jobs_generator.py:64-333): a fixed list of jobs and inter-arrival gaps."""

    def __init__(self, jobs, gaps):
        self._jobs, self._gaps = list(jobs), list(gaps)
        self.jobs_params = {}

    def __len__(self):
        return len(self._jobs)

    def sample_job(self):
        return self._jobs.pop(0)

    def sample_interarrival_time(self, size=None):
        if len(self._jobs) == 0:
            return float('inf')           # jobs_generator.py:270-272
        return self._gaps.pop(0)
