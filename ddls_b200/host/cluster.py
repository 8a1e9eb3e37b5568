"""Drop-in ``RampClusterEnvironment`` backed by the CUDA engine.

Same constructor / ``reset`` / ``step`` / ``is_done`` signatures and the same caller-visible state as the reference
(ddls/environments/ramp_cluster/ramp_cluster_environment.py, "RCE"): ``job_queue``, ``jobs_running / jobs_completed /
jobs_blocked``, ``stopwatch``, ``step_counter``, ``step_stats``, ``steps_log``, ``episode_stats``, ``job_op_placement``,
``job_dep_placement``, ``job_id_to_job_idx``, per-worker ``mounted_job_idx_to_ops`` / ``memory_occupied`` and per-channel
``mounted_job_idx_to_deps`` -- the agents (which stay the reference's Python) read these.

What moves to the GPU is exactly the hot path: the memoised lookahead tick loop (RCE:379-518), the lookahead
registration (RCE:793-888), the outer arrival/completion event loop and the step statistics (RCE:942-1106).  The host
side keeps the dict/set bookkeeping of mount / unmount (RCE:1305-1464) because the agents read it, lowers each new
Action to flat arrays (ddls_b200/lowering.py, cached by content), and replays the engine's events into the
reference's ``episode_stats`` lists.  There is NO CPU fallback: without the CUDA library this class raises.

Random-stream note: the reference samples the next job (and its inter-arrival gap) when that job arrives, inside
``step`` (RCE:1026 -> RCE:351-377).  The engine needs the gap before the step that contains the arrival, so the next
job is drawn at the start of the ``step`` in which it may arrive -- after the agents have produced that step's Action,
exactly one draw ahead; in the RampJobPartitioningEnvironment loop (RJPE:300-420) this is the same draw order.
"""
from __future__ import annotations

import copy
from collections import defaultdict

import numpy as np

from .. import engine as _engine
from ..lowering import ModelRegistry, lower_job
from ._classes import A100, JobQueue
from .devices import gen_job_dep_str
from .topology import Ramp

SS = _engine.SS


class Stopwatch:
    """ddls/utils.py:485-496 view of the engine clock."""

    def __init__(self):
        self._time = 0

    def reset(self):
        self._time = 0

    def tick(self, tick=1):
        self._time += tick

    def time(self):
        return self._time


class RampClusterEnvironment:
    def __init__(self, topology_config: dict, node_config: dict, name: str = 'ramp_cluster', path_to_save: str = None,
                 save_freq: int = 1, use_sqlite_database: bool = False, suppress_warnings=False, machine_epsilon=1e-7,
                 device: int = 0, max_jobs: int = 4096, jobs_generator_cls=None):
        self.suppress_warnings = suppress_warnings
        self.topology_config, self.node_config = topology_config, node_config
        self.name = name
        self.path_to_save = None          # log saving (RCE:1566-1598) is orchestration, not hot path: not mirrored
        self.save_freq = save_freq
        self.machine_epsilon = machine_epsilon
        self.device = device
        self._max_jobs = max_jobs
        self._jobs_generator_cls = jobs_generator_cls
        if topology_config['type'] != 'ramp':
            raise Exception(f'Unrecognised topology type {topology_config["type"]}. May need to implement here.')   # RCE:161
        self.topology = Ramp(**topology_config['kwargs'])
        num_nodes = sum(node_config[t]['num_nodes'] for t in node_config)
        if num_nodes != len(self.topology.graph.nodes):                                                               # RCE:164-167
            raise Exception(f'topology_config generated a topology with {len(self.topology.graph.nodes)} nodes, but node_config '
                            f'specified a total of {num_nodes} nodes, which is incompatible.')
        self._populate_topology(self.topology, node_config)
        self.stopwatch = Stopwatch()
        self.reset_counter = 0
        self._engine = None
        self._models = ModelRegistry()
        self._template_cache = {}

    # ---- RCE:169-198 ----
    def _populate_topology(self, topology, node_config):
        node_ids = iter(list(topology.graph.nodes))
        g = topology.graph.graph
        g['worker_to_node'], g['worker_to_type'], g['worker_types'], g['num_workers'] = {}, {}, set(), 0
        for node_type in node_config.keys():
            for _ in range(node_config[node_type]['num_nodes']):
                node_id = next(node_ids)
                topology.graph.nodes[node_id]['workers'] = {}
                for wc in node_config[node_type]['workers_config']:
                    if wc['num_workers'] > 1:
                        raise Exception('ERROR: Current RAMP implementation only supports 1 worker per server. Set worker_config["num_workers"] = 1.')
                    for i in range(wc['num_workers']):
                        Worker = wc['worker']
                        if isinstance(Worker, str):
                            Worker = A100 if Worker.endswith('A100') else None
                            if Worker is None:
                                raise Exception(f'Unknown worker class path {wc["worker"]}')
                        worker = Worker(processor_id=f'node_{node_id}_worker_{i}')
                        topology.graph.nodes[node_id]['workers'][worker.processor_id] = worker
                        g['worker_to_node'][worker.processor_id] = node_id
                        g['worker_to_type'][worker.processor_id] = worker.device_type
                        g['num_workers'] += 1
                        g['worker_types'].add(worker.device_type)

    # ---- RCE:202-295 ----
    def reset(self, jobs_config, max_simulation_run_time=float('inf'), job_queue_capacity: int = 10, seed: int = None,
              verbose=False):
        self.reset_counter += 1
        if seed is not None:
            raise Exception('RampClusterEnvironment.reset(seed=<int>) crashes in the reference (RCE:218 vs utils.py:20); seed the modules yourself.')
        self.seed = seed
        self.stopwatch.reset()
        if isinstance(jobs_config, dict):
            cls = self._jobs_generator_cls
            if cls is None:
                from ddls.demands.jobs.jobs_generator import JobsGenerator as cls      # ingest stays the reference's
            self.jobs_generator = cls(**jobs_config)
        else:
            self.jobs_generator = jobs_config                                          # an already built generator (duck-typed)
        self.max_simulation_run_time = max_simulation_run_time
        self._reset_steps_log()
        self.sim_log = defaultdict(list)
        self.episode_stats = self._init_episode_stats()
        for node_id in self.topology.graph.nodes:
            for worker in self.topology.graph.nodes[node_id]['workers'].values():
                worker.reset()
        for channel in self.topology.channel_id_to_channel.values():
            channel.reset()
        self.job_queue = JobQueue(queue_capacity=job_queue_capacity)
        self.num_jobs_arrived = 0
        self.num_mounted_ops = self.num_mounted_deps = 0
        self.load_rates = []
        self.mounted_workers, self.mounted_channels = set(), set()
        self.jobs_running, self.jobs_completed, self.jobs_blocked = {}, {}, {}
        self.job_op_to_worker = {}
        self.job_dep_to_channels = defaultdict(set)
        self.job_idx_to_job_id, self.job_id_to_job_idx = {}, {}
        self.step_counter = 0
        self.action = None
        self.job_model_to_max_num_partitions_to_init_details = defaultdict(lambda: defaultdict(lambda: defaultdict(lambda: None)))
        self.job_op_placement, self.job_dep_placement = {}, {}
        self._jobs_by_idx = {}
        self._pending = None           # (job, gap) drawn one ahead of its arrival
        self._done = False

        # engine: ONE episode, created once per environment and re-used by every reset(); ramp_reset clears the memo tables
        # like RCE:269-275 (registered templates are immutable and stay)
        n_workers = self.topology.graph.graph['num_workers']
        if self._engine is None:
            self._engine = _engine.RampEngine(
                n_episodes=1, n_cluster_workers=n_workers, max_jobs=self._max_jobs, device=self.device,
                job_queue_capacity=job_queue_capacity, machine_epsilon=self.machine_epsilon,
                max_simulation_run_time=float(max_simulation_run_time), memo_mode=_engine.MEMO_REFERENCE)
            # a step's outer loop runs once per job completion (<= one running job per worker) plus the arrival / the end of time
            self._engine.enable_tick_lists(n_workers + 16)
            self._template_cache = {}
        else:
            self._engine.set_limits(float(max_simulation_run_time), job_queue_capacity)
        # first job (RCE:280-281).  The arrival stream is fed one job ahead of its arrival (see step()); the engine's
        # "len(jobs_generator) > 0" is `job count - arrived > 0`, kept exact with ramp_set_job_count, so generators that never
        # run dry ('remove_and_repeat', 'replace') keep producing arrivals until max_simulation_run_time like the reference.
        self.time_next_job_to_arrive = 0
        job, gap = self._draw_job()
        rows = np.zeros((1, 1), dtype=_engine.ARRIVAL_DTYPE)
        rows[0, 0] = self._arrival_row(job, gap)
        self._engine.reset(rows)
        self._register_arrival(job, gap)
        self.job_queue.add(job)
        self._rec_prev = self._engine.job_records()[0].copy()
        return None

    def _reset_steps_log(self):
        self.steps_log = defaultdict(list)

    def _init_episode_stats(self):       # RCE:340-349
        es = defaultdict(list)
        es['num_jobs_arrived'] = es['num_jobs_completed'] = es['num_jobs_blocked'] = 0
        es['episode_start_time'] = copy.copy(self.stopwatch.time())
        return es

    # ---- arrivals (RCE:351-377) ----
    def _draw_job(self):
        job = self.jobs_generator.sample_job()
        gap = self.jobs_generator.sample_interarrival_time(size=None)
        return job, gap

    @staticmethod
    def _arrival_row(job, gap):
        return (float(gap), float(job.original_job.details['job_total_op_memory_cost']),
                float(job.original_job.details['job_total_dep_size']))

    def _register_arrival(self, job, gap):
        job_idx = copy.copy(self.num_jobs_arrived)
        job.original_job.job_id = job.job_id
        job.original_job.details['job_idx'] = job_idx
        job.register_job_arrived(time_arrived=self.stopwatch.time(), job_idx=job_idx)
        self.time_last_job_arrived = copy.copy(self.stopwatch.time())
        self.time_next_job_to_arrive += gap
        self.load_rates.append((job.original_job.details['job_total_op_memory_cost'] + job.original_job.details['job_total_dep_size'])
                               / (self.time_next_job_to_arrive - self.time_last_job_arrived))
        if job_idx in self.job_idx_to_job_id:
            raise Exception(f'job idx {job_idx} is already in arrived job_idx_to_job_id and is therefore not unique.')
        self.job_idx_to_job_id[job_idx] = job.job_id
        if job.job_id in self.job_id_to_job_idx:
            raise Exception(f'job id {job.job_id} is already in arrived job_id_to_job_idx and is therefore not unique.')
        self.job_id_to_job_idx[job.job_id] = job_idx
        self._jobs_by_idx[job_idx] = job
        self.num_jobs_arrived += 1
        self.last_job_arrived_job_idx = job_idx
        self.episode_stats['num_jobs_arrived'] += 1

    # ---- RCE:894-1179 ----
    def step(self, action, verbose: bool = False):
        if self._engine is None:
            raise Exception('reset() must be called before step()')
        self.action = action
        eng = self._engine
        # draw the next job one ahead (see module docstring), stream its arrival row to the engine and tell it whether the
        # generator still holds a job (RCE:1019-1040)
        if self._pending is None and len(self.jobs_generator) > 0:
            if self.num_jobs_arrived >= self._max_jobs:
                raise Exception(f'more than max_jobs={self._max_jobs} arrivals in one episode: construct the environment with a larger max_jobs')
            job, gap = self._draw_job()
            self._pending = (job, gap)
            eng.set_arrivals(0, self.num_jobs_arrived, np.array([self._arrival_row(job, gap)], dtype=_engine.ARRIVAL_DTYPE))
        eng.set_job_count(0, self.num_jobs_arrived + (1 if self._pending is not None else 0))

        job_ids = list(action.job_ids)
        if len(job_ids) > 1:
            raise Exception('More than one job per Action is not supported (the reference assumes one, action.py:37).')
        newly_blocked_host = []
        for job_id, job in list(self.job_queue.jobs.items()):            # RCE:914-919
            if job_id not in action.job_ids:
                newly_blocked_host.append(job)
        actions = eng.make_actions()
        mounted_job = None
        if len(job_ids) == 1:
            job_id = job_ids[0]
            parts = action.actions
            self._partition_ops(parts['op_partition'])
            self._place_ops(parts['op_placement'])
            self._schedule_ops(parts['op_schedule'])
            self._place_deps(parts['dep_placement'])
            self._schedule_deps(parts['dep_schedule'])
            lj = lower_job(self, action, job_id, self._models)
            key = (lj.fingerprint(), lj.model_id, lj.degree)
            tid = self._template_cache.get(key)
            if tid is None:
                tid = eng.register_template(lj)
                self._template_cache[key] = tid
            _engine.action_row(actions, 0, tid, lj.mount)
            mounted_job = self.jobs_running[self.job_id_to_job_idx[job_id]]
            mounted_job._lowered = lj

        before = self._rec_prev                                           # records only change inside a step
        stats = eng.step(actions)[0]
        eng.check_status()                                                # raises like RCE:462
        after = eng.job_records()[0].copy()
        self._rec_prev = after
        self._replay(stats, before, after, mounted_job, newly_blocked_host)
        done = bool(stats[SS['done']])
        self._done = done
        return None, None, None, done, None

    def is_done(self, verbose=False):     # RCE:1542-1557
        return self._done

    # ---- host-side mount bookkeeping (RCE:1285-1415), kept because the agents read it ----
    def _partition_ops(self, op_partition):
        self.op_partition = op_partition
        for job_id in op_partition.action:
            self.job_queue.jobs[job_id] = op_partition.partitioned_jobs[job_id]

    def _place_ops(self, op_placement):
        for job_id in op_placement.action:
            job = self.job_queue.jobs[job_id]
            for op_id in op_placement.action[job_id]:
                worker_id = op_placement.action[job_id][op_id]
                node_id = self.topology.graph.graph['worker_to_node'][worker_id]
                worker = self.topology.graph.nodes[node_id]['workers'][worker_id]
                if job.details['job_idx'] not in worker.mounted_job_idx_to_ops and len(worker.mounted_job_idx_to_ops) > 0:
                    raise Exception(f'Placement for job index {job.details["job_idx"]} job ID {job_id} op ID {op_id} worker ID '
                                    f'{worker_id} breaks the following Ramp rules: [\'one_job_per_worker\'].')      # RCE:1326-1328
                worker.mount(job=job, op_id=op_id)
                job.details['mounted_workers'].add(worker_id)
                self.num_mounted_ops += 1
                self.job_op_to_worker[gen_job_dep_str(job.details['job_idx'], job.job_id, op_id)] = worker_id
            job.register_job_running(time_started=self.stopwatch.time())      # RCE:1417-1420
            self.jobs_running[job.details['job_idx']] = job
            self.job_queue.remove(job)
            self.job_op_placement[job_id] = op_placement.action[job_id]

    def _schedule_ops(self, op_schedule):
        for worker_id in op_schedule.action.keys():
            node_id = self.topology.graph.graph['worker_to_node'][worker_id]
            worker = self.topology.graph.nodes[node_id]['workers'][worker_id]
            for job_idx in sorted(worker.mounted_job_idx_to_ops.keys()):
                job = self.jobs_running[job_idx]
                for op_id in sorted(worker.mounted_job_idx_to_ops[job_idx]):
                    worker.mounted_job_op_to_priority[gen_job_dep_str(job_idx, job.job_id, op_id)] = op_schedule.action[worker_id][job.job_id][op_id]

    def _place_deps(self, dep_placement):
        for job_id in dep_placement.action:
            job_idx = self.job_id_to_job_idx[job_id]
            job = self.jobs_running[job_idx]
            for dep_id in dep_placement.action[job_id].keys():
                for channel_id in dep_placement.action[job_id][dep_id]:
                    if channel_id is None:
                        continue
                    channel = self.topology.channel_id_to_channel[channel_id]
                    if job_idx not in channel.mounted_job_idx_to_deps and len(channel.mounted_job_idx_to_deps) > 0:
                        raise Exception(f'Dep placement for job index {job_idx} job ID {job_id} dep ID {dep_id} channel ID {channel_id} '
                                        f'breaks the following Ramp rules: [\'one_job_per_channel\']')                 # RCE:1367-1369
                    channel.mount(job, dep_id)
                    job.details['mounted_channels'].add(channel_id)
                    self.num_mounted_deps += 1
                    self.job_dep_to_channels[gen_job_dep_str(job_idx, job.job_id, dep_id)].add(channel_id)
            self.job_dep_placement[job_id] = dep_placement.action[job_id]

    def _schedule_deps(self, dep_schedule):
        for channel_id in dep_schedule.action.keys():
            if channel_id is None:
                continue
            channel = self.topology.channel_id_to_channel[channel_id]
            for job_idx in sorted(channel.mounted_job_idx_to_deps.keys()):
                job = self.jobs_running[job_idx]
                for dep_id in sorted(channel.mounted_job_idx_to_deps[job_idx]):
                    channel.mounted_job_dep_to_priority[gen_job_dep_str(job_idx, job.job_id, dep_id)] = dep_schedule.action[channel_id][job.job_id][dep_id]

    def _remove_job_from_cluster(self, job):       # RCE:1425-1464
        if job.job_id in self.job_queue.jobs:
            self.job_queue.remove(job)
        self.jobs_running.pop(job.details['job_idx'], None)
        for op_id in job.computation_graph.nodes:
            k = gen_job_dep_str(job.details['job_idx'], job.job_id, op_id)
            if k in self.job_op_to_worker:
                worker_id = self.job_op_to_worker.pop(k)
                node_id = self.topology.graph.graph['worker_to_node'][worker_id]
                self.topology.graph.nodes[node_id]['workers'][worker_id].unmount(job=job, op_id=op_id)
                self.num_mounted_ops -= 1
        for dep_id in job.computation_graph.edges:
            k = gen_job_dep_str(job.details['job_idx'], job.job_id, dep_id)
            if k in self.job_dep_to_channels:
                for channel_id in self.job_dep_to_channels.pop(k):
                    self.topology.channel_id_to_channel[channel_id].unmount(job, dep_id)
                    self.num_mounted_deps -= 1
        self.job_op_placement.pop(job.job_id, None)
        self.job_dep_placement.pop(job.job_id, None)

    # ---- replay the engine's events into the reference's Python state ----
    def _device_type(self):
        return list(self.topology.graph.graph['worker_types'])[0]

    def _register_blocked_job(self, job):          # RCE:1504-1540
        if job.job_id in self.job_queue.jobs:
            self.job_queue.remove(job)
        self.jobs_running.pop(job.details['job_idx'], None)
        if job.details['job_idx'] in self.jobs_blocked:
            return
        self.jobs_blocked[job.details['job_idx']] = job
        dt, es = self._device_type(), self.episode_stats
        es['num_jobs_blocked'] += 1
        es['jobs_blocked_num_nodes'].append(len(job.computation_graph.nodes))
        es['jobs_blocked_num_edges'].append(len(job.computation_graph.edges))
        es['jobs_blocked_total_operation_memory_cost'].append(job.job_total_operation_memory_cost)
        es['jobs_blocked_total_dependency_size'].append(job.job_total_dependency_size)
        es['jobs_blocked_job_sequential_completion_time'].append(job.details['job_sequential_completion_time'][dt])
        es['jobs_blocked_max_acceptable_job_completion_time_frac'].append(job.max_acceptable_job_completion_time_frac)
        es['jobs_blocked_max_acceptable_job_completion_time'].append(job.details['max_acceptable_job_completion_time'][dt])
        es['jobs_blocked_original_demand_num_nodes'].append(len(job.original_job.computation_graph.nodes))
        es['jobs_blocked_original_demand_num_edges'].append(len(job.original_job.computation_graph.edges))
        es['jobs_blocked_original_demand_total_operation_memory_cost'].append(job.original_job.job_total_operation_memory_cost)
        es['jobs_blocked_original_demand_total_dependency_size'].append(job.original_job.job_total_dependency_size)

    def _register_completed_job(self, job, rec):   # RCE:1466-1502
        job.register_job_completed(time_completed=float(rec['time_completed']))
        self.jobs_completed[job.details['job_idx']] = job
        dt, es = self._device_type(), self.episode_stats
        es['num_jobs_completed'] += 1
        jct = job.details['time_completed'] - job.details['time_arrived']
        es['job_completion_time'].append(jct)
        es['job_completion_time_speedup'].append(job.details['job_sequential_completion_time'][dt] / jct)
        es['job_communication_overhead_time'].append(job.details['communication_overhead_time'])
        es['job_computation_overhead_time'].append(job.details['computation_overhead_time'])
        es['jobs_completed_num_nodes'].append(len(job.computation_graph.nodes))
        es['jobs_completed_num_edges'].append(len(job.computation_graph.edges))
        es['jobs_completed_total_operation_memory_cost'].append(job.job_total_operation_memory_cost)
        es['jobs_completed_total_dependency_size'].append(job.job_total_dependency_size)
        es['jobs_completed_max_partitions_per_op'].append(job.details.get('max_partitions_per_op'))
        es['jobs_completed_job_sequential_completion_time'].append(job.details['job_sequential_completion_time'][dt])
        es['jobs_completed_max_acceptable_job_completion_time_frac'].append(job.max_acceptable_job_completion_time_frac)
        es['jobs_completed_max_acceptable_job_completion_time'].append(job.details['max_acceptable_job_completion_time'][dt])
        es['jobs_completed_num_mounted_workers'].append(len(job.details['mounted_workers']))
        es['jobs_completed_num_mounted_channels'].append(len(job.details['mounted_channels']))
        es['jobs_completed_mean_mounted_worker_utilisation_frac'].append(job.details['mean_mounted_worker_utilisation_frac'])
        es['jobs_completed_original_demand_num_nodes'].append(len(job.original_job.computation_graph.nodes))
        es['jobs_completed_original_demand_num_edges'].append(len(job.original_job.computation_graph.edges))
        es['jobs_completed_original_demand_total_operation_memory_cost'].append(job.original_job.job_total_operation_memory_cost)
        es['jobs_completed_original_demand_total_dependency_size'].append(job.original_job.job_total_dependency_size)
        self._remove_job_from_cluster(job)

    def _replay(self, stats, before, after, mounted_job, newly_blocked_host):
        JS = _engine
        step_stats = defaultdict(lambda: 0)
        for k in _engine.STEP_STATS:
            step_stats[k] = float(stats[SS[k]])
        for k in ('step_counter', 'num_jobs_completed', 'num_jobs_arrived', 'num_jobs_blocked', 'job_queue_length'):
            step_stats[k] = int(step_stats[k])
        # the reference leaves these two as per-tick lists (RCE:989-994): the engine kept every entry of the step
        tick_mounted, tick_cluster = self._engine.tick_lists(0)
        assert len(tick_mounted) == int(stats[SS['num_ticks']])
        step_stats['mean_mounted_worker_utilisation_frac'] = [float(x) for x in tick_mounted]
        step_stats['mean_cluster_worker_utilisation_frac'] = [float(x) for x in tick_cluster]
        self.step_stats = step_stats

        # 1. queued jobs the action did not handle (RCE:914-919)
        for job in newly_blocked_host:
            self._register_blocked_job(job)
        # 2. the mounted job: lookahead results and block decision (RCE:793-888)
        if mounted_job is not None:
            idx = mounted_job.details['job_idx']
            rec = after[idx]
            # a record that is BLOCKED with its lookahead results filled in was accepted (RCE:826-888) and then blocked because
            # the simulation ended in this very step with the job still running (RCE:1111-1121): handled with the events below
            if rec['status'] == JS.JS_BLOCKED and float(rec['jct']) == 0.0:
                self._register_blocked_job(mounted_job.original_job)
                self._remove_job_from_cluster(mounted_job)
            else:
                details = {'lookahead_job_completion_time': float(rec['jct']),
                           'communication_overhead_time': float(rec['comm']),
                           'computation_overhead_time': float(rec['comp']),
                           'mounted_workers': mounted_job.details['mounted_workers'],
                           'mounted_channels': mounted_job.details['mounted_channels'],
                           'mean_mounted_worker_utilisation_frac': float(rec['util'])}
                if hasattr(mounted_job, 'reset_job') and hasattr(self.op_partition, 'job_id_to_max_partition_degree'):
                    # RCE:848-879: reset the whole job for the actual simulation, re-using (and then refreshing) the
                    # per-(model, max partition degree) init details that OpPartition reads (op_partition.py:47-50)
                    job_id = mounted_job.job_id
                    mnp = self.op_partition.job_id_to_max_partition_degree[job_id]
                    model = mounted_job.details['model']
                    memo = self.job_model_to_max_num_partitions_to_init_details
                    tot_mem = tot_dep = imm = None
                    if model in memo and mnp in memo[model]:
                        tot_mem = memo[model][mnp]['job_total_operation_memory_cost']
                        tot_dep = memo[model][mnp]['job_total_dependency_size']
                        imm = memo[model][mnp]['init_job_immutable_details']
                    memo[model]
                    mounted_job.reset_job(details=details, job_total_operation_memory_cost=tot_mem,
                                          job_total_dependency_size=tot_dep, init_job_immutable_details=imm)
                    memo[model][mnp]['job_total_operation_memory_cost'] = mounted_job.job_total_operation_memory_cost
                    memo[model][mnp]['job_total_dependency_size'] = mounted_job.job_total_dependency_size
                    memo[model][mnp]['init_job_immutable_details'] = mounted_job.init_job_immutable_details
                    memo[model][mnp]['partitioned_computation_graph'] = self.op_partition.job_id_to_partitioned_computation_graph[job_id]
                else:
                    mounted_job.details.update(details)
                mounted_job.details['job_total_flow_size'] = mounted_job._lowered.mount.flow_size          # RCE:882-888
        # 3. events of the outer loop, in event order (RCE:1004-1037)
        n_blocked_at_sim_end = 0
        changed = [i for i in range(self.num_jobs_arrived) if after[i]['status'] != before[i]['status']
                   and after[i]['status'] in (JS.JS_COMPLETED, JS.JS_BLOCKED)]
        for i in sorted(changed, key=lambda i: int(after[i]['event_seq'])):
            job = self.jobs_running.get(i)
            if after[i]['status'] == JS.JS_COMPLETED and job is not None:
                self.stopwatch._time = float(after[i]['time_completed'])
                self._register_completed_job(job, after[i])
            elif after[i]['status'] == JS.JS_BLOCKED and job is not None:     # still running when the simulation ended (RCE:1111-1121)
                self._register_blocked_job(job.original_job)
                self._remove_job_from_cluster(job)
                n_blocked_at_sim_end += 1
        self.stopwatch._time = float(stats[SS['step_end_time']])
        if step_stats['num_jobs_arrived'] > 0:
            job, gap = self._pending
            self._pending = None
            self._register_arrival(job, gap)
            if self.job_queue.can_fit(job):
                self.job_queue.add(job)
            else:
                self._register_blocked_job(job)
        if len(self.jobs_generator) == 0 and self._pending is None:
            self.time_next_job_to_arrive = float('inf')                       # RCE:1040
        self.mounted_workers, self.mounted_channels = set(), set()
        for job in self.jobs_running.values():
            self.mounted_workers.update(job.details['mounted_workers'])
            self.mounted_channels.update(job.details['mounted_channels'])

        # logs (RCE:1082-1109)
        # the reference appends to steps_log BEFORE it blocks the jobs still running at the end of the simulation (RCE:1082-1090 vs
        # RCE:1111-1121), so the log's last num_jobs_blocked lacks them while step_stats has them
        for key, val in step_stats.items():
            self.steps_log[key].append(val - n_blocked_at_sim_end if key == 'num_jobs_blocked' else val)
        for metric in ('compute_info_processed', 'dep_info_processed', 'flow_info_processed', 'cluster_info_processed',
                       'demand_compute_info_processed', 'demand_dep_info_processed', 'demand_total_info_processed',
                       'mean_compute_overhead_frac', 'mean_communication_overhead_frac', 'mean_num_jobs_running',
                       'mean_num_mounted_workers', 'mean_mounted_worker_utilisation_frac', 'mean_cluster_worker_utilisation_frac'):
            self.episode_stats[metric].append(step_stats[metric])
        self.step_counter += 1
        if bool(stats[SS['done']]):
            self._finalise_episode()

    def _finalise_episode(self):                   # RCE:1123-1167
        es = self.episode_stats
        es['episode_end_time'] = copy.copy(self.stopwatch.time())
        es['episode_time'] = es['episode_end_time'] - es['episode_start_time']
        es['mean_load_rate'] = np.mean(self.load_rates)
        es['blocking_rate'] = es['num_jobs_blocked'] / es['num_jobs_arrived'] if es['num_jobs_arrived'] else 0
        es['acceptance_rate'] = es['num_jobs_completed'] / es['num_jobs_arrived'] if es['num_jobs_arrived'] else 0
        for tp, info in {'mean_compute_throughput': 'compute_info_processed', 'mean_dep_throughput': 'dep_info_processed',
                         'mean_flow_throughput': 'flow_info_processed', 'mean_cluster_throughput': 'cluster_info_processed',
                         'mean_demand_compute_throughput': 'demand_compute_info_processed',
                         'mean_demand_dep_throughput': 'demand_dep_info_processed',
                         'mean_demand_total_throughput': 'demand_total_info_processed'}.items():
            es[info] = np.sum(es[info])
            es[tp] = es[info] / es['episode_time'] if es[info] != 0 and es['episode_time'] != 0 else 0
        for m in ('mean_compute_overhead_frac', 'mean_communication_overhead_frac', 'mean_num_jobs_running',
                  'mean_num_mounted_workers', 'mean_mounted_worker_utilisation_frac', 'mean_cluster_worker_utilisation_frac'):
            vals = es[m]
            flat = [x for v in vals for x in (v if isinstance(v, list) else [v])]
            es[m] = float(np.mean(flat)) if len(flat) and es['episode_time'] != 0 else 0

    def __str__(self):
        return (f'Cluster {type(self)} | Topology: {type(self.topology)} with {len(self.topology.graph.nodes)} nodes '
                f'| Topology config: {self.topology_config} | Node config: {self.node_config}')
