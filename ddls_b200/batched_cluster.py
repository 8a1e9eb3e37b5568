"""B ``RampClusterEnvironment`` instances in lock step on ONE engine, driven with the reference's own ``Action`` objects.

``BatchedRampClusterEnvironment.step(actions)`` takes one ``Action`` per episode (``None`` or an empty ``Action()`` for "place
nothing", RCE:914-919) -- the reference's ``Action{OpPartition, OpPlacement, OpSchedule, DepPlacement, DepSchedule}`` classes or
anything with the same attributes -- and is, per episode, ``RampClusterEnvironment.step(action)`` (RCE:894-1179): the queued job the
action handles is mounted, its lookahead is run or recalled from the episode's memo (RCE:469-518), the job is registered or blocked
(RCE:793-888) and the cluster runs to the next event.  What comes back is the engine's view: the ``step_stats`` row of every
episode, job records, episode state.

Lowering an ``Action`` (``ddls_b200.lowering.lower_job``: tens of milliseconds of dict walking for a job with 10^5 deps) is done
once per (model, maximum partition degree, placement signature): the placement signature is the op -> worker map of the job, and
under the reference's schedulers and dep placer (SRPTOpScheduler / SRPTDepScheduler / FirstFitDepPlacer with one channel per server
pair -- the only ones RampJobPartitioningEnvironment supports, RJPE:171-197) priorities, channels and dep run times are functions
of it; an agent for which they are not sets ``action.lowering_key`` (any hashable) to tell such actions apart.  Pass
``verify_cache=True`` to lower every action anyway and check that cache hits are byte-identical (the tests do).

This class keeps no per-episode mirror of workers and channels for agents to read (``ddls_b200.host.RampClusterEnvironment`` is
the drop-in with that state, one episode per instance); ``workers_in_use(b)`` gives the occupancy a placer needs.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import engine as _engine
from .host.topology import Ramp
from .lowered import MountScalars
from .lowering import ModelRegistry, lower_job


class BatchedRampClusterEnvironment:
    def __init__(self, topology_config: dict, node_config: Optional[dict] = None, n_episodes: int = 1, max_jobs: int = 64,
                 device: int = 0, machine_epsilon: float = 1e-7, memo_mode: int = _engine.MEMO_REFERENCE, verify_cache: bool = False,
                 worker_type: str = 'A100'):
        if topology_config.get('type', 'ramp') != 'ramp':
            raise Exception(f'Unrecognised topology type {topology_config["type"]}')                 # RCE:148-151
        self.topology = Ramp(**topology_config['kwargs'])
        g = self.topology.graph.graph
        g['worker_to_node'], g['worker_to_type'] = {}, {}
        n_per_node = 1
        if node_config:
            cfgs = [c for t in node_config.values() for c in t['workers_config']]
            n_per_node = int(cfgs[0]['num_workers'])
            if n_per_node != 1 or len(cfgs) != 1:
                raise Exception('RAMP clusters have one worker per server (ramp_cluster_environment.py:169-198, ramp_rules.py)')
        for node in self.topology.graph.nodes:                                                       # RCE:169-198
            w = f'node_{node}_worker_0'
            g['worker_to_node'][w] = node
            g['worker_to_type'][w] = worker_type
        g['num_workers'] = len(g['worker_to_node'])
        self.B, self.J = int(n_episodes), int(max_jobs)
        self.eng = _engine.RampEngine(n_episodes=self.B, n_cluster_workers=g['num_workers'], max_jobs=self.J, device=device,
                                      machine_epsilon=machine_epsilon, memo_mode=memo_mode)
        self.verify_cache = verify_cache
        self._models = ModelRegistry()
        self._cache: Dict[tuple, tuple] = {}             # (model, degree, placement signature) -> (template id, LoweredJob)
        self._by_fingerprint: Dict[tuple, int] = {}      # identical lowered jobs reached through different placements share a template
        self._mounted: List[Dict[int, tuple]] = [dict() for _ in range(self.B)]   # episode -> {job idx: worker ids}
        self._mounts_fresh = True
        self.stats = {'actions': 0, 'lowerings': 0, 'cache_hits': 0, 'templates': 0}
        self.last_step_stats = None

    def close(self):
        self.eng.close()

    # ---- RCE:202-295 -------------------------------------------------------------------------------------------------
    def reset(self, arrivals, max_simulation_run_time: float = float('inf'), job_queue_capacity: int = 10):
        """arrivals: [B, J'] rows of (gap to the NEXT arrival, original op memory, original dep size) in arrival order, either as
        ``engine.ARRIVAL_DTYPE`` or as a float array [B, J', 3] (what ``JobsGenerator`` + ``Job`` hold: jobs_generator.py:278-333)."""
        arr = np.asarray(arrivals)
        if arr.dtype != _engine.ARRIVAL_DTYPE:
            a = np.zeros(arr.shape[:2], dtype=_engine.ARRIVAL_DTYPE)
            a['interarrival'], a['orig_op_mem'], a['orig_dep_size'] = arr[..., 0], arr[..., 1], arr[..., 2]
            arr = a
        if arr.shape[0] != self.B or arr.shape[1] > self.J:
            raise Exception(f'arrivals must be [{self.B}, <= {self.J}]')
        self.eng.set_limits(float(max_simulation_run_time), int(job_queue_capacity))
        self.eng.reset(np.ascontiguousarray(arr))
        self._mounted = [dict() for _ in range(self.B)]
        self._mounts_fresh = True
        self.last_step_stats = None

    # ---- the lowering cache ------------------------------------------------------------------------------------------
    def _signature(self, action, job_id):
        parts = action.actions
        job = parts['op_partition'].partitioned_jobs[job_id]
        placement = parts['op_placement'].action[job_id]
        model = job.details['model'] if 'model' in job.details else ''
        degree = int(parts['op_partition'].job_id_to_max_partition_degree[job_id])
        # `action.lowering_key` (optional, any hashable): for agents whose schedules / run times are NOT functions of the placement
        return (model, degree, tuple(sorted(placement.items(), key=lambda kv: str(kv[0]))), getattr(action, 'lowering_key', None)), job, placement

    def _template_for(self, action, job_id):
        key, job, placement = self._signature(action, job_id)
        hit = self._cache.get(key)
        if hit is not None and not self.verify_cache:
            self.stats['cache_hits'] += 1
            return hit[0], hit[1], job, placement
        lj = lower_job(self, action, job_id, self._models)
        self.stats['lowerings'] += 1
        if hit is not None:                               # verify_cache: the signature must determine the lowered job
            if lj.fingerprint() != hit[1].fingerprint():
                raise Exception('two actions with the same (model, degree, op placement) lowered to different jobs: the op / dep '
                                'schedules or channels are not functions of the placement; construct with a signature that covers them')
            self.stats['cache_hits'] += 1
            return hit[0], hit[1], job, placement
        fp = (lj.fingerprint(), lj.model_id, lj.degree)
        tid = self._by_fingerprint.get(fp)
        if tid is None:
            tid = self.eng.register_template(lj)
            self._by_fingerprint[fp] = tid
            self.stats['templates'] += 1
        self._cache[key] = (tid, lj)
        return tid, lj, job, placement

    # ---- RCE:894-1179 for every episode ------------------------------------------------------------------------------
    def step(self, actions: Sequence, fuse_empty_steps: bool = False) -> np.ndarray:
        """actions[b]: the Action for episode b's queued job, or None / an Action that handles no job.  Returns the step_stats rows
        [B, STEP_STATS_LEN] (columns: ``engine.SS``)."""
        if len(actions) != self.B:
            raise Exception(f'{len(actions)} actions for {self.B} episodes')
        rows = self.eng.make_actions()
        queued = None
        for b, action in enumerate(actions):
            if action is None:
                continue
            job_ids = list(getattr(action, 'job_ids', ()))
            if not job_ids:
                continue
            if len(job_ids) != 1:
                raise Exception(f'episode {b}: an Action handles {len(job_ids)} jobs; RampJobPartitioningEnvironment places one job per step (RJPE:300-343)')
            self.stats['actions'] += 1
            job_id = job_ids[0]
            tid, lj, job, placement = self._template_for(action, job_id)
            wtype = self.topology.graph.graph['worker_to_type'][next(iter(placement.values()))]
            mount = MountScalars(float(job.details['max_acceptable_job_completion_time'][wtype]),      # RCE:815
                                 float(job.details['job_total_op_memory_cost']), float(job.details['job_total_dep_size']),   # RCE:966-967
                                 float(lj.mount.flow_size), int(lj.mount.n_mounted_workers), int(lj.mount.n_mounted_channels))
            _engine.action_row(rows, b, tid, mount)
            if queued is None:
                queued = self.queued_job()
            self._mounted[b][int(queued[b])] = tuple(sorted(set(placement.values())))
        stats = self.eng.step(rows, fuse_empty_steps=fuse_empty_steps)
        self.eng.check_status()
        self._mounts_fresh = False
        self.last_step_stats = stats
        return stats

    # ---- state -------------------------------------------------------------------------------------------------------
    def queued_job(self) -> np.ndarray:
        """[B] index of the job waiting in each episode's queue, -1 if none (RCE:351-377: at most one job queues at a time)."""
        return self.eng.episode_state()[:, _engine.EP['queued_job']].astype(np.int64)

    def done(self) -> np.ndarray:
        return self.eng.episode_state()[:, _engine.EP['done']] != 0                                    # RCE:1542-1557

    def time(self) -> np.ndarray:
        return self.eng.episode_state()[:, _engine.EP['time']].copy()

    def job_records(self) -> np.ndarray:
        return self.eng.job_records()

    def workers_in_use(self, b: int) -> set:
        """Workers of episode b's running jobs (one job per worker: ramp_rules.py:6-39) -- what a placer must avoid."""
        if not self._mounts_fresh:
            rec = self.eng.job_records()
            for e in range(self.B):
                gone = [j for j in self._mounted[e] if rec[e, j]['status'] != _engine.JS_RUNNING]
                for j in gone:
                    del self._mounted[e][j]
            self._mounts_fresh = True
        return {w for ws in self._mounted[b].values() for w in ws}
