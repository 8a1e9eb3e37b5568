"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONHASHSEED=0 python oracle/gen_golden.py            # all cases
    PYTHONHASHSEED=0 python oracle/gen_golden.py chain8     # one case

For every episode it drives ``RampJobPartitioningEnvironment`` exactly like
``ddls/loops/eval_loop.py:35-42`` and records, through monkey-patched hooks on the
reference's ``RampClusterEnvironment``:

  * per ``step``: the lowered Action (``ddls_b200.lowering.lower_job`` applied to the
    reference's own Action / Job / topology objects) or -1 for ``Action()``, and the
    reference's ``step_stats``;
  * per un-memoised ``_run_lookahead``: ``(jct, comm, comp)`` and the complete
    ``tick_counter_to_active_workers_tick_size`` trace (RCE:467);
  * per arrival (``_get_next_job``): the inter-arrival gap drawn and the original job's
    totals (RCE:363-364);
  * at the end: ``episode_stats``.

The fixtures pin oracle/ramp_oracle.c (tests/test_oracle_golden.py) -- the reference has no
tests or golden vectors of its own.
"""
import os
import random
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402

ref_shim.install()

from ddls.environments.ramp_job_partitioning.ramp_job_partitioning_environment import RampJobPartitioningEnvironment  # noqa: E402
from ddls.environments.ramp_cluster.ramp_cluster_environment import RampClusterEnvironment  # noqa: E402
from ddls.devices.processors.gpus.A100 import A100  # noqa: E402
from ddls.distributions.fixed import Fixed  # noqa: E402
from ddls.distributions.distribution import Distribution  # noqa: E402
from ddls.distributions.uniform import Uniform  # noqa: E402
from ddls.environments.ramp_job_partitioning.agents.sip_ml import SiPML  # noqa: E402
from ddls.environments.ramp_job_partitioning.agents.random import Random  # noqa: E402
from ddls.environments.ramp_job_partitioning.agents.acceptable_jct import AcceptableJCT  # noqa: E402

from ddls_b200 import synth  # noqa: E402
from ddls_b200.lowering import lower_job, ModelRegistry  # noqa: E402
from oracle.oracle import STEP_STATS  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

STEP_KEYS_SCALAR = [k for k in STEP_STATS if k not in ('util_mounted_sum', 'util_cluster_sum', 'num_ticks', 'done',
                                                        'lookahead_ran')]


class Recorder:
    """Hooks on the reference RampClusterEnvironment (class-level monkey patches)."""

    def __init__(self):
        self.models = ModelRegistry()
        self.templates = []          # unique LoweredJob by fingerprint
        self.fp_to_tid = {}
        self.steps = []              # dict per cluster.step
        self.lookaheads = []         # dict per un-memoised _run_lookahead
        self.arrivals = []           # (interarrival, orig_op_mem, orig_dep_size) per job idx
        self._orig = {}

    def install(self):
        rec = self
        self._orig['step'] = RampClusterEnvironment.step
        self._orig['_run_lookahead'] = RampClusterEnvironment._run_lookahead
        self._orig['_get_next_job'] = RampClusterEnvironment._get_next_job

        def step(cluster, action, verbose=False):
            entry = {'tid': -1, 'mount': None, 'lookahead': -1}
            job_ids = list(action.job_ids)
            if len(job_ids) > 1:
                raise Exception('golden recorder assumes <= 1 job per step')
            rec._current = entry
            if len(job_ids) == 1:
                # lower BEFORE the step mutates the cluster (placement is all in the Action)
                lj = lower_job(cluster, action, job_ids[0], rec.models)
                fp = (lj.fingerprint(), lj.model_id, lj.degree)
                if fp not in rec.fp_to_tid:
                    rec.fp_to_tid[fp] = len(rec.templates)
                    rec.templates.append(lj)
                entry['tid'] = rec.fp_to_tid[fp]
                entry['mount'] = lj.mount
            out = rec._orig['step'](cluster, action, verbose=verbose)
            ss = cluster.step_stats
            vec = {}
            for k in STEP_KEYS_SCALAR:
                vec[k] = float(ss[k]) if k in ss else 0.0
            lst = ss['mean_mounted_worker_utilisation_frac']
            vec['util_mounted_sum'] = float(np.sum(lst)) if isinstance(lst, list) else float(lst)
            lst2 = ss['mean_cluster_worker_utilisation_frac']
            vec['util_cluster_sum'] = float(np.sum(lst2)) if isinstance(lst2, list) else float(lst2)
            vec['num_ticks'] = float(len(lst)) if isinstance(lst, list) else 0.0
            vec['done'] = float(out[3])
            vec['lookahead_ran'] = 1.0 if entry['lookahead'] >= 0 else 0.0
            entry['stats'] = vec
            entry['time'] = float(cluster.stopwatch.time())
            rec.steps.append(entry)
            return out

        def _run_lookahead(cluster, job_id, *args, **kwargs):
            out = rec._orig['_run_lookahead'](cluster, job_id, *args, **kwargs)
            job, jct, comm, comp, trace = out
            ticks = sorted(trace.keys())
            assert ticks == list(range(1, len(ticks) + 1))
            rec._current['lookahead'] = len(rec.lookaheads)
            rec.lookaheads.append({'tid': rec._current['tid'], 'jct': float(jct), 'comm': float(comm), 'comp': float(comp),
                                   'trace_n': np.array([trace[t][0] for t in ticks], dtype=np.int32),
                                   'trace_tick': np.array([trace[t][1] for t in ticks], dtype=np.float64)})
            return out

        def _get_next_job(cluster):
            before = cluster.time_next_job_to_arrive
            job = rec._orig['_get_next_job'](cluster)
            gap = cluster.time_next_job_to_arrive - before
            rec.arrivals.append((float(gap), float(job.original_job.details['job_total_op_memory_cost']),
                                 float(job.original_job.details['job_total_dep_size'])))
            return job

        RampClusterEnvironment.step = step
        RampClusterEnvironment._run_lookahead = _run_lookahead
        RampClusterEnvironment._get_next_job = _get_next_job

    def uninstall(self):
        for k, f in self._orig.items():
            setattr(RampClusterEnvironment, k, f)


class Exponential(Distribution):
    """Exponential inter-arrival times (BASELINE config 5); the reference ships no such distribution, its JobsGenerator only
    calls .sample() (jobs_generator.py).  Drawn from numpy's global generator like the reference's own distributions, so the
    episode stays a function of the seed."""

    def __init__(self, mean):
        self.mean = float(mean)

    def sample(self, size=None):
        return float(np.random.exponential(self.mean)) if size is None else np.random.exponential(self.mean, size=size)


def make_env(graph_dir, shape, n_jobs, max_partitions, interarrival, frac_dist, max_sim_time=1e6, quantum=0.01,
             num_training_steps=50, sampling_mode='remove', num_channels=1):
    c, r, s = shape
    interarrival_dist = Exponential(interarrival[1]) if isinstance(interarrival, tuple) else Fixed(val=interarrival)
    return RampJobPartitioningEnvironment(
        topology_config={'type': 'ramp', 'kwargs': {'num_communication_groups': c, 'num_racks_per_communication_group': r,
                                                    'num_servers_per_rack': s, 'num_channels': num_channels,
                                                    'total_node_bandwidth': 1.6e12,
                                                    'intra_gpu_propagation_latency': 50e-9, 'worker_io_latency': 100e-9}},
        node_config={'type_1': {'num_nodes': c * r * s, 'workers_config': [{'num_workers': 1, 'worker': A100}]}},
        jobs_config={'path_to_files': graph_dir, 'job_interarrival_time_dist': interarrival_dist,
                     'max_acceptable_job_completion_time_frac_dist': frac_dist, 'replication_factor': n_jobs,
                     'job_sampling_mode': sampling_mode, 'num_training_steps': num_training_steps, 'shuffle_files': True},
        max_partitions_per_op=max_partitions, min_op_run_time_quantum=quantum, reward_function='job_acceptance',
        reward_function_kwargs={'fail_reward': -1, 'success_reward': 1}, pad_obs_kwargs={'max_nodes': 400},
        max_simulation_run_time=max_sim_time, suppress_warnings=True)


CASES = {
    # name: (graphs, shape, n_jobs per graph, max_partitions, interarrival, frac (lo, hi), actor, seed, extra kwargs)
    'chain8': dict(graphs=[synth.chain_graph(6, 'chain6')], shape=(2, 2, 2), n_jobs=20, max_partitions=8,
                   interarrival=1000.0, frac=(0.1, 1.0), actor='random', seed=0),
    'chain8_busy': dict(graphs=[synth.chain_graph(6, 'chain6')], shape=(2, 2, 2), n_jobs=24, max_partitions=8,
                        interarrival=150.0, frac=(0.1, 1.0), actor='random', seed=5),
    'chain8_maxtime': dict(graphs=[synth.chain_graph(6, 'chain6')], shape=(2, 2, 2), n_jobs=12, max_partitions=4,
                           interarrival=400.0, frac=(0.5, 1.0), actor='sipml', seed=2, max_sim_time=2500.0),
    'residual8_deg4': dict(graphs=[synth.residual_small_graph()], shape=(2, 2, 2), n_jobs=6, max_partitions=4,
                           interarrival=1000.0, frac=(0.1, 1.0), actor='sipml', seed=1),
    'mixed16': dict(graphs=[synth.chain_graph(5, 'chain5'), synth.resnet_like_graph(n_blocks=2, stem=2, name='res2', seed=7, body_per_block=3),
                            synth.transformer_like_graph(n_layers=1, name='tfm1', seed=4)],
                    shape=(2, 2, 4), n_jobs=5, max_partitions=8, interarrival=300.0, frac=(0.1, 1.0), actor='random', seed=3),
    'tfm32_acceptable': dict(graphs=[synth.transformer_like_graph(n_layers=2, name='tfm2', seed=9)], shape=(4, 4, 2), n_jobs=6,
                             max_partitions=16, interarrival=600.0, frac=(0.05, 0.6), actor='acceptable_jct', seed=4),
    # many jobs in flight on a 64-worker cluster, three models sharing one memo, placements that fail when the cluster is full
    'mixed64_busy': dict(graphs=[synth.chain_graph(4, 'chain4'), synth.resnet_like_graph(n_blocks=1, stem=2, name='res1', seed=11, body_per_block=2),
                                 synth.transformer_like_graph(n_layers=1, name='tfm1b', seed=6)],
                         shape=(4, 4, 4), n_jobs=8, max_partitions=8, interarrival=60.0, frac=(0.2, 1.0), actor='random', seed=12),
    # arrivals far faster than completions: most jobs are blocked (no free workers / JCT above the acceptable one)
    'res16_flood': dict(graphs=[synth.resnet_like_graph(n_blocks=1, stem=1, name='res1s', seed=3, body_per_block=2)], shape=(2, 2, 4),
                        n_jobs=30, max_partitions=4, interarrival=20.0, frac=(0.3, 1.0), actor='sipml', seed=8),
    # BASELINE.json's job at full size: the bench's ResNet-50-like graph at degree 16 on the 64-worker cluster
    # (N=5,280 ops, E=132,016 deps, T=1,169 ticks) -- one job, one lookahead of the reference itself
    'resnet64_deg16_full': dict(graphs=[synth.resnet_like_graph()], shape=(4, 4, 4), n_jobs=1, max_partitions=16,
                                interarrival=1000.0, frac=(1.0, 1.0), actor='sipml', seed=1),
    'resnet64_deg8_full': dict(graphs=[synth.resnet_like_graph()], shape=(4, 4, 4), n_jobs=1, max_partitions=8,
                               interarrival=1000.0, frac=(1.0, 1.0), actor='sipml', seed=1),
    'resnet64_deg4_full': dict(graphs=[synth.resnet_like_graph()], shape=(4, 4, 4), n_jobs=1, max_partitions=4,
                               interarrival=1000.0, frac=(1.0, 1.0), actor='sipml', seed=1),
    'resnet64_deg2_full': dict(graphs=[synth.resnet_like_graph()], shape=(4, 4, 4), n_jobs=1, max_partitions=2,
                               interarrival=1000.0, frac=(1.0, 1.0), actor='sipml', seed=1),
    # BASELINE configs 4 and 5 at sizes the Python reference finishes in about a minute: clusters of 256 and 128 workers (bit sets
    # of 4 and 2 words in the batched environments), the configs' BERT-like / ResNet-50-like / GPT-2-like job graphs at full
    # size with small partition degrees, exponential arrivals
    'bert256_shard': dict(graphs=[synth.transformer_like_graph(n_layers=12, name='bert_base_like', seed=2)], shape=(8, 8, 4), n_jobs=4,
                          max_partitions=8, interarrival=700.0, frac=(0.3, 1.0), actor='sipml', seed=31),
    'mix128_exp': dict(graphs=[synth.resnet_like_graph(), synth.transformer_like_graph(n_layers=12, name='gpt2_small_like', seed=5, gpt=True)],
                       shape=(8, 4, 4), n_jobs=3, max_partitions=4, interarrival=('exponential', 500.0), frac=(0.2, 1.0), actor='random',
                       seed=32),
    # BASELINE config 2's cluster and job: the ResNet-50-like graph at full size on the 32-worker 4x4x2 RAMP, random degrees up to 8
    'resnet32_cfg2': dict(graphs=[synth.resnet_like_graph()], shape=(4, 4, 2), n_jobs=4, max_partitions=8, interarrival=900.0,
                          frac=(0.2, 1.0), actor='random', seed=41),
    'residual32_deg16': dict(graphs=[synth.residual_small_graph()], shape=(4, 4, 2), n_jobs=3, max_partitions=16,
                             interarrival=1000.0, frac=(0.1, 1.0), actor='sipml', seed=1),
}


def run_case(name, spec):
    seed = spec['seed']
    np.random.seed(seed)
    random.seed(seed)
    d = tempfile.mkdtemp(prefix='golden_graphs_')
    for g in spec['graphs']:
        g.write(d)
    rec = Recorder()
    rec.install()
    t0 = time.perf_counter()
    try:
        env = make_env(d, spec['shape'], spec['n_jobs'], spec['max_partitions'], spec['interarrival'],
                       Uniform(spec['frac'][0], spec['frac'][1], decimals=2),
                       max_sim_time=spec.get('max_sim_time', 1e6))
        # the env constructor already called reset() once (RJPE:...: "self.reset()"); start a clean recording
        rec.steps.clear(); rec.lookaheads.clear(); rec.arrivals.clear(); rec.templates.clear(); rec.fp_to_tid.clear()
        np.random.seed(seed)
        random.seed(seed)
        obs = env.reset()
        actor = {'random': Random(), 'sipml': SiPML(spec['max_partitions']), 'acceptable_jct': AcceptableJCT()}[spec['actor']]
        done, n_env_steps = False, 0
        while not done:
            job_to_place = list(env.cluster.job_queue.jobs.values())[0]
            a = actor.compute_action(obs, job_to_place=job_to_place)
            obs, _, done, _ = env.step(int(a))
            n_env_steps += 1
        wall = time.perf_counter() - t0
        cluster = env.cluster
        es = cluster.episode_stats
    finally:
        rec.uninstall()

    out = {}
    out['meta_n_cluster_workers'] = np.array(cluster.topology.graph.graph['num_workers'])
    out['meta_max_sim_time'] = np.array(float(cluster.max_simulation_run_time))
    out['meta_n_env_steps'] = np.array(n_env_steps)
    out['meta_reference_wall_s'] = np.array(wall)
    out['meta_n_models'] = np.array(len(rec.models))
    out['n_templates'] = np.array(len(rec.templates))
    for t, lj in enumerate(rec.templates):
        out.update(lj.to_npz_dict(prefix=f't{t}_'))
    out['arrivals'] = np.array(rec.arrivals, dtype=np.float64).reshape(-1, 3)
    out['step_tid'] = np.array([s['tid'] for s in rec.steps], dtype=np.int32)
    out['step_lookahead'] = np.array([s['lookahead'] for s in rec.steps], dtype=np.int32)
    out['step_mount'] = np.array([[s['mount'].max_acceptable_jct, s['mount'].part_op_mem, s['mount'].part_dep_size,
                                   s['mount'].flow_size, s['mount'].n_mounted_workers, s['mount'].n_mounted_channels]
                                  if s['mount'] is not None else [0] * 6 for s in rec.steps], dtype=np.float64)
    out['step_stats'] = np.array([[s['stats'][k] for k in STEP_STATS] for s in rec.steps], dtype=np.float64)
    out['step_time'] = np.array([s['time'] for s in rec.steps], dtype=np.float64)
    out['n_lookaheads'] = np.array(len(rec.lookaheads))
    for i, la in enumerate(rec.lookaheads):
        out[f'la{i}_res'] = np.array([la['jct'], la['comm'], la['comp']], dtype=np.float64)
        out[f'la{i}_tid'] = np.array(la['tid'])
        out[f'la{i}_trace_n'] = la['trace_n']
        out[f'la{i}_trace_tick'] = la['trace_tick']
    # episode stats
    for k in ('num_jobs_arrived', 'num_jobs_completed', 'num_jobs_blocked'):
        out[f'es_{k}'] = np.array(int(es[k]))
    for k in ('episode_start_time', 'episode_end_time', 'episode_time', 'mean_load_rate', 'blocking_rate', 'acceptance_rate',
              'compute_info_processed', 'dep_info_processed', 'flow_info_processed', 'cluster_info_processed',
              'demand_compute_info_processed', 'demand_dep_info_processed', 'demand_total_info_processed',
              'mean_compute_throughput', 'mean_dep_throughput', 'mean_flow_throughput', 'mean_cluster_throughput',
              'mean_demand_compute_throughput', 'mean_demand_dep_throughput', 'mean_demand_total_throughput',
              'mean_compute_overhead_frac', 'mean_communication_overhead_frac', 'mean_num_jobs_running',
              'mean_num_mounted_workers'):
        out[f'es_{k}'] = np.array(float(es[k]))
    for k in ('job_completion_time', 'job_completion_time_speedup', 'job_communication_overhead_time',
              'job_computation_overhead_time', 'jobs_completed_mean_mounted_worker_utilisation_frac',
              'jobs_completed_num_mounted_workers', 'jobs_completed_num_mounted_channels',
              'jobs_completed_max_acceptable_job_completion_time', 'jobs_blocked_max_acceptable_job_completion_time'):
        out[f'es_{k}'] = np.array([float(x) for x in es[k]], dtype=np.float64)
    out['es_completed_job_idxs'] = np.array(list(cluster.jobs_completed.keys()), dtype=np.int32)
    out['es_blocked_job_idxs'] = np.array(list(cluster.jobs_blocked.keys()), dtype=np.int32)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, f'{name}.npz')
    np.savez_compressed(path, **out)
    sizes = [(lj.n_ops, lj.n_deps) for lj in rec.templates]
    print(f'{name}: {n_env_steps} env steps, {len(rec.steps)} cluster steps, {len(rec.lookaheads)} lookaheads, '
          f'{len(rec.templates)} templates {sizes}, arrived/completed/blocked = {es["num_jobs_arrived"]}/'
          f'{es["num_jobs_completed"]}/{es["num_jobs_blocked"]}, reference wall {wall:.1f}s -> {path} '
          f'({os.path.getsize(path) / 1024:.0f} KiB)', flush=True)


if __name__ == '__main__':
    names = sys.argv[1:] or list(CASES)
    if os.environ.get('PYTHONHASHSEED') != '0':
        print('note: run with PYTHONHASHSEED=0 for byte-identical regeneration', file=sys.stderr)
    for n in names:
        run_case(n, CASES[n])
