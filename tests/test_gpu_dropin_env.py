"""-m gpu: the drop-in ``ddls_b200.host.RampClusterEnvironment`` (reference class surface: reset/step/is_done +
the state the agents read) driven with reference-shaped Action objects, against the reference's own recorded run."""
import copy

import numpy as np
import pytest

from conftest import golden_files
from golden_io import Golden

pytestmark = pytest.mark.gpu

SHAPES = {8: (2, 2, 2), 16: (2, 2, 4), 32: (4, 4, 2), 64: (4, 4, 4), 128: (8, 4, 4), 256: (8, 8, 4)}


def _make_env(g):
    from ddls_b200.host import RampClusterEnvironment, A100
    c, r, s = SHAPES[g.n_cluster_workers]
    return RampClusterEnvironment(
        topology_config={'type': 'ramp', 'kwargs': {'num_communication_groups': c, 'num_racks_per_communication_group': r,
                                                    'num_servers_per_rack': s, 'num_channels': 1,
                                                    'total_node_bandwidth': 1.6e12, 'intra_gpu_propagation_latency': 50e-9,
                                                    'worker_io_latency': 100e-9}},
        node_config={'type_1': {'num_nodes': c * r * s, 'workers_config': [{'num_workers': 1, 'worker': A100}]}},
        suppress_warnings=True, max_jobs=64)


@pytest.mark.parametrize('fname', golden_files())
def test_dropin_env_replays_reference_episode(fname):
    from ddls_b200.host import synthetic
    from ddls_b200.engine import SS, STEP_STATS
    g = Golden(fname)
    arr = g.d['arrivals']
    n_jobs = len(arr)
    # which template handles job k: the k-th handled/blocked decision in arrival order is not needed -- the mirror only
    # needs an original job per arrival
    jobs = [synthetic.build_original_job(job_id=100 + k, model='m', orig_op_mem=float(arr[k, 1]), orig_dep_size=float(arr[k, 2]),
                                         frac=0.5, seq_time=1000.0, num_training_steps=50) for k in range(n_jobs)]
    gaps = [float(x) for x in arr[:, 0] if np.isfinite(x)]       # gap drawn when job k arrives
    if np.isfinite(arr[-1, 0]):
        # the reference's generator still held jobs when the episode ended (max_simulation_run_time): keep ours non-empty too
        jobs += [synthetic.build_original_job(900 + k, 'm', 1.0, 1.0, 0.5, 1000.0, 50) for k in range(2)]
        gaps += [float(arr[-1, 0])] * 2
    gen = synthetic.SyntheticJobsGenerator(jobs, gaps)
    env = _make_env(g)
    env.reset(gen, max_simulation_run_time=g.max_sim_time)
    ref = g.d['step_stats']
    for s in range(g.n_steps):
        tmpl = g.step_job(s)
        if tmpl is not None:
            queued = list(env.job_queue.jobs.values())[0]
            tmpl.model_id = 0
            # the fixtures do not store global worker ids: give the job any free workers (the reference's own run had
            # as many free at this point, the dynamics being identical)
            w2n = env.topology.graph.graph['worker_to_node']
            free = [w for w in sorted(w2n) if len(env.topology.graph.nodes[w2n[w]]['workers'][w].mounted_job_idx_to_ops) == 0]
            assert len(free) >= tmpl.n_workers
            action, _ = synthetic.build_action(tmpl, queued, env, worker_ids=free[:tmpl.n_workers])
            # the fixture's memo key is (model, degree); reproduce the model identity through the job's model name
            action.actions['op_partition'].partitioned_jobs[queued.job_id].details['model'] = f'model{g.templates[int(g.d["step_tid"][s])].model_id}'
        else:
            action = synthetic.SyntheticAction()
        _, _, _, done, _ = env.step(action)
        for k in STEP_STATS:
            if k in ('util_mounted_sum', 'util_cluster_sum', 'num_ticks', 'done', 'lookahead_ran'):
                continue
            assert float(env.step_stats[k]) == pytest.approx(ref[s, SS[k]], rel=1e-6, abs=0), (fname, s, k)
        assert done == bool(ref[s, SS['done']])
        assert env.stopwatch.time() == pytest.approx(float(g.d['step_time'][s]), rel=1e-6)
        # state the agents read stays consistent: one job per worker, mounted sets match the running jobs
        for job in env.jobs_running.values():
            for w in job.details['mounted_workers']:
                node = env.topology.graph.graph['worker_to_node'][w]
                assert list(env.topology.graph.nodes[node]['workers'][w].mounted_job_idx_to_ops.keys()) == [job.details['job_idx']]
    es = env.episode_stats
    assert es['num_jobs_arrived'] == int(g.d['es_num_jobs_arrived'])
    assert es['num_jobs_completed'] == int(g.d['es_num_jobs_completed'])
    assert es['num_jobs_blocked'] == int(g.d['es_num_jobs_blocked'])
    assert list(env.jobs_completed.keys()) == list(g.d['es_completed_job_idxs'])
    assert sorted(env.jobs_blocked.keys()) == sorted(g.d['es_blocked_job_idxs'])
    np.testing.assert_allclose(es['job_completion_time'], g.d['es_job_completion_time'], rtol=1e-6, atol=0)
    np.testing.assert_allclose(es['job_communication_overhead_time'], g.d['es_job_communication_overhead_time'], rtol=1e-6, atol=0)
    np.testing.assert_allclose(es['jobs_completed_mean_mounted_worker_utilisation_frac'],
                               g.d['es_jobs_completed_mean_mounted_worker_utilisation_frac'], rtol=1e-6, atol=0)
    for k in ('blocking_rate', 'acceptance_rate', 'mean_load_rate', 'episode_time', 'compute_info_processed',
              'mean_compute_throughput', 'mean_cluster_throughput', 'mean_num_jobs_running', 'mean_num_mounted_workers',
              'mean_compute_overhead_frac', 'mean_communication_overhead_frac'):
        assert float(es[k]) == pytest.approx(float(g.d[f'es_{k}']), rel=1e-6, abs=1e-12), (fname, k)
    # all workers / channels released at the end of the episode
    for node in env.topology.graph.nodes:
        for w in env.topology.graph.nodes[node]['workers'].values():
            assert len(w.mounted_job_idx_to_ops) == 0 and w.memory_occupied == pytest.approx(0, abs=1e-3)


def test_dropin_env_enforces_ramp_rule_one_job_per_worker():
    """RCE:1326-1328: mounting a second job on an occupied worker raises like the reference."""
    from ddls_b200.host import synthetic
    g = Golden('chain8')
    arr = g.d['arrivals']
    jobs = [synthetic.build_original_job(100 + k, 'm', float(arr[k, 1]), float(arr[k, 2]), 1.0, 1e9, 50) for k in range(3)]
    env = _make_env(g)
    env.reset(synthetic.SyntheticJobsGenerator(jobs, [1.0, 1.0]), max_simulation_run_time=1e9)
    t = copy.copy(g.templates[0])
    t.mount = copy.copy(t.mount)
    t.mount.max_acceptable_jct = float('inf')
    q = list(env.job_queue.jobs.values())[0]
    w = sorted(env.topology.graph.graph['worker_to_node'])[:t.n_workers]
    a, _ = synthetic.build_action(t, q, env, worker_ids=w)
    env.step(a)                       # job 0 now runs on its workers; job 1 arrives 1 time unit later
    q = list(env.job_queue.jobs.values())[0]
    a, _ = synthetic.build_action(t, q, env, worker_ids=w)
    with pytest.raises(Exception, match='one_job_per_worker'):
        env.step(a)
