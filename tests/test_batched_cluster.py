"""CPU: the host logic of ``ddls_b200.batched_cluster.BatchedRampClusterEnvironment`` (Action -> signature -> cached lowering ->
action rows, occupancy bookkeeping) with the engine calls answered by the oracle (tests/fake_engine.py, one episode); the same
class over B episodes on the CUDA engine is tests/test_gpu_batched_cluster.py."""
import numpy as np
import pytest

from golden_io import Golden

SHAPES = {8: (2, 2, 2), 16: (2, 2, 4), 32: (4, 4, 2), 64: (4, 4, 4)}


@pytest.mark.parametrize('name', ['chain8_busy', 'mixed16', 'res16_flood', 'tfm32_acceptable'])
def test_one_episode_through_the_batched_class_equals_the_reference(name, monkeypatch):
    from fake_engine import FakeEngine
    from ddls_b200 import batched_cluster
    from ddls_b200.engine import SS, STEP_STATS
    from ddls_b200.host import synthetic
    monkeypatch.setattr(batched_cluster._engine, 'RampEngine', FakeEngine)
    g = Golden(name)
    c, r, s_ = SHAPES[g.n_cluster_workers]
    env = batched_cluster.BatchedRampClusterEnvironment(
        {'type': 'ramp', 'kwargs': {'num_communication_groups': c, 'num_racks_per_communication_group': r, 'num_servers_per_rack': s_,
                                    'num_channels': 1, 'total_node_bandwidth': 1.6e12, 'intra_gpu_propagation_latency': 50e-9,
                                    'worker_io_latency': 100e-9}}, n_episodes=1, max_jobs=len(g.d['arrivals']), verify_cache=True)
    env.reset(g.d['arrivals'][None, :, :], max_simulation_run_time=g.max_sim_time)
    workers = sorted(env.topology.graph.graph['worker_to_node'])
    for s in range(g.n_steps):
        tmpl = g.step_job(s)
        action = None
        if tmpl is not None:
            k = int(env.queued_job()[0])
            arr = g.d['arrivals'][k]
            orig = synthetic.build_original_job(job_id=k, model=f'model{tmpl.model_id}', orig_op_mem=float(arr[1]), orig_dep_size=float(arr[2]),
                                                frac=0.5, seq_time=1000.0, num_training_steps=tmpl.num_training_steps)
            orig.details['job_idx'] = k
            free = [w for w in workers if w not in env.workers_in_use(0)]
            assert len(free) >= tmpl.n_workers
            action, _ = synthetic.build_action(tmpl, orig, env, worker_ids=free[:tmpl.n_workers])
            action.actions['op_partition'].partitioned_jobs[orig.job_id].details['model'] = f'model{tmpl.model_id}'
            action.lowering_key = int(g.d['step_tid'][s])
        stats = env.step([action])
        for key in STEP_STATS:
            if key in ('util_mounted_sum', 'util_cluster_sum', 'num_ticks', 'lookahead_ran'):
                continue
            assert float(stats[0, SS[key]]) == pytest.approx(float(g.d['step_stats'][s][SS[key]]), rel=1e-6, abs=0), (s, key)
    assert env.done()[0]
    assert env.stats['actions'] == env.stats['lowerings'] and env.stats['templates'] <= env.stats['actions']
    assert env.workers_in_use(0) == set()
