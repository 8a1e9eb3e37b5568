"""-m gpu: ``ddls_b200.batched_cluster.BatchedRampClusterEnvironment`` -- B RampClusterEnvironment episodes in lock step on one
engine, each stepped with a reference-shaped ``Action`` object (the five action parts, duck-typed like the reference's classes) --
replays the reference's recorded golden episodes SIDE BY SIDE: every cluster step of every episode must reproduce the step
statistics the reference recorded, and the lowering cache must make a repeated (model, degree, placement) free."""
import numpy as np
import pytest

from golden_io import Golden

pytestmark = pytest.mark.gpu

SHAPES = {8: (2, 2, 2), 16: (2, 2, 4), 32: (4, 4, 2), 64: (4, 4, 4), 128: (8, 4, 4), 256: (8, 8, 4)}
BATCHES = [
    ['chain8', 'chain8_busy', 'residual8_deg4', 'chain8', 'chain8_busy'],
    ['mixed16', 'res16_flood', 'mixed16', 'res16_flood'],
    ['tfm32_acceptable', 'residual32_deg16', 'resnet32_cfg2'],
    ['mixed64_busy', 'resnet64_deg2_full', 'resnet64_deg4_full', 'mixed64_busy'],
    ['mix128_exp', 'mix128_exp'],
]


def _topology(n_workers):
    c, r, s = SHAPES[n_workers]
    return {'type': 'ramp', 'kwargs': {'num_communication_groups': c, 'num_racks_per_communication_group': r, 'num_servers_per_rack': s,
                                       'num_channels': 1, 'total_node_bandwidth': 1.6e12, 'intra_gpu_propagation_latency': 50e-9,
                                       'worker_io_latency': 100e-9}}


@pytest.mark.parametrize('verify', [False, True], ids=['cached', 'verified'])
@pytest.mark.parametrize('names', BATCHES, ids=lambda n: '+'.join(n))
def test_batched_cluster_env_replays_reference_episodes_with_action_objects(names, verify):
    from ddls_b200.batched_cluster import BatchedRampClusterEnvironment
    from ddls_b200.engine import SS, STEP_STATS, JS_COMPLETED, JS_BLOCKED
    from ddls_b200.host import synthetic
    goldens = [Golden(n) for n in names]
    B = len(goldens)
    n_workers = goldens[0].n_cluster_workers
    assert all(g.n_cluster_workers == n_workers and g.max_sim_time == goldens[0].max_sim_time for g in goldens)
    J = max(len(g.d['arrivals']) for g in goldens)
    env = BatchedRampClusterEnvironment(_topology(n_workers), n_episodes=B, max_jobs=J, verify_cache=verify)
    arrivals = np.zeros((B, J, 3))
    arrivals[:, :, 0] = np.inf
    for b, g in enumerate(goldens):
        arrivals[b, :len(g.d['arrivals'])] = g.d['arrivals']
    env.reset(arrivals, max_simulation_run_time=goldens[0].max_sim_time)
    for b, g in enumerate(goldens):
        env.eng.set_job_count(b, len(g.d['arrivals']))
    all_workers = sorted(env.topology.graph.graph['worker_to_node'])
    n_steps = max(g.n_steps for g in goldens)
    for s in range(n_steps):
        queued = env.queued_job()
        actions = []
        for b, g in enumerate(goldens):
            tmpl = g.step_job(s) if s < g.n_steps else None
            if tmpl is None:
                actions.append(synthetic.SyntheticAction() if s % 2 else None)     # both spellings of "place nothing"
                continue
            k = int(queued[b])
            assert k >= 0, (names[b], s)
            arr = g.d['arrivals'][k]
            orig = synthetic.build_original_job(job_id=1000 * b + k, model=f'model{tmpl.model_id}', orig_op_mem=float(arr[1]),
                                                orig_dep_size=float(arr[2]), frac=0.5, seq_time=1000.0, num_training_steps=tmpl.num_training_steps)
            orig.details['job_idx'] = k
            busy = env.workers_in_use(b)
            free = [w for w in all_workers if w not in busy]
            assert len(free) >= tmpl.n_workers, (names[b], s)
            action, _ = synthetic.build_action(tmpl, orig, env, worker_ids=free[:tmpl.n_workers])
            action.actions['op_partition'].partitioned_jobs[orig.job_id].details['model'] = f'model{tmpl.model_id}'
            # the fixtures do not store global worker ids, so the replay places every job on the first free workers: two recorded
            # jobs of one (model, degree) whose original blocks differed (different collective run times) can then share a placement
            action.lowering_key = int(g.d['step_tid'][s])
            actions.append(action)
        stats = env.step(actions)
        done = env.done()
        for b, g in enumerate(goldens):
            if s >= g.n_steps:
                assert done[b], (names[b], s)
                continue
            ref = g.d['step_stats'][s]
            for key in STEP_STATS:
                if key in ('util_mounted_sum', 'util_cluster_sum', 'num_ticks', 'lookahead_ran'):
                    continue
                assert float(stats[b, SS[key]]) == pytest.approx(float(ref[SS[key]]), rel=1e-6, abs=0), (names[b], s, key)
            assert env.time()[b] == pytest.approx(float(g.d['step_time'][s]), rel=1e-6)
    assert env.done().all()
    rec = env.job_records()
    for b, g in enumerate(goldens):
        r = rec[b][:len(g.d['arrivals'])]
        order = np.argsort(r['event_seq'], kind='stable')
        completed = [int(i) for i in order if r['status'][i] == JS_COMPLETED]
        assert completed == list(g.d['es_completed_job_idxs']), names[b]
        assert sorted(int(i) for i in order if r['status'][i] == JS_BLOCKED) == sorted(g.d['es_blocked_job_idxs']), names[b]
        np.testing.assert_allclose(r['time_completed'][completed] - r['time_arrived'][completed], g.d['es_job_completion_time'], rtol=1e-6, atol=0)
    # repeated (model, degree, placement) decisions were not lowered again (verify=True lowers everything on purpose)
    n_actions = env.stats['actions']
    assert n_actions > 0 and env.stats['cache_hits'] + env.stats['templates'] <= n_actions
    if not verify:
        assert env.stats['lowerings'] == n_actions - env.stats['cache_hits']
        if len(set(names)) < len(names):
            assert env.stats['cache_hits'] > 0
    env.close()


def test_batched_cluster_env_rejects_bad_input():
    from ddls_b200.batched_cluster import BatchedRampClusterEnvironment
    env = BatchedRampClusterEnvironment(_topology(8), n_episodes=2, max_jobs=4)
    with pytest.raises(Exception, match='arrivals must be'):
        env.reset(np.zeros((3, 4, 3)))
    env.reset(np.full((2, 4, 3), 1.0))
    with pytest.raises(Exception, match='actions for 2 episodes'):
        env.step([None])
    env.close()
    with pytest.raises(Exception, match='Unrecognised topology'):
        BatchedRampClusterEnvironment({'type': 'torus', 'kwargs': {}})
