"""ddls_b200/observation.py against 81 observations recorded from the unmodified reference's encoder in seeded episodes
(tests/fixtures/obs_cases.npz, written by oracle/gen_obs_cases.py): every array of the observation dict identical."""
import os

import numpy as np
import pytest

from ddls_b200.observation import encode_observation

D = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fixtures', 'obs_cases.npz'))
N_CASES = int(D['n_cases'])
KEYS = ('action_set', 'action_mask', 'node_features', 'edge_features', 'graph_features', 'edges_src', 'edges_dst',
        'node_split', 'edge_split')


def test_fixture_has_busy_clusters_and_masked_actions():
    assert N_CASES >= 50
    assert any(D[f'c{i}_scalars'][13] > 0 for i in range(N_CASES))                   # mounted workers > 0
    assert any((D[f'c{i}_obs_action_mask'] == 0).any() for i in range(N_CASES))


@pytest.mark.parametrize('i', range(N_CASES))
def test_observation_matches_reference(i):
    p = f'c{i}_'
    s = D[p + 'scalars']
    got = encode_observation(
        D[p + 'op_compute'], D[p + 'op_memory'], D[p + 'op_depth'], D[p + 'edge_src'], D[p + 'edge_dst'], D[p + 'edge_size'],
        D[p + 'params'], max_compute_cost=s[0], max_compute_op=int(s[1]), max_memory_cost=s[2], max_memory_op=int(s[3]),
        max_dep_size=s[4], max_dep_index=int(s[5]), max_depth=s[6], sequential_completion_time=s[7], max_acceptable_jct=s[8],
        max_acceptable_frac=s[9], total_op_memory=s[10], total_dep_size=s[11], num_training_steps=s[12],
        n_mounted_workers=int(s[13]), n_jobs_running=int(s[14]), n_workers=int(s[15]), shape=(int(s[16]), int(s[17]), int(s[18])),
        max_partitions_per_op=int(s[19]), max_nodes=int(s[20]), machine_epsilon=float(s[21]))
    for k in KEYS:
        want = D[p + 'obs_' + k]
        assert got[k].dtype == want.dtype, (k, got[k].dtype, want.dtype)
        np.testing.assert_array_equal(got[k], want, err_msg=k)


def _catalogue():
    from ddls_b200 import synth
    return [synth.chain_graph(6, 'chain6'), synth.chain_graph(5, 'chain5'), synth.chain_graph(4, 'chain4'),
            synth.resnet_like_graph(n_blocks=2, stem=2, name='res2', seed=7, body_per_block=3),
            synth.resnet_like_graph(n_blocks=1, stem=2, name='res1', seed=11, body_per_block=2),
            synth.transformer_like_graph(n_layers=1, name='tfm1', seed=4), synth.transformer_like_graph(n_layers=1, name='tfm1b', seed=6),
            synth.transformer_like_graph(n_layers=2, name='tfm2', seed=9)]


def test_job_arrays_from_the_profile_alone_equal_the_reference_job_graph():
    """observation.job_arrays derives, from the forward profile only, the node / edge order, depths, first-maximum ops and
    totals of the mirrored job graph the reference builds (utils.py:342-398, job.py:250-325): identical to what was read off
    the reference's Job objects in all 81 recorded cases, and the observation encoded from them is the recorded one."""
    from ddls_b200.observation import job_arrays, static_observation
    graphs = _catalogue()
    arrs = [job_arrays(g) for g in graphs]
    seen = set()
    for i in range(N_CASES):
        p = f'c{i}_'
        s = D[p + 'scalars']
        m = [k for k, a in enumerate(arrs) if len(a['op_compute']) == len(D[p + 'op_compute']) and np.array_equal(a['op_compute'], D[p + 'op_compute'])]
        assert len(m) == 1
        a = arrs[m[0]]
        seen.add(graphs[m[0]].name)
        for k in ('op_memory', 'op_depth', 'edge_src', 'edge_dst', 'edge_size'):
            np.testing.assert_array_equal(a[k], D[p + k], err_msg=k)
        assert [a['max_compute_cost'], a['max_compute_op'], a['max_memory_cost'], a['max_memory_op'], a['max_dep_size'],
                a['max_dep_index'], a['max_depth']] == list(s[:7])
        assert a['sequential_completion_time'] * s[12] == pytest.approx(s[7], rel=1e-12)
        assert a['total_op_memory'] == s[10] and a['total_dep_size'] == s[11]
        kw = {k: v for k, v in a.items() if k != 'sequential_completion_time'}
        got = encode_observation(params=D[p + 'params'], sequential_completion_time=s[7], max_acceptable_jct=s[8], max_acceptable_frac=s[9],
                                 num_training_steps=s[12], n_mounted_workers=int(s[13]), n_jobs_running=int(s[14]), n_workers=int(s[15]),
                                 shape=(int(s[16]), int(s[17]), int(s[18])), max_partitions_per_op=int(s[19]), max_nodes=int(s[20]),
                                 machine_epsilon=float(s[21]), **kw)
        for k in KEYS:
            np.testing.assert_array_equal(got[k], D[p + 'obs_' + k], err_msg=k)
        st = static_observation(graphs[m[0]])
        N, E = len(a['op_compute']), len(a['edge_size'])
        np.testing.assert_array_equal(st['node_features'], D[p + 'obs_node_features'][:N])
        np.testing.assert_array_equal(st['edge_features'], D[p + 'obs_edge_features'][:E])
        np.testing.assert_array_equal(st['graph_static'], D[p + 'obs_graph_features'][9:15])
    assert len(seen) == len(graphs)
