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
