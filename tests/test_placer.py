"""ddls_b200/placer.py against the decisions the unmodified reference's RampFirstFitOpPlacer took in seeded busy-cluster
episodes (tests/golden/placer_cases.json, written by oracle/gen_placer_cases.py): same servers for every sub-op, same failures."""
import json
import os

import pytest

from ddls_b200.placer import first_fit_place

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'placer_cases.json')
CASES = json.load(open(PATH))


def test_fixture_covers_busy_clusters_and_a_failure():
    assert len(CASES) >= 50
    assert any(c['placement'] is None for c in CASES)
    assert any(any(j for _, _, j in c['ramp']) for c in CASES)            # some servers already hold other jobs


@pytest.mark.parametrize('i', range(len(CASES)))
def test_first_fit_matches_reference(i):
    c = CASES[i]
    ramp = {tuple(k): {'mem': m, 'job_idxs': set(j)} for k, m, j in c['ramp']}
    split_of = dict(zip(c['mp_split_ids'], c['mp_splits']))
    got = first_fit_place(c['nodes'], c['mem'], c['in_edges'], c['out_edges'], split_of, ramp, tuple(c['shape']),
                          [tuple(s) for s in c['servers']], c['job_idx'])
    if c['placement'] is None:
        assert got is None
    else:
        assert got is not None
        assert {k: list(v) for k, v in got.items()} == c['placement']
