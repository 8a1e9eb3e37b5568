"""ddls_b200/placer.py against the decisions the unmodified reference's RampFirstFitOpPlacer took in seeded busy-cluster
episodes (tests/fixtures/placer_cases.json, written by oracle/gen_placer_cases.py): same servers for every sub-op, same failures."""
import json
import os

import pytest

from ddls_b200.placer import first_fit_place

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fixtures', 'placer_cases.json')
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


@pytest.mark.parametrize('i', range(len(CASES)))
def test_native_first_fit_matches_reference(i):
    """The same decisions through the C ABI (ramp_first_fit_place)."""
    from ddls_b200.placer import first_fit_place_native
    c = CASES[i]
    n = len(c['nodes'])
    assert c['nodes'] == [str(k) for k in range(1, n + 1)]
    split_of = dict(zip(c['mp_split_ids'], c['mp_splits']))
    got = first_fit_place_native(n, c['mem'], [(int(u), int(v)) for u, v in c['edges']], [split_of.get(str(k), 1) for k in range(1, n + 1)],
                                 {tuple(k): m for k, m, _ in c['ramp']}, {tuple(k): bool(j) for k, _, j in c['ramp']}, tuple(c['shape']))
    if c['placement'] is None:
        assert got is None
    else:
        assert got is not None and {k: list(v) for k, v in got.items()} == c['placement']
