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


def test_place_many_equals_one_by_one():
    """ramp_first_fit_place_many (bit-set cluster states, one call) == ramp_first_fit_place per state."""
    import ctypes as C
    import numpy as np
    from ddls_b200 import engine, synth
    from ddls_b200.expand import _FwdGraph
    from ddls_b200.placer import first_fit_place_native
    g = synth.resnet_like_graph(n_blocks=2, stem=2, name='res2', seed=7, body_per_block=3)
    shape = (4, 4, 4)
    servers = [(c, r, s) for c in range(4) for r in range(4) for s in range(4)]
    mem = np.ascontiguousarray([a + p for a, p in zip(g.act, g.par)], dtype=np.float64)
    zero = np.zeros(g.n)
    es = np.ascontiguousarray([u for u, _ in g.edges], dtype=np.int32)
    ed = np.ascontiguousarray([v for _, v in g.edges], dtype=np.int32)
    cg = _FwdGraph(g.n, len(g.edges), zero.ctypes.data, zero.ctypes.data, mem.ctypes.data, zero.ctypes.data, es.ctypes.data, ed.ctypes.data)
    L = engine.load_library()
    L.ramp_first_fit_place_many.restype = C.c_int
    L.ramp_first_fit_place_many.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    for degree in (2, 4, 8, 16):
        splits = np.full(g.n, degree, dtype=np.int32)
        states = np.zeros((40, 1), dtype=np.uint64)
        for k in range(40):
            busy = rng.random(64) < rng.uniform(0.0, 0.9)
            states[k, 0] = np.uint64(sum(1 << i for i in range(64) if busy[i]))
        masks = np.zeros((40, 1), dtype=np.uint64)
        ok = np.zeros(40, dtype=np.uint8)
        rc = L.ramp_first_fit_place_many(C.byref(cg), splits.ctypes.data, (C.c_int32 * 3)(*shape), 80e9, 40, 1, states.ctypes.data,
                                         masks.ctypes.data, ok.ctypes.data)
        assert rc == 0
        for k in range(40):
            busy = {sv: bool((int(states[k, 0]) >> i) & 1) for i, sv in enumerate(servers)}
            where = first_fit_place_native(g.n, mem.tolist(), g.edges, splits.tolist(), {sv: 80e9 for sv in servers}, busy, shape)
            if where is None:
                assert ok[k] == 0
            else:
                want = sum(1 << servers.index(sv) for sv in set(where.values()))
                assert ok[k] == 1 and int(masks[k, 0]) == want
