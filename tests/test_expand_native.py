"""Native template expansion (include/ramp_b200.h: ramp_expand_template, ddls_b200/csrc/ramp_expand.cpp)
against its Python twin (ddls_b200/template_builder.py, itself pinned against the reference pipeline's lowered jobs in
tests/test_lowering_roundtrip.py) and, transitively and directly, against the reference's own fixtures.  Host-only code:
runs without a GPU."""
import numpy as np
import pytest

from golden_io import Golden
from ddls_b200 import synth
from ddls_b200.expand import expand_template
from ddls_b200.template_builder import build_template, RampShape

ARRAYS = ('row_ptr', 'dep_dst', 'op_cost', 'op_n_parents', 'op_worker', 'op_prio', 'dep_is_flow', 'dep_channel',
          'dep_run_time', 'dep_prio')
GRAPHS = {
    'resnet': lambda: synth.resnet_like_graph(n_blocks=4, name='res4'),
    'residual': synth.residual_small_graph,
    'chain': lambda: synth.chain_graph(6, 'chain6'),
    'transformer': lambda: synth.transformer_like_graph(n_layers=2, name='tfm2', seed=9),
}


@pytest.mark.parametrize('mode', ['one_to_one', 'reference'])
@pytest.mark.parametrize('degree', [1, 2, 4, 8, 16])
@pytest.mark.parametrize('gname', sorted(GRAPHS))
def test_native_expansion_equals_python_builder(gname, degree, mode):
    g = GRAPHS[gname]()
    shape = RampShape(4, 4, 4)
    a = build_template(g, degree, shape, run_times=mode)
    b = expand_template(g, degree, shape, run_times=mode)
    assert (a.n_ops, a.n_deps, a.n_workers, a.n_channels) == (b.n_ops, b.n_deps, b.n_workers, b.n_channels)
    for f in ARRAYS:
        np.testing.assert_array_equal(np.asarray(getattr(a, f)), np.asarray(getattr(b, f)), err_msg=f)     # bit-exact f64
    assert a.mount == b.mount


@pytest.mark.parametrize('block_start', [16, 48])
def test_native_expansion_other_blocks_and_topologies(block_start):
    g = synth.residual_small_graph()
    for shape in (RampShape(4, 4, 4), RampShape(8, 8, 4), RampShape(4, 4, 8)):
        a = build_template(g, 8, shape, block_start=block_start)
        b = expand_template(g, 8, shape, block_start=block_start)
        for f in ARRAYS:
            np.testing.assert_array_equal(np.asarray(getattr(a, f)), np.asarray(getattr(b, f)), err_msg=f)
        assert a.mount == b.mount


@pytest.mark.parametrize('degree', [2, 4, 8, 16])
def test_native_expansion_reproduces_reference_lowered_job(degree):
    """Directly against the job the unmodified reference lowered (empty 64-worker cluster, bench graph): every array but the
    hash-ordered priority ties inside groups of equal run time (see tests/test_lowering_roundtrip.py)."""
    ref = Golden(f'resnet64_deg{degree}_full').templates[0]
    mine = expand_template(synth.resnet_like_graph(), degree, RampShape(4, 4, 4), run_times='reference')
    for f in ARRAYS[:-1]:
        np.testing.assert_array_equal(np.asarray(getattr(mine, f)), np.asarray(getattr(ref, f)), err_msg=f)


def test_native_expansion_rejects_bad_arguments():
    with pytest.raises(Exception):
        expand_template(synth.chain_graph(4, 'c4'), 3, RampShape(2, 2, 2))           # odd degree (op_partition.py:26-27)
    with pytest.raises(Exception):
        expand_template(synth.chain_graph(4, 'c4'), 16, RampShape(2, 2, 2))          # block larger than the cluster
    with pytest.raises(Exception, match='more sub-ops than the block has servers'):
        expand_template(synth.chain_graph(6, 'chain6'), 16, RampShape(4, 4, 2), quantum=2.0, coords=[(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1)])


def test_block_may_be_smaller_than_the_action_when_no_op_splits_that_far():
    """RJPE:332-343 gives every op clamp(even(ceil(cost / quantum)), 1, action) sub-ops: with a coarse quantum the most-split op may
    take fewer sub-ops than the action allows, the placer then hands out that many servers, and the lowered job is the one of the
    smaller action."""
    g = synth.chain_graph(6, 'chain6')
    coords = [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1)]
    for mode in ('one_to_one', 'reference'):
        a = expand_template(g, 16, RampShape(4, 4, 2), quantum=2.0, coords=coords, run_times=mode)      # splits 2, 2, 4, 4, 4, 6
        b = expand_template(g, 6, RampShape(4, 4, 2), quantum=2.0, coords=coords, run_times=mode)
        assert a.n_workers == 6 and a.n_ops == b.n_ops == 44
        for f in ARRAYS:
            np.testing.assert_array_equal(np.asarray(getattr(a, f)), np.asarray(getattr(b, f)), err_msg=f)
        assert a.mount == b.mount


# graphs / topologies of the seeded reference episodes behind tests/fixtures/placer_cases.json (oracle/gen_golden.py CASES,
# restated here because that script imports the reference) and how many placements each episode recorded
EPISODES = [
    ('chain8_busy', lambda: [synth.chain_graph(6, 'chain6')], (2, 2, 2), 14),
    ('mixed16', lambda: [synth.chain_graph(5, 'chain5'), synth.resnet_like_graph(n_blocks=2, stem=2, name='res2', seed=7, body_per_block=3),
                         synth.transformer_like_graph(n_layers=1, name='tfm1', seed=4)], (2, 2, 4), 15),
    ('mixed64_busy', lambda: [synth.chain_graph(4, 'chain4'), synth.resnet_like_graph(n_blocks=1, stem=2, name='res1', seed=11, body_per_block=2),
                              synth.transformer_like_graph(n_layers=1, name='tfm1b', seed=6)], (4, 4, 4), 24),
    ('res16_flood', lambda: [synth.resnet_like_graph(n_blocks=1, stem=1, name='res1s', seed=3, body_per_block=2)], (2, 2, 4), 6),
    ('tfm32_acceptable', lambda: [synth.transformer_like_graph(n_layers=2, name='tfm2', seed=9)], (4, 4, 2), 6),
    ('residual32_deg16', lambda: [synth.residual_small_graph()], (4, 4, 2), 3),
]


def test_placer_plus_native_expansion_reproduce_reference_jobs_on_busy_clusters():
    """The whole host chain of SURVEY 8f-1/-4 -- ddls_b200/placer.py picks the servers, ramp_expand_template builds the job --
    against every job the unmodified reference's agents lowered in six seeded busy-cluster episodes (67 jobs: chain, ResNet-like
    and transformer-like graphs, degrees 2-16, four topologies): every array equal but the hash-ordered priority ties."""
    import json
    import os
    from ddls_b200.placer import first_fit_place
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fixtures', 'placer_cases.json')))
    at, checked = 0, 0
    for name, graphs, shape, n_cases in EPISODES:
        g = Golden(name)
        mine_cases, at = cases[at:at + n_cases], at + n_cases
        tids = [int(t) for t in g.d['step_tid'] if t >= 0]
        placed = [c for c in mine_cases if c['placement'] is not None]
        assert len(placed) == len(tids)
        for c, tid in zip(placed, tids):
            ref = g.templates[tid]
            fwd = [x for x in graphs() if x.n == len(c['nodes']) and np.allclose(np.add(x.act, x.par), c['mem'])]
            assert len(fwd) == 1
            ramp = {tuple(k): {'mem': m, 'job_idxs': set(j)} for k, m, j in c['ramp']}
            where = first_fit_place(c['nodes'], c['mem'], c['in_edges'], c['out_edges'], dict(zip(c['mp_split_ids'], c['mp_splits'])),
                                    ramp, tuple(c['shape']), [tuple(s) for s in c['servers']], c['job_idx'])
            degree = max(c['mp_splits']) if c['mp_splits'] else 1
            mine = expand_template(fwd[0], degree, RampShape(*shape), run_times='reference', coords=sorted(set(where.values())))
            for f in ARRAYS[:-1]:
                np.testing.assert_array_equal(np.asarray(getattr(mine, f)), np.asarray(getattr(ref, f)), err_msg=f'{name} tid {tid} {f}')
            checked += 1
    assert checked == 67


def test_native_expansion_depends_on_the_block_only_through_its_geometry():
    """Blocks that are equal after an order-preserving relabelling of each coordinate axis (communication group, rack, server)
    give byte-identical lowered jobs -- what the batched environment's template cache keys on."""
    g = synth.resnet_like_graph(n_blocks=2, stem=2, name='res2', seed=7, body_per_block=3)
    shape = RampShape(4, 4, 4)
    groups = [[[(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0)], [(2, 2, 1), (3, 2, 1), (2, 3, 1), (3, 3, 1)], [(0, 1, 3), (2, 1, 3), (0, 3, 3), (2, 3, 3)]],
              [[(0, 0, 0), (0, 0, 1), (0, 0, 2), (0, 0, 3)], [(1, 2, 0), (1, 2, 1), (1, 2, 2), (1, 2, 3)]],
              [[(0, 0, 0), (1, 1, 0)], [(2, 0, 3), (3, 2, 3)]]]
    for blocks in groups:
        first = None
        for b in blocks:
            t = expand_template(g, len(b), shape, run_times='reference', coords=b)
            if first is None:
                first = t
                continue
            for f in ARRAYS:
                np.testing.assert_array_equal(np.asarray(getattr(t, f)), np.asarray(getattr(first, f)), err_msg=f)
            assert t.mount == first.mount
