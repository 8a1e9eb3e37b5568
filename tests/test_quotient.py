"""Symmetry quotient of a lowered job (include/ramp_b200.h: ramp_quotient_template, ddls_b200/csrc/ramp_quotient.cpp),
which ramp_register_template applies before a job goes to the device.  Host-only: runs without a GPU.

Exactness is checked end to end: the tick loop the kernels implement (tests/kernel_model.py), run on the QUOTIENT job
with class weights and scaled parent counters (tests/quotient_model.py), must reproduce the reference's recorded
lookaheads (tests/golden) and the oracle bit for bit -- (jct, comm, comp) and every tick of the trace."""
import numpy as np
import pytest

from conftest import golden_files
from golden_io import Golden
import quotient_model as qm
from ddls_b200 import synth
from ddls_b200.quotient import quotient
from ddls_b200.template_builder import RampShape, build_template, random_dag_template

FIELDS = ('op_cost', 'op_key', 'op_worker', 'op_weight', 'op_threshold', 'row_ptr', 'dep_dst', 'dep_run_time', 'dep_key',
          'dep_channel', 'dep_is_flow', 'dep_inc', 'op_class', 'dep_entry', 'dep_group_mask')
MAX_WORK = 40_000_000


def _assert_same_quotient(a, b):
    assert (a.n_ops, a.n_deps, a.n_workers, a.n_channels, a.merged, a.masks_valid) == (b.n_ops, b.n_deps, b.n_workers, b.n_channels, b.merged, b.masks_valid)
    for f in FIELDS:
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f)


def _assert_lookahead(out, ref_trace_n, ref_trace_tick, jct, comm, comp):
    assert out['finished']
    np.testing.assert_array_equal(out['trace_n_active'], ref_trace_n)
    np.testing.assert_array_equal(out['trace_tick'], ref_trace_tick)            # bit-exact f64
    assert out['jct'] == jct and out['comm'] == comm and out['comp'] == comp


def _assert_matches_oracle(out, ref, oracle_lib):
    assert out['n_ticks'] == ref['n_ticks']
    np.testing.assert_array_equal(out['trace_n_active'], ref['trace_n_active'])
    np.testing.assert_array_equal(out['trace_tick'], ref['trace_tick'])
    if ref['status'] == oracle_lib.ORC_ERR_INFINITE_TICK:                            # deadlocked job graph (RCE:462)
        assert not out['finished'] and np.isinf(out['trace_tick'][-1])
    else:
        assert out['jct'] == ref['jct'] and out['comm'] == ref['comm'] and out['comp'] == ref['comp']


def _check_structure(job, q):
    """Invariants of any quotient: sizes add up, every original op / dep went somewhere consistent."""
    assert int(q.op_weight.sum()) == job.n_ops and int(q.dep_inc.sum()) == job.n_deps
    assert np.array_equal(np.bincount(q.op_class, minlength=q.n_ops), q.op_weight)
    if job.n_deps:
        assert np.array_equal(np.bincount(q.dep_entry, minlength=q.n_deps), q.dep_inc)
        src = np.repeat(np.arange(job.n_ops), np.diff(job.row_ptr))
        qsrc = np.repeat(np.arange(q.n_ops), np.diff(q.row_ptr))
        assert np.array_equal(q.op_class[src], qsrc[q.dep_entry])
        assert np.array_equal(q.op_class[job.dep_dst], q.dep_dst[q.dep_entry])
        assert np.array_equal(job.dep_run_time + 0.0, q.dep_run_time[q.dep_entry])
    assert np.array_equal(job.op_cost + 0.0, q.op_cost[q.op_class])
    assert np.array_equal(job.op_n_parents.astype(np.int64) * q.op_weight[q.op_class], q.op_threshold[q.op_class])


@pytest.mark.parametrize('fname', golden_files())
def test_quotient_reproduces_reference_goldens(fname):
    g = Golden(fname)
    for i in range(g.n_lookaheads):
        la = g.lookahead(i)
        job = g.templates[la['tid']]
        q = quotient(job)
        _check_structure(job, q)
        assert q.n_ops <= job.n_ops and q.n_deps <= job.n_deps
        if job.n_deps <= 8000:
            _assert_same_quotient(q, qm.quotient(job))                          # C++ == Python twin, array for array
        if len(la['trace_tick']) * max(q.n_deps, 1) <= MAX_WORK:
            _assert_lookahead(qm.run_lookahead_quotient(q), la['trace_n'], la['trace_tick'], la['jct'], la['comm'], la['comp'])


@pytest.mark.parametrize('mode', ['one_to_one', 'reference'])
@pytest.mark.parametrize('degree', [1, 2, 4, 8, 16])
def test_quotient_of_partitioned_jobs_collapses_the_sub_ops(degree, mode, oracle_lib):
    """Sub-op k of every op on server k: all degrees collapse to one op per original op; the lookahead on the quotient equals
    the oracle's on the full job."""
    g = synth.resnet_like_graph(n_blocks=4, name='res4')
    job = build_template(g, degree, RampShape(4, 4, 4), run_times=mode)
    q = quotient(job)
    _check_structure(job, q)
    assert q.n_ops == 2 * g.n and q.n_workers == 1
    if degree > 1:
        assert np.all(q.op_weight == degree)
        assert q.n_deps < job.n_deps * 2 / degree
    ref = oracle_lib.run_lookahead(job)
    _assert_lookahead(qm.run_lookahead_quotient(q), ref['trace_n_active'], ref['trace_tick'], ref['jct'], ref['comm'], ref['comp'])


def test_quotient_of_the_bench_job_at_full_size(oracle_lib):
    """BASELINE.json config 3's job at degree 16 (N=5,280, E=132,016): 330 op classes, <= 2 entries per original edge."""
    job = build_template(synth.resnet_like_graph(), 16, RampShape(4, 4, 4))
    q = quotient(job)
    _check_structure(job, q)
    assert (q.n_ops, q.n_workers, q.n_channels) == (330, 1, 1) and q.n_deps == 887
    ref = oracle_lib.run_lookahead(job)
    _assert_lookahead(qm.run_lookahead_quotient(q), ref['trace_n_active'], ref['trace_tick'], ref['jct'], ref['comm'], ref['comp'])


@pytest.mark.parametrize('seed', range(12))
def test_quotient_of_adversarial_random_jobs(seed, oracle_lib):
    """Random DAGs (priority ties, zero-cost ops, zero-time flows, channel-less flows, mutual edges): little or no symmetry;
    whatever the pass merges must still give the oracle's lookahead."""
    job = random_dag_template(np.random.default_rng(seed), n_ops=40 + 7 * seed, n_workers=2 + seed % 5)
    q = quotient(job)
    _check_structure(job, q)
    _assert_same_quotient(q, qm.quotient(job))
    ref = oracle_lib.run_lookahead(job)
    _assert_matches_oracle(qm.run_lookahead_quotient(q), ref, oracle_lib)


@pytest.mark.parametrize('seed', range(6))
def test_quotient_of_replicated_random_jobs(seed, oracle_lib):
    """k disjoint copies of one random job on k disjoint worker sets, costs of one copy perturbed in half the seeds: the
    identical copies must merge (weight k), a perturbed copy must stay apart, the lookahead must equal the oracle's."""
    rng = np.random.default_rng(100 + seed)
    base = random_dag_template(rng, n_ops=30, n_workers=3, p_tie=0.0)
    k = 2 + seed % 3
    from ddls_b200.lowered import LoweredJob
    N, E, W, C = base.n_ops, base.n_deps, base.n_workers, base.n_channels
    cost = np.tile(base.op_cost, k)
    if seed % 2:
        cost[N * (k - 1)] += 0.125                                               # break the symmetry of the last copy
    job = LoweredJob(
        n_ops=N * k, n_deps=E * k, n_workers=W * k, n_channels=C * k, num_training_steps=base.num_training_steps, model_id=0,
        degree=2, op_cost=cost, op_prio=np.tile(base.op_prio, k),
        op_worker=np.concatenate([base.op_worker.astype(np.int64) + W * c for c in range(k)]),
        op_n_parents=np.tile(base.op_n_parents, k),
        row_ptr=np.concatenate([[0]] + [base.row_ptr[1:].astype(np.int64) + E * c for c in range(k)]),
        dep_dst=np.concatenate([base.dep_dst.astype(np.int64) + N * c for c in range(k)]),
        dep_run_time=np.tile(base.dep_run_time, k), dep_prio=np.tile(base.dep_prio, k),
        dep_channel=np.concatenate([np.where(base.dep_channel == 0xFFFF, 0xFFFF, base.dep_channel.astype(np.int64) + C * c)
                                    for c in range(k)]),
        dep_is_flow=np.tile(base.dep_is_flow, k), mount=base.mount).canonicalise()
    q = quotient(job)
    _check_structure(job, q)
    _assert_same_quotient(q, qm.quotient(job))
    if seed % 2 == 0:
        assert q.n_ops == N and np.all(q.op_weight == k)
    else:
        assert N < q.n_ops <= 2 * N
    ref = oracle_lib.run_lookahead(job)
    _assert_matches_oracle(qm.run_lookahead_quotient(q), ref, oracle_lib)


@pytest.mark.parametrize('degree', [4, 8, 16])
def test_quotient_merges_entries_over_channel_groups(degree, oracle_lib):
    """Reference run times (collectives whose time depends on which racks / communication groups a server pair spans): several
    channel groups, but one global priority order -> ONE entry per dep class carrying its group set; the lookahead on it equals
    the oracle's on the full job at BASELINE.json's size."""
    from ddls_b200 import workload
    job = workload.reference_template(synth.resnet_like_graph(), degree, RampShape(4, 4, 4))
    q = quotient(job)
    _check_structure(job, q)
    assert q.merged == 1 and q.masks_valid == 1 and q.n_channels >= 2 and q.n_ops == 330
    assert q.n_deps < 1500
    assert int(np.bitwise_or.reduce(q.dep_group_mask)) == (1 << q.n_channels) - 1
    ref = oracle_lib.run_lookahead(job)
    _assert_lookahead(qm.run_lookahead_quotient(q), ref['trace_n_active'], ref['trace_tick'], ref['jct'], ref['comm'], ref['comp'])
