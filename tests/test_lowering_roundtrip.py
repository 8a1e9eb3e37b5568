"""CPU: lower_job() applied to reference-shaped objects (ddls_b200.host.synthetic, same attribute names as the
reference's Job / Action classes) reproduces the lowered arrays they were built from -- exercises the boundary
without the reference installed.  tests/golden/*.npz cover the same function on the reference's own objects."""
import numpy as np
import pytest

from conftest import golden_files
from golden_io import Golden
from ddls_b200.host import synthetic
from ddls_b200.host.topology import Ramp
from ddls_b200.host.devices import A100
from ddls_b200.lowering import lower_job, ModelRegistry
from ddls_b200.lowered import LoweredJob


class _Cluster:
    def __init__(self, shape):
        c, r, s = shape
        self.topology = Ramp(c, r, s)
        g = self.topology.graph.graph
        g['worker_to_node'], g['worker_to_type'] = {}, {}
        for node in self.topology.graph.nodes:
            w = A100(processor_id=f'node_{node}_worker_0')
            g['worker_to_node'][w.processor_id] = node
            g['worker_to_type'][w.processor_id] = 'A100'


@pytest.mark.parametrize('fname', golden_files())
def test_roundtrip_golden_templates(fname):
    g = Golden(fname)
    shape = {8: (2, 2, 2), 16: (2, 2, 4), 32: (4, 4, 2), 64: (4, 4, 4), 128: (8, 4, 4), 256: (8, 8, 4)}[g.n_cluster_workers]
    cluster = _Cluster(shape)
    for t, lj in enumerate(g.templates):
        orig = synthetic.build_original_job(job_id=t, model=f'm{lj.model_id}', orig_op_mem=1.0, orig_dep_size=2.0,
                                            frac=0.5, seq_time=100.0, num_training_steps=lj.num_training_steps)
        orig.details['job_idx'] = t
        action, _ = synthetic.build_action(lj, orig, cluster)
        back = lower_job(cluster, action, t, ModelRegistry())
        for name, _ in LoweredJob.ARRAYS:
            np.testing.assert_array_equal(getattr(back, name), getattr(lj, name), err_msg=f'{fname} t{t} {name}')
        assert (back.n_ops, back.n_deps, back.n_workers, back.n_channels) == (lj.n_ops, lj.n_deps, lj.n_workers, lj.n_channels)
        assert back.degree == lj.degree and back.num_training_steps == lj.num_training_steps
        assert back.mount.n_mounted_workers == lj.mount.n_mounted_workers
        assert back.mount.n_mounted_channels == lj.mount.n_mounted_channels
        assert back.mount.max_acceptable_jct == lj.mount.max_acceptable_jct
        assert back.mount.flow_size == pytest.approx(lj.mount.flow_size, rel=1e-9)


def test_lowering_rejects_unsupported_inputs():
    g = Golden('chain8')
    lj = g.templates[0]
    cluster = _Cluster((2, 2, 2))
    orig = synthetic.build_original_job(0, 'm', 1.0, 2.0, 0.5, 100.0, lj.num_training_steps)
    orig.details['job_idx'] = 0
    action, pjob = synthetic.build_action(lj, orig, cluster)
    # a dep placed on two channels (multi-hop) is not a RAMP placement
    dep = next(d for d, chans in action.actions['dep_placement'].action[0].items() if None not in chans)
    action.actions['dep_placement'].action[0][dep].add('src_0-0-0_dst_1-1-1_channel_0')
    with pytest.raises(Exception, match='multi-channel'):
        lower_job(cluster, action, 0)
    # an op without a worker
    action2, _ = synthetic.build_action(lj, orig, cluster)
    some_op = next(iter(action2.actions['op_placement'].action[0]))
    del action2.actions['op_placement'].action[0][some_op]
    with pytest.raises(Exception, match='no worker'):
        lower_job(cluster, action2, 0)


def test_synthetic_builder_matches_reference_pipeline_structure():
    """The bench's synthetic Action generator (ddls_b200/template_builder.py) against the template the unmodified reference
    pipeline (OpPartition -> RampFirstFitOpPlacer -> SRPT schedulers -> FirstFitDepPlacer, lowered by lower_job) produced for
    the same graph at degree 16 on an empty 64-worker cluster (tests/golden/resnet64_deg16_full.npz): the partitioned graph,
    op costs, parent counts, job-local placement, op priorities and the flow / non-flow split are IDENTICAL; only the dep
    run times (the reference's collective formulas, actions/utils.py:40-99, are not restated) and the priorities derived
    from them differ."""
    import numpy as np
    from golden_io import Golden
    from ddls_b200 import synth
    from ddls_b200.template_builder import build_template, RampShape
    ref = Golden('resnet64_deg16_full').templates[0]
    mine = build_template(synth.resnet_like_graph(), 16, RampShape(4, 4, 4))
    assert (mine.n_ops, mine.n_deps, mine.n_workers, mine.n_channels) == (ref.n_ops, ref.n_deps, ref.n_workers, ref.n_channels)
    for f in ('row_ptr', 'dep_dst', 'op_cost', 'op_n_parents', 'op_worker', 'op_prio', 'dep_is_flow'):
        np.testing.assert_array_equal(np.asarray(getattr(mine, f)), np.asarray(getattr(ref, f)), err_msg=f)
    nonflow = np.asarray(ref.dep_is_flow) == 0
    assert (np.asarray(mine.dep_run_time)[nonflow] == 0).all() and (np.asarray(ref.dep_run_time)[nonflow] == 0).all()


@pytest.mark.parametrize('degree', [2, 4, 8, 16])
def test_reference_run_time_formulas_reproduce_reference_pipeline(degree, oracle_lib):
    """run_times='reference' restates update_dep_run_times (collective all-reduce / sync / one-to-one, actions/utils.py:13-393)
    and the SRPT dep schedule: against the job the unmodified reference lowered for the same graph (empty 64-worker cluster)
    EVERY array is identical except the dep priorities inside groups of equal run time -- the reference breaks those ties in
    the iteration order of a Python set of strings (DepPlacement.jobdeps, dep_placement.py:16), i.e. by string hash -- and the
    lookahead on the generated job equals the reference's recorded (jct, comm, comp) and tick trace bit for bit."""
    import numpy as np
    from golden_io import Golden
    from ddls_b200 import synth
    from ddls_b200.template_builder import build_template, RampShape
    g = Golden(f'resnet64_deg{degree}_full')
    ref, la = g.templates[0], g.lookahead(0)
    mine = build_template(synth.resnet_like_graph(), degree, RampShape(4, 4, 4), run_times='reference')
    for f in ('row_ptr', 'dep_dst', 'op_cost', 'op_n_parents', 'op_worker', 'op_prio', 'dep_is_flow', 'dep_channel', 'dep_run_time'):
        np.testing.assert_array_equal(np.asarray(getattr(mine, f)), np.asarray(getattr(ref, f)), err_msg=f)
    # priorities: both are SRPT orders (non-increasing run time along increasing priority); only the order inside groups
    # of equal run time (and where the non-flow members of a collective fall inside their group) may differ
    rt, pm, pr = np.asarray(ref.dep_run_time), np.asarray(mine.dep_prio), np.asarray(ref.dep_prio)
    flow = np.nonzero(np.asarray(ref.dep_is_flow) == 1)[0]
    for prio in (pm, pr):
        along = rt[flow][np.argsort(prio[flow], kind='stable')]
        assert (np.diff(along) <= 0).all()
        assert len(np.unique(prio[flow])) == len(flow)
    out = oracle_lib.run_lookahead(mine)
    np.testing.assert_array_equal(out['trace_tick'], la['trace_tick'])
    np.testing.assert_array_equal(out['trace_n_active'], la['trace_n'])
    assert (out['jct'], out['comm'], out['comp']) == (la['jct'], la['comm'], la['comp'])
