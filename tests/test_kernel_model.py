"""The kernel's tick algorithm (tests/kernel_model.py: rank keys, double-buffered channel winners, one-tick non-flow
list, zero-length ticks that skip the flow pass) reproduces the reference's recorded lookaheads bit for bit, and the
oracle on adversarial random jobs.  Runs on CPU: it checks the DESIGN the CUDA kernels implement."""
import numpy as np
import pytest

from conftest import golden_files
from golden_io import Golden
from kernel_model import run_lookahead_model

FILES = golden_files()
MAX_WORK = 40_000_000     # ticks x deps budget per lookahead for the pure-Python model (a few seconds)


@pytest.mark.parametrize('fname', FILES)
def test_model_matches_reference_goldens(fname):
    g = Golden(fname)
    checked = 0
    for i in range(g.n_lookaheads):
        la = g.lookahead(i)
        job = g.templates[la['tid']]
        if len(la['trace_tick']) * max(job.n_deps, 1) > MAX_WORK:
            continue
        out = run_lookahead_model(job)
        assert out['finished']
        np.testing.assert_array_equal(out['trace_n_active'], la['trace_n'])
        np.testing.assert_array_equal(out['trace_tick'], la['trace_tick'])      # bit-exact f64
        assert out['jct'] == la['jct'] and out['comm'] == la['comm'] and out['comp'] == la['comp']
        checked += 1
    if checked == 0:
        pytest.skip('every lookahead of this fixture is too large for the pure-Python model')


@pytest.mark.parametrize('seed', range(6))
def test_model_matches_oracle_on_random_jobs(seed, oracle_lib):
    from ddls_b200.template_builder import random_dag_template
    job = random_dag_template(np.random.default_rng(seed), n_ops=60, n_workers=5)
    ref = oracle_lib.run_lookahead(job)
    out = run_lookahead_model(job)
    assert out['n_ticks'] == ref['n_ticks']
    np.testing.assert_array_equal(out['trace_n_active'], ref['trace_n_active'])
    np.testing.assert_array_equal(out['trace_tick'], ref['trace_tick'])
    assert out['jct'] == ref['jct'] and out['comm'] == ref['comm'] and out['comp'] == ref['comp']


def test_model_matches_oracle_on_wide_fan(oracle_lib):
    """Hundreds of flows sharing a handful of channels, bulk completions, zero-cost ops (the template of the GPU overflow test)."""
    from test_gpu_parity import _fan_template
    job = _fan_template(200, 4, seed=200)
    ref = oracle_lib.run_lookahead(job)
    out = run_lookahead_model(job)
    np.testing.assert_array_equal(out['trace_n_active'], ref['trace_n_active'])
    np.testing.assert_array_equal(out['trace_tick'], ref['trace_tick'])
    assert out['jct'] == ref['jct'] and out['comm'] == ref['comm'] and out['comp'] == ref['comp']
