"""The drop-in class keeps the reference's call signatures.  Needs the read-only reference checkout (through
oracle/ref_shim.py), so it runs in the build container only and is skipped on the GPU box."""
import inspect

import pytest

from oracle import ref_shim


@pytest.mark.skipif(not ref_shim.reference_available(), reason='reference checkout not present (GPU box)')
def test_dropin_class_signatures_match_reference():
    ref_shim.install()
    from ddls.environments.ramp_cluster.ramp_cluster_environment import RampClusterEnvironment as Ref
    from ddls_b200.host.cluster import RampClusterEnvironment as Mine
    for method in ('__init__', 'reset', 'step', 'is_done'):
        ref = list(inspect.signature(getattr(Ref, method)).parameters.values())
        mine = list(inspect.signature(getattr(Mine, method)).parameters.values())
        assert len(mine) >= len(ref), method
        for r, m in zip(ref, mine):                                # same names, order and defaults; extras only at the end
            assert (r.name, r.default, r.kind) == (m.name, m.default, m.kind), (method, r, m)
        for extra in mine[len(ref):]:
            assert extra.default is not inspect.Parameter.empty, (method, extra)
