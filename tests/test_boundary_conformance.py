"""The drop-in class keeps the reference's call signatures.  Needs the read-only reference checkout (through
the oracle's import shim), so it runs in the build container only and is skipped on the GPU box."""
import inspect
import os
import sys

import pytest

REF = '/root/reference'


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present (GPU box)')
def test_dropin_class_signatures_match_reference():
    from oracle import ref_shim
    if hasattr(ref_shim, 'install'):
        ref_shim.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
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
