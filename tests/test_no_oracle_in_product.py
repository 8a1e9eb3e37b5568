"""The oracle is test infrastructure: nothing under ddls_b200/ may import, link or execute it, and the product path has no
CPU fallback (it must raise when the CUDA library is missing -- see test_capi_symbols.test_missing_library_fails_loudly)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_never_touches_the_oracle():
    pat = re.compile(r'^\s*(from|import)\s+oracle\b|libramp_oracle|ramp_oracle\.(c|h)|orc_[a-z_]+\(', re.M)
    offenders = []
    for base, _, files in os.walk(os.path.join(ROOT, 'ddls_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                src = open(os.path.join(base, f), errors='ignore').read()
                if pat.search(src):
                    offenders.append(os.path.join(base, f))
    assert offenders == []


def test_reference_is_not_read_at_run_time():
    """The reference checkout does not exist on the GPU box: only the scripts that run in the build container (golden
    generators, the staging recipe) and oracle/ref_shim.py (which falls back to the staged copy oracle/_ref) may name it."""
    allowed = {os.path.join(ROOT, 'oracle', f) for f in ('gen_golden.py', 'ref_shim.py', 'stage_ref.py')}
    needle = '/root/' + 'reference'
    offenders = []
    for base in ('ddls_b200', 'tests', 'oracle'):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            if os.sep + '_ref' in d:            # the staged, unmodified reference itself
                continue
            for f in files:
                path = os.path.join(d, f)
                if f.endswith('.py') and path not in allowed:
                    if needle in open(path, errors='ignore').read():
                        offenders.append(path)
    for f in ('bench.py', '__graft_entry__.py'):
        if needle in open(os.path.join(ROOT, f)).read():
            offenders.append(f)
    assert offenders == []
