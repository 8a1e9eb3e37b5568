"""TEST INFRASTRUCTURE ONLY -- makes the *unmodified* Python reference importable.

The reference (cwfparsonson/ddls, mounted read-only at /root/reference) imports a
number of third-party modules at module scope that are absent from this image
(ray, gym, dgl, hydra, omegaconf, matplotlib, seaborn, sigfig, pygraphviz,
sqlitedict).  None of them does arithmetic on the hot path
(RampClusterEnvironment.step, ramp_cluster_environment.py:894), so they are
replaced by inert MagicMock packages.  ``ray.remote`` becomes a pass-through
decorator (ramp_cluster_environment.py:39, job.py:19) and ``gym.Env`` a plain
base class (ramp_job_partitioning_environment.py:42).

This file exists so that ``oracle/gen_golden.py`` and the reference cross-check tests can run the
reference *in this container* to pin the C restatement in ``oracle/ramp_oracle.c``, and so that
``bench.py --impl reference`` / the drop-in integration test can run the copy staged at
``oracle/_ref`` (oracle/stage_ref.py) on the GPU box, where /root/reference does not exist.
Nothing in the product path (``ddls_b200/``) imports it.
"""
import importlib.abc
import importlib.machinery
import os
import sys
from unittest.mock import MagicMock

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')     # oracle/stage_ref.py (travels to the GPU box)
REFERENCE_ROOT = os.environ.get('DDLS_REFERENCE_ROOT') or ('/root/reference' if os.path.isdir('/root/reference/ddls') else _STAGED)

MISSING = {'ray', 'sqlitedict', 'dgl', 'omegaconf', 'hydra', 'matplotlib', 'seaborn',
           'pygraphviz', 'sigfig', 'gym', 'wandb'}


class _Loader(importlib.abc.Loader):
    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        if name.split('.')[0] in MISSING:
            return importlib.machinery.ModuleSpec(name, _Loader(), is_package=True)
        return None


_installed = False


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'ddls'))


def install():
    """Idempotently installs the import shim and puts the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f'reference not found under {REFERENCE_ROOT}')
    really_missing = set()
    for name in sorted(MISSING):
        try:
            __import__(name)
        except Exception:
            really_missing.add(name)
    MISSING.intersection_update(really_missing)
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, REFERENCE_ROOT)
    import ray
    import gym
    if 'ray' in MISSING:
        ray.remote = lambda f=None, **kw: (f if f is not None else (lambda g: g))
        ray.init = lambda **kw: None
    if 'gym' in MISSING:
        gym.Env = type('Env', (), {})
    _installed = True
