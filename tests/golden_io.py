"""Loader for the fixtures written by oracle/gen_golden.py."""
import os

import numpy as np

from ddls_b200.lowered import LoweredJob, MountScalars

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class Golden:
    def __init__(self, name):
        path = os.path.join(GOLDEN_DIR, name if name.endswith('.npz') else name + '.npz')
        self.d = np.load(path)
        self.name = name
        self.templates = [LoweredJob.from_npz_dict(self.d, prefix=f't{t}_') for t in range(int(self.d['n_templates']))]
        self.n_lookaheads = int(self.d['n_lookaheads'])
        self.n_cluster_workers = int(self.d['meta_n_cluster_workers'])
        self.max_sim_time = float(self.d['meta_max_sim_time'])
        self.n_models = int(self.d['meta_n_models'])

    def lookahead(self, i):
        res = self.d[f'la{i}_res']
        return dict(tid=int(self.d[f'la{i}_tid']), jct=float(res[0]), comm=float(res[1]), comp=float(res[2]),
                    trace_n=self.d[f'la{i}_trace_n'], trace_tick=self.d[f'la{i}_trace_tick'])

    def step_job(self, s):
        """LoweredJob (with this step's mount scalars) for cluster step s, or None for Action()."""
        tid = int(self.d['step_tid'][s])
        if tid < 0:
            return None
        import copy
        job = copy.copy(self.templates[tid])
        m = self.d['step_mount'][s]
        job.mount = MountScalars(float(m[0]), float(m[1]), float(m[2]), float(m[3]), int(m[4]), int(m[5]))
        return job

    @property
    def n_steps(self):
        return len(self.d['step_tid'])

    def arrivals(self):
        from oracle.oracle import ARRIVAL_DTYPE
        a = self.d['arrivals']
        out = np.zeros(len(a), dtype=ARRIVAL_DTYPE)
        out['interarrival'], out['orig_op_mem'], out['orig_dep_size'] = a[:, 0], a[:, 1], a[:, 2]
        return out
