"""Flat-array form of one partitioned + placed + scheduled job: the complete input of
the reference's ``RampClusterEnvironment._run_lookahead``
(ddls/environments/ramp_cluster/ramp_cluster_environment.py:379-467).

Index conventions (they carry the reference's tie-break rules):
  * op index  = rank of the op id in ``sorted(graph.nodes)``;
  * dep index = rank of the ``(u, v, k)`` tuple in ``sorted(graph.edges)``.
    Because tuples sort by ``u`` first, the CSR-by-source position of a dep *is* its index.
  * "first in sorted order wins priority ties" (RCE:56-66, RCE:672-685) == lowest index wins.
  * worker / channel ids are job-local dense ids; ``worker_ids`` / ``channel_ids`` map them back
    to the cluster's global string ids.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

NO_CHANNEL = 0xFFFF


@dataclass
class MountScalars:
    """Per-mount scalars read off the partitioned ``Job.details`` (RCE:815, RCE:966-973)."""
    max_acceptable_jct: float = float('inf')
    part_op_mem: float = 0.0
    part_dep_size: float = 0.0
    flow_size: float = 0.0
    n_mounted_workers: int = 0
    n_mounted_channels: int = 0


@dataclass
class LoweredJob:
    n_ops: int
    n_deps: int
    n_workers: int
    n_channels: int
    num_training_steps: int
    model_id: int
    degree: int
    op_cost: np.ndarray        # f64[N]  compute_cost[device_type] -> initial remaining_run_time (RCE:1334)
    op_prio: np.ndarray        # i64[N]  worker.mounted_job_op_to_priority (RCE:1397)
    op_worker: np.ndarray      # u16[N]  job-local worker id (RCE:1336)
    op_n_parents: np.ndarray   # u16[N]  predecessors that are not also successors (JOB:508-523)
    row_ptr: np.ndarray        # i32[N+1]
    dep_dst: np.ndarray        # i32[E]
    dep_run_time: np.ndarray   # f64[E]  init_run_time after RCE:542-560
    dep_prio: np.ndarray       # i64[E]  channel.mounted_job_dep_to_priority (RCE:1412)
    dep_channel: np.ndarray    # u16[E]  job-local channel id or NO_CHANNEL
    dep_is_flow: np.ndarray    # u8[E]   RCE:531-536
    mount: MountScalars = field(default_factory=MountScalars)
    # host-side metadata (not shipped to the device)
    model: str = ''
    op_ids: Optional[List] = None
    dep_ids: Optional[List] = None
    worker_ids: Optional[List] = None    # local -> global worker id
    channel_ids: Optional[List] = None   # local -> global channel id

    ARRAYS = (('op_cost', np.float64), ('op_prio', np.int64), ('op_worker', np.uint16),
              ('op_n_parents', np.uint16), ('row_ptr', np.int32), ('dep_dst', np.int32),
              ('dep_run_time', np.float64), ('dep_prio', np.int64), ('dep_channel', np.uint16),
              ('dep_is_flow', np.uint8))

    def canonicalise(self):
        """Contiguous arrays of the wire dtypes; validates shapes and value ranges."""
        for name, dt in self.ARRAYS:
            a = np.ascontiguousarray(getattr(self, name), dtype=dt)
            setattr(self, name, a)
        N, E = self.n_ops, self.n_deps
        assert self.op_cost.shape == (N,) and self.op_prio.shape == (N,) and self.op_worker.shape == (N,)
        assert self.op_n_parents.shape == (N,) and self.row_ptr.shape == (N + 1,)
        for name in ('dep_dst', 'dep_run_time', 'dep_prio', 'dep_channel', 'dep_is_flow'):
            assert getattr(self, name).shape == (E,), name
        if N:
            if self.row_ptr[0] != 0 or self.row_ptr[-1] != E or np.any(np.diff(self.row_ptr) < 0):
                raise Exception('LoweredJob.row_ptr is not a valid CSR offset array')
            if int(self.op_worker.max()) >= self.n_workers:
                raise Exception('LoweredJob.op_worker refers to a worker >= n_workers')
        if E:
            if self.dep_dst.min() < 0 or self.dep_dst.max() >= N:
                raise Exception('LoweredJob.dep_dst out of range')
            ch = self.dep_channel[self.dep_channel != NO_CHANNEL]
            if ch.size and int(ch.max()) >= self.n_channels:
                raise Exception('LoweredJob.dep_channel refers to a channel >= n_channels')
        if np.any(~(self.op_cost >= 0)) or np.any(~(self.dep_run_time >= 0)):
            raise Exception('LoweredJob run times must be non-negative and not NaN')
        return self

    def fingerprint(self) -> int:
        """64-bit content hash of everything the lookahead result depends on."""
        h = hashlib.blake2b(digest_size=8)
        h.update(np.array([self.n_ops, self.n_deps, self.n_workers, self.n_channels,
                           self.num_training_steps], dtype=np.int64).tobytes())
        for name, _ in self.ARRAYS:
            h.update(getattr(self, name).tobytes())
        return int.from_bytes(h.digest(), 'little')

    def algorithmic_bytes(self, n_ticks: int) -> int:
        """bytes one un-memoised lookahead must move:
        20 B per op (cost 8 + priority key 4 + worker 2 + n_parents 2 + row_ptr 4),
        19 B per dep (run_time 8 + priority key 4 + dst 4 + channel 2 + is_flow 1),
        12 B per tick of trace written, 24 B of results."""
        return 20 * self.n_ops + 19 * self.n_deps + 12 * n_ticks + 24

    # ---- (de)serialisation for fixtures -------------------------------------------------
    def to_npz_dict(self, prefix=''):
        d = {prefix + 'scalars': np.array([self.n_ops, self.n_deps, self.n_workers, self.n_channels,
                                           self.num_training_steps, self.model_id, self.degree], dtype=np.int64),
             prefix + 'mount': np.array([self.mount.max_acceptable_jct, self.mount.part_op_mem,
                                         self.mount.part_dep_size, self.mount.flow_size,
                                         self.mount.n_mounted_workers, self.mount.n_mounted_channels], dtype=np.float64)}
        for name, _ in self.ARRAYS:
            d[prefix + name] = getattr(self, name)
        return d

    @classmethod
    def from_npz_dict(cls, d, prefix=''):
        s = [int(x) for x in d[prefix + 'scalars']]
        m = d[prefix + 'mount']
        kw = {name: np.array(d[prefix + name]) for name, _ in cls.ARRAYS}
        job = cls(n_ops=s[0], n_deps=s[1], n_workers=s[2], n_channels=s[3], num_training_steps=s[4],
                  model_id=s[5], degree=s[6],
                  mount=MountScalars(float(m[0]), float(m[1]), float(m[2]), float(m[3]), int(m[4]), int(m[5])), **kw)
        return job.canonicalise()
