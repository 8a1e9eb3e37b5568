"""ctypes binding of ``ramp_quotient_template`` (ddls_b200/csrc/ramp_quotient.cpp): the symmetry quotient
``ramp_register_template`` applies to every lowered job before it goes to the device.  Host-only; exported for
inspection and for the CPU tests (tests/test_quotient.py)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import engine
from .lowered import LoweredJob


class _Quotient(C.Structure):
    _fields_ = [('n_ops', C.c_int32), ('n_deps', C.c_int32), ('n_workers', C.c_int32), ('n_channels', C.c_int32),
                ('op_cost', C.POINTER(C.c_double)), ('op_key', C.POINTER(C.c_uint32)), ('op_worker', C.POINTER(C.c_uint32)),
                ('op_weight', C.POINTER(C.c_uint32)), ('op_threshold', C.POINTER(C.c_uint32)),
                ('row_ptr', C.POINTER(C.c_int32)), ('dep_dst', C.POINTER(C.c_int32)),
                ('dep_run_time', C.POINTER(C.c_double)), ('dep_key', C.POINTER(C.c_uint32)),
                ('dep_channel', C.POINTER(C.c_uint32)), ('dep_group_mask', C.POINTER(C.c_uint64)), ('dep_is_flow', C.POINTER(C.c_uint8)),
                ('dep_inc', C.POINTER(C.c_uint32)), ('op_class', C.POINTER(C.c_int32)), ('dep_entry', C.POINTER(C.c_int32)),
                ('merged', C.c_int32), ('masks_valid', C.c_int32)]


@dataclass
class QuotientJob:
    n_ops: int
    n_deps: int
    n_workers: int
    n_channels: int
    num_training_steps: int
    op_cost: np.ndarray
    op_key: np.ndarray
    op_worker: np.ndarray
    op_weight: np.ndarray
    op_threshold: np.ndarray
    row_ptr: np.ndarray
    dep_dst: np.ndarray
    dep_run_time: np.ndarray
    dep_key: np.ndarray
    dep_channel: np.ndarray      # split entries: channel group; 0xFFFFFFFF = none or merged
    dep_is_flow: np.ndarray
    dep_inc: np.ndarray
    op_class: np.ndarray
    dep_entry: np.ndarray
    dep_group_mask: np.ndarray = None   # uint64: bit g set = members on group g
    merged: int = 0
    masks_valid: int = 1


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def quotient(job: LoweredJob) -> QuotientJob:
    L = engine.load_library()
    L.ramp_quotient_template.restype = C.c_int
    L.ramp_quotient_template.argtypes = [C.POINTER(engine._LoweredJob), C.POINTER(_Quotient)]
    L.ramp_free_quotient.restype = None
    L.ramp_free_quotient.argtypes = [C.POINTER(_Quotient)]
    job.canonicalise()
    cj = engine._LoweredJob(job.n_ops, job.n_deps, job.n_workers, job.n_channels, job.num_training_steps,
                            job.model_id, job.degree, 0,
                            job.op_cost.ctypes.data, job.op_prio.ctypes.data, job.op_worker.ctypes.data,
                            job.op_n_parents.ctypes.data, job.row_ptr.ctypes.data, job.dep_dst.ctypes.data,
                            job.dep_run_time.ctypes.data, job.dep_prio.ctypes.data, job.dep_channel.ctypes.data,
                            job.dep_is_flow.ctypes.data)
    q = _Quotient()
    rc = L.ramp_quotient_template(C.byref(cj), C.byref(q))
    if rc != 0:
        raise Exception(f'ramp_quotient_template failed ({rc})')
    try:
        n, e = q.n_ops, q.n_deps
        return QuotientJob(n_ops=n, n_deps=e, n_workers=q.n_workers, n_channels=q.n_channels,
                           num_training_steps=job.num_training_steps,
                           op_cost=_arr(q.op_cost, n, np.float64), op_key=_arr(q.op_key, n, np.int64),
                           op_worker=_arr(q.op_worker, n, np.int64), op_weight=_arr(q.op_weight, n, np.int64),
                           op_threshold=_arr(q.op_threshold, n, np.int64), row_ptr=_arr(q.row_ptr, n + 1, np.int64),
                           dep_dst=_arr(q.dep_dst, e, np.int64), dep_run_time=_arr(q.dep_run_time, e, np.float64),
                           dep_key=_arr(q.dep_key, e, np.int64), dep_channel=_arr(q.dep_channel, e, np.int64),
                           dep_is_flow=_arr(q.dep_is_flow, e, np.uint8), dep_inc=_arr(q.dep_inc, e, np.int64),
                           op_class=_arr(q.op_class, job.n_ops, np.int64), dep_entry=_arr(q.dep_entry, job.n_deps, np.int64),
                           dep_group_mask=_arr(q.dep_group_mask, e, np.uint64), merged=int(q.merged), masks_valid=int(q.masks_valid))
    finally:
        L.ramp_free_quotient(C.byref(q))
