"""ctypes binding of the native template expansion (include/ramp_b200.h: ramp_expand_template; csrc/ramp_expand.cpp).

Same inputs and outputs as ``template_builder.build_template`` -- forward graph + partition degree + a block of servers ->
LoweredJob -- computed in C++ (host only: works without a GPU).  SURVEY.md 8f-1."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import engine
from .lowered import LoweredJob, MountScalars
from .synth import ForwardGraph
from .template_builder import RampShape, REFERENCE_BLOCK_4x4x4


class _FwdGraph(C.Structure):
    _fields_ = [('n_fwd', C.c_int32), ('n_edges', C.c_int32), ('fwd_cost', C.c_void_p), ('bwd_cost', C.c_void_p),
                ('act_size', C.c_void_p), ('par_size', C.c_void_p), ('edge_src', C.c_void_p), ('edge_dst', C.c_void_p)]


class _Aux(C.Structure):
    _fields_ = [('dep_size', C.c_void_p), ('op_mem', C.c_void_p), ('node_order', C.c_void_p)]


class _Block(C.Structure):
    _fields_ = [('n_servers', C.c_int32), ('num_communication_groups', C.c_int32), ('coords', C.c_void_p),
                ('channel_bandwidth', C.c_double), ('latency', C.c_double), ('io_latency', C.c_double)]


def block_coords(shape: RampShape, degree: int, block_start: int = 0, run_times: str = 'one_to_one'):
    """(cg, rack, server) of the servers sub-op 0..degree-1 go to: the probed reference block on an empty 4x4x4 cluster for
    run_times='reference', else the aligned block [block_start, block_start + degree) of template_builder."""
    if run_times == 'reference':
        if (shape.c, shape.r, shape.s) != (4, 4, 4) or degree not in REFERENCE_BLOCK_4x4x4 or block_start != 0:
            raise Exception("run_times='reference' is available for the probed empty-cluster blocks of a 4x4x4 RAMP only")
        return list(REFERENCE_BLOCK_4x4x4[degree])
    out = []
    for w in range(block_start, block_start + max(degree, 1)):
        c, rem = divmod(w, shape.r * shape.s)
        r, s = divmod(rem, shape.s)
        out.append((c, r, s))
    return out


def expand_template(fwd: ForwardGraph, degree: int, shape: RampShape, block_start: int = 0, quantum: float = 0.01,
                    num_training_steps: int = 50, model_id: int = 0, max_acceptable_frac: float = 1.0,
                    run_times: str = 'one_to_one', coords=None) -> LoweredJob:
    """coords: explicit (cg, rack, server) of the block's servers in sorted server-id order (e.g. from placer.first_fit_place);
    overrides block_start / the probed empty-cluster blocks."""
    L = engine.load_library()
    L.ramp_expand_template.restype = C.c_int
    L.ramp_expand_template.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.ramp_free_expanded_aux.restype = None
    L.ramp_free_expanded_aux.argtypes = [C.c_void_p]
    L.ramp_free_expanded_job.restype = None
    L.ramp_free_expanded_job.argtypes = [C.c_void_p]
    if degree != 1 and degree % 2 != 0:
        raise Exception(f'Invalid num_partitions={degree}; RAMP placer expects even numbers.')   # op_partition.py:26-27
    if coords is None and block_start + max(degree, 1) > shape.n_workers:
        raise Exception('worker block does not fit in the cluster')
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    fc, bc, ac, pc = f64(fwd.fwd), f64(fwd.bwd), f64(fwd.act), f64(fwd.par)
    es = np.ascontiguousarray([u for (u, _) in fwd.edges], dtype=np.int32)
    ed = np.ascontiguousarray([v for (_, v) in fwd.edges], dtype=np.int32)
    g = _FwdGraph(fwd.n, len(fwd.edges), fc.ctypes.data, bc.ctypes.data, ac.ctypes.data, pc.ctypes.data, es.ctypes.data, ed.ctypes.data)
    coords = np.ascontiguousarray(block_coords(shape, degree, block_start, run_times) if coords is None else sorted(coords),
                                  dtype=np.int32).reshape(-1, 3)
    blk = _Block(len(coords), shape.c, coords.ctypes.data, shape.channel_bandwidth, shape.latency, shape.io_latency)
    mode = 1 if run_times == 'reference' else 0
    out, aux = engine._LoweredJob(), _Aux()
    engine._check(L.ramp_expand_template(C.byref(g), degree, quantum, C.byref(blk), mode, num_training_steps, C.byref(out), C.byref(aux)))
    N, E = out.n_ops, out.n_deps

    def arr(ptr, n, dt):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy() if n else np.zeros(0, dtype=dt)
    try:
        sizes, op_mem, order = arr(aux.dep_size, E, np.float64), arr(aux.op_mem, N, np.float64), arr(aux.node_order, N, np.int32)
        op_cost = arr(out.op_cost, N, np.float64)
        is_flow = arr(out.dep_is_flow, E, np.uint8)
        seq_time = float(sum(float(op_cost[i]) for i in order)) * num_training_steps            # JOB:224-235, graph order
        mount = MountScalars(max_acceptable_jct=max_acceptable_frac * seq_time,
                             part_op_mem=float(sum(float(op_mem[i]) for i in order)),
                             part_dep_size=float(sizes.sum()), flow_size=float(sizes[is_flow == 1].sum()),
                             n_mounted_workers=out.n_workers, n_mounted_channels=out.n_channels)
        lj = LoweredJob(n_ops=N, n_deps=E, n_workers=out.n_workers, n_channels=out.n_channels,
                        num_training_steps=num_training_steps, model_id=model_id, degree=degree, op_cost=op_cost,
                        op_prio=arr(out.op_prio, N, np.int64), op_worker=arr(out.op_worker, N, np.uint16),
                        op_n_parents=arr(out.op_n_parents, N, np.uint16), row_ptr=arr(out.row_ptr, N + 1, np.int32),
                        dep_dst=arr(out.dep_dst, E, np.int32), dep_run_time=arr(out.dep_run_time, E, np.float64),
                        dep_prio=arr(out.dep_prio, E, np.int64), dep_channel=arr(out.dep_channel, E, np.uint16),
                        dep_is_flow=is_flow, mount=mount, model=fwd.name)
    finally:
        L.ramp_free_expanded_job(C.byref(out))
        L.ramp_free_expanded_aux(C.byref(aux))
    lj.seq_time = seq_time
    return lj.canonicalise()
