"""First-fit op placement of one partitioned job on a (possibly busy) RAMP cluster: which server every sub-op goes to.

Restates RampFirstFitOpPlacer (agents/placers/ramp_first_fit_op_placer.py:27-113) and the helpers it calls
(agents/placers/utils.py: get_allocation_preamble :68, topo_sort :100, parent_collective_placement :258,
regular_collective_placement :333, find_sub_block :385, ff_block :394, get_factor_pairs :445, get_block :464,
get_block_shapes :491, allocate :532, check_block :215) on plain dicts; pinned against the reference's recorded decisions in
tests/test_placer.py (fixture tests/fixtures/placer_cases.json).  SURVEY.md 8f rows 1 and 4 (host side)."""
from __future__ import annotations

import math
from collections import deque
from typing import Dict, List, Optional, Sequence, Tuple

Server = Tuple[int, int, int]


def backward_op_id(forward_op_id: str, num_nodes: int) -> str:
    return str((2 * num_nodes) - (int(forward_op_id) - 1))            # placers/utils.py:316-322


def partitioned_op_id(op_id: str, split_id: int) -> str:
    return str(int(op_id)) + chr(97 + split_id)                       # placers/utils.py:324-331


def _factor_pairs(n):
    return [(n // i, i) for i in range(1, n + 1) if n % i == 0]       # placers/utils.py:445-462


def _block_shapes(pairs, meta):
    blocks = []
    for p0, p1 in pairs:                                              # placers/utils.py:491-530
        var = math.sqrt(p0)
        if (var % 1 == 0) and (var <= meta[0] and var <= meta[1] and p1 <= meta[2]):
            blocks.append((int(var), int(var), p1))
        if p0 > meta[0] or p0 > meta[1] or p1 > meta[2]:
            continue
        blocks.append((p0, 1, p1))
        blocks.append((p0, p1, 1))
    return blocks


def _get_block(C, R, S, shape, origin):
    i, j, k = origin                                                  # placers/utils.py:464-489
    if S == -1:
        return [((i + n) % (shape[0] + 1), (j + n) % (shape[1] + 1), k % shape[2]) for n in range(C)]
    return [((i + c) % shape[0], (j + r) % shape[1], (k + s) % shape[2]) for c in range(C) for r in range(R) for s in range(S)]


def _check_block(ramp, block, op_size, job_idx):
    if not block:                                                     # placers/utils.py:215-233
        return False
    for server in block:
        if len(ramp[server]['job_idxs']) != 0 and job_idx not in ramp[server]['job_idxs']:
            return False
        if ramp[server]['mem'] < op_size:
            return False
    return True


def _ff_block(block_shapes, meta_shape, shape, ramp, job_idx, op_size):
    for bs in block_shapes:                                           # placers/utils.py:394-443 (origin (0, 0, 0))
        I, J, K = (meta_shape[0] - bs[0]) + 1, (meta_shape[1] - bs[1]) + 1, (meta_shape[2] - bs[2]) + 1
        if I <= 0 or J <= 0 or K <= 0:
            continue
        for i in range(I):
            for j in range(J):
                for k in range(K):
                    block = _get_block(bs[0], bs[1], bs[2], shape, (i, j, k))
                    if _check_block(ramp, block, op_size, job_idx):
                        return block
    return None


def first_fit_place(nodes: Sequence[str], mem: Sequence[float], in_edges: Dict[str, List[str]], out_edges: Dict[str, List[str]],
                    split_of: Dict[str, int], ramp: Dict[Server, dict], shape: Server, servers: Sequence[Server],
                    job_idx: int) -> Optional[Dict[str, Server]]:
    """nodes / mem / in_edges / out_edges: the job's un-partitioned forward graph in networkx iteration order; split_of: forward
    op id -> number of sub-ops (1 if absent); ramp: server -> {'mem': free bytes, 'job_idxs': set of mounted job idxs} (it is
    updated in place like the reference's dummy ramp); servers: every server of the cluster (the meta block).
    Returns op id (forward, backward, sub-ops 'ida'..) -> server, or None when the job cannot be placed."""
    n = len(nodes)
    memory = dict(zip(nodes, mem))
    parents = {v: list(in_edges[v]) for v in nodes}
    children = {v: list(out_edges[v]) for v in nodes}
    # topo_sort (placers/utils.py:100-115) consumes a copy of the parent lists
    left = {v: list(p) for v, p in parents.items()}
    sequence, queue = [], deque()
    for v in nodes:
        if left[v] == []:
            queue.append(v); sequence.append(v)
    while queue:
        v = queue.popleft()
        for c in children[v]:
            left[c].remove(v)
            if left[c] == []:
                queue.append(c); sequence.append(c)
    meta = set(servers)
    where: Dict[str, List[Server]] = {v: [] for v in nodes}
    out: Dict[str, Server] = {}

    def put(op, split, j, server):
        bwd = backward_op_id(op, n)
        if split > 1:
            out[partitioned_op_id(op, j)] = server
            out[partitioned_op_id(bwd, j)] = server
        else:
            out[op] = server
            out[bwd] = server
        where[op].append(server)
    for op in sequence:
        split = split_of.get(op, 1)
        placed = False
        # parent_collective_placement (placers/utils.py:258-314): reuse the servers of a parent split the same number of times
        for servers_p in [where[p] for p in parents[op] if set(where[p]).issubset(meta)]:
            if split != len(servers_p):
                continue
            if sum(ramp[s]['mem'] for s in servers_p) >= memory[op]:
                for j, s in enumerate(servers_p):
                    ramp[s]['mem'] -= memory[op] / split
                    put(op, split, j, s)
                placed = True
                break
        if placed:
            continue
        # regular_collective_placement (placers/utils.py:333-383): first block shape / origin that is free and has the memory
        if split > len(servers):
            return None
        op_size = memory[op] / split
        shapes = _block_shapes(_factor_pairs(split), shape) + [(split, split, -1), (split, 1, 1)]
        block = _ff_block(shapes, shape, shape, ramp, job_idx, op_size)
        if not block:
            return None
        for j, s in enumerate(block):
            ramp[s]['mem'] -= op_size
            put(op, split, j, s)
    return out


def first_fit_place_native(n_fwd: int, mem: Sequence[float], edges: Sequence[Tuple[int, int]], splits: Sequence[int],
                           free_mem: Dict[Server, float], busy: Dict[Server, bool], shape: Server) -> Optional[Dict[str, Server]]:
    """Same decision through the C ABI (include/ramp_b200.h: ramp_first_fit_place; csrc/ramp_expand.cpp).  Forward ops are
    1..n_fwd in node order, `edges` in the graph's edge order; returns the same mapping as first_fit_place."""
    import ctypes as C
    import numpy as np
    from . import engine
    from .expand import _FwdGraph

    class _State(C.Structure):
        _fields_ = [('shape', C.c_int32 * 3), ('_pad', C.c_int32), ('free_mem', C.c_void_p), ('busy', C.c_void_p)]
    L = engine.load_library()
    L.ramp_first_fit_place.restype = C.c_int
    L.ramp_first_fit_place.argtypes = [C.c_void_p] * 5
    c_, r_, s_ = shape
    idx = lambda sv: (sv[0] * r_ + sv[1]) * s_ + sv[2]
    fm = np.zeros(c_ * r_ * s_, dtype=np.float64)
    bz = np.zeros(c_ * r_ * s_, dtype=np.uint8)
    for sv, m in free_mem.items():
        fm[idx(sv)] = m
    for sv, b in busy.items():
        bz[idx(sv)] = 1 if b else 0
    memv = np.ascontiguousarray(mem, dtype=np.float64)
    zero = np.zeros(n_fwd, dtype=np.float64)
    es = np.ascontiguousarray([u for (u, _) in edges], dtype=np.int32)
    ed = np.ascontiguousarray([v for (_, v) in edges], dtype=np.int32)
    g = _FwdGraph(n_fwd, len(edges), zero.ctypes.data, zero.ctypes.data, memv.ctypes.data, zero.ctypes.data, es.ctypes.data, ed.ctypes.data)
    st = _State((C.c_int32 * 3)(c_, r_, s_), 0, fm.ctypes.data, bz.ctypes.data)
    sp = np.ascontiguousarray(splits, dtype=np.int32)
    server_out = np.zeros(int(np.maximum(sp, 1).sum()), dtype=np.int32)
    offset_out = np.zeros(n_fwd + 1, dtype=np.int32)
    rc = L.ramp_first_fit_place(C.byref(g), sp.ctypes.data, C.byref(st), server_out.ctypes.data, offset_out.ctypes.data)
    if rc == 1:
        return None
    engine._check(rc)
    out: Dict[str, Server] = {}
    for i in range(1, n_fwd + 1):
        k = max(int(sp[i - 1]), 1)
        for j in range(k):
            sv = int(server_out[offset_out[i - 1] + j])
            coord = (sv // (r_ * s_), (sv // s_) % r_, sv % s_)
            fwd_id, bwd_id = str(i), backward_op_id(str(i), n_fwd)
            if k > 1:
                out[partitioned_op_id(fwd_id, j)] = coord
                out[partitioned_op_id(bwd_id, j)] = coord
            else:
                out[fwd_id] = coord
                out[bwd_id] = coord
    return out
