"""Latency of ONE wave of identical lookaheads per partition degree on the thread-per-lookahead kernel (run on the GPU box).

    python scripts/microbench_thread_degree.py [run_times] [n]
"""
import json
import sys

import numpy as np

sys.path.insert(0, '.')
from ddls_b200 import synth, engine, workload
from ddls_b200.template_builder import build_template, RampShape

mode = sys.argv[1] if len(sys.argv) > 1 else 'reference'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
g = synth.resnet_like_graph()
eng = engine.RampEngine(n_episodes=4096, n_cluster_workers=64, max_jobs=1, trace_cap=4096)
for d in (2, 4, 8, 16):
    lj = workload.reference_template(g, d, RampShape(4, 4, 4)) if mode == 'reference' else build_template(g, d, RampShape(4, 4, 4), run_times=mode)
    tid = eng.register_template(lj)
    ids = np.full(n, tid, dtype=np.int32)
    eng.run_lookaheads(ids)
    best = min(eng.run_lookaheads(ids)[1] for _ in range(5))
    res, _ = eng.run_lookaheads(ids)
    T = int(res['n_ticks'].max())
    print(json.dumps(dict(mode=mode, degree=d, n=n, ms=round(best, 4), T=T, us_per_tick=round(best * 1e3 / T, 4))), flush=True)
