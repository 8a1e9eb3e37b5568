"""Latency / throughput of one degree at a few concurrency levels (GPU box): argv = degree, n..."""
import sys, json
import numpy as np
sys.path.insert(0, '.')
from ddls_b200 import synth, engine
from ddls_b200.template_builder import build_template, RampShape
degree = int(sys.argv[1])
ns = [int(x) for x in sys.argv[2:]] or [1, 296, 592]
t = build_template(synth.resnet_like_graph(), degree, RampShape(4, 4, 4))
eng = engine.RampEngine(n_episodes=1, n_cluster_workers=64, max_jobs=1, trace_cap=4096)
tid = eng.register_template(t)
eng.run_lookaheads(np.full(64, tid, dtype=np.int32))
out = []
for n in ns:
    ids = np.full(n, tid, dtype=np.int32)
    out.append((n, round(min(eng.run_lookaheads(ids)[1] for _ in range(3)), 3)))
print(degree, out, flush=True)
