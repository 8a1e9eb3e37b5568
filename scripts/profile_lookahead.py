"""One standalone lookahead-kernel launch on the BASELINE-sized template, for ncu (run on the GPU box)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from ddls_b200 import synth, engine
from ddls_b200.template_builder import build_template, RampShape
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
degree = int(sys.argv[2]) if len(sys.argv) > 2 else 16
t = build_template(synth.resnet_like_graph(), degree, RampShape(4, 4, 4))
eng = engine.RampEngine(n_episodes=1, n_cluster_workers=64, max_jobs=1, trace_cap=4096)
tid = eng.register_template(t)
ids = np.full(n, tid, dtype=np.int32)
for _ in range(3):
    res, ms = eng.run_lookaheads(ids)
print('ms', ms, 'T', int(res['n_ticks'][0]))
