"""One launch of the thread-per-lookahead kernel for ncu (run on the GPU box):
    ncu --set full --clock-control none --import-source on -k regex:lookahead_thread -s 1 -c 1 -o gpurun_out/prof_thread python scripts/profile_thread.py [n]
"""
import sys

import numpy as np

sys.path.insert(0, '.')
from ddls_b200 import synth, engine
from ddls_b200.template_builder import build_template, RampShape

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mode = sys.argv[2] if len(sys.argv) > 2 else 'reference'
g = synth.resnet_like_graph()
eng = engine.RampEngine(n_episodes=4096, n_cluster_workers=64, max_jobs=1, trace_cap=4096)
tids = [eng.register_template(build_template(g, d, RampShape(4, 4, 4), run_times=mode)) for d in (2, 4, 8, 16)]
ids = np.array([tids[k % 4] for k in range(n)], dtype=np.int32)
eng.run_lookaheads(ids)
res, ms = eng.run_lookaheads(ids)
print(ms, int(res['n_ticks'].max()))
