"""Per-step breakdown of one scripted segment of the bench workload (GPU box)."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
import torch
from ddls_b200 import engine, workload
cfgname = 'cfg3-resnet50-64w'
cfg = workload.CONFIGS[cfgname]
B, L = 4096, 8
eng = engine.RampEngine(n_episodes=B, n_cluster_workers=64, max_jobs=L, trace_cap=4096)
tmap = {}
def jcts(ts):
    for i, t in enumerate(ts):
        tmap[i] = eng.register_template(t)
    res, _ = eng.run_lookaheads([tmap[i] for i in range(len(ts))])
    return res['jct']
wl = workload.generate(cfgname, jcts, n_episodes=B, n_steps=L, seed=0)
acts = []
for p in range(L):
    a = wl.actions[p].copy()
    m = a['template_id'] >= 0
    a['template_id'][m] = np.array([tmap[int(t)] for t in a['template_id'][m]], dtype=np.int32)
    acts.append(a)
for rep in range(2):
    eng.reset(wl.arrivals)
    rows = []
    for p in range(L):
        eng.lookahead_kernel_time(reset=True)
        t0 = time.perf_counter()
        eng.step(acts[p], fuse_empty_steps=True)
        dt = (time.perf_counter() - t0) * 1e3
        kt = eng.lookahead_kernel_time(reset=True)
        deg = np.bincount(acts[p]['template_id'][acts[p]['template_id'] >= 0], minlength=4)
        rows.append((p, round(dt, 2), round(kt['total_ms'], 2), kt['work_items'], deg.tolist()))
    if rep == 1:
        for r in rows:
            print('step %d: wall %.2f ms, lookahead kernels %.2f ms, %d lookaheads, placed per template %s' % r)
