"""A few bench-like steps for ncu (run on the GPU box): the cfg3 workload with reference run times, device-resident actions.

    ncu --set full --clock-control none --import-source on -k regex:lookahead_thread -s 2 -c 3 -o gpurun_out/prof_bench_r2 python scripts/profile_bench_step.py
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python scripts/profile_bench_step.py
Prints the number of lookaheads every step executed (for scripts/ncu_summary.py)."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from ddls_b200 import engine, workload

cfg = workload.CONFIGS['cfg3-resnet50-64w']
B, L = 4096, 8
eng = engine.RampEngine(n_episodes=B, n_cluster_workers=64, max_jobs=L, trace_cap=4096)
tmap = {}


def jcts(templates):
    for i, t in enumerate(templates):
        tmap[i] = eng.register_template(t)
    res, _ = eng.run_lookaheads([tmap[i] for i in range(len(templates))])
    return res['jct']


wl = workload.generate('cfg3-resnet50-64w', jcts, n_episodes=B, n_steps=L, seed=0, run_times='reference')
dev = []
for p in range(L):
    a = wl.actions[p].copy()
    placed = a['template_id'] >= 0
    a['template_id'][placed] = np.array([tmap[int(t)] for t in a['template_id'][placed]], dtype=np.int32)
    dev.append(torch.from_numpy(a.view(np.uint8).reshape(B, -1).copy()).cuda())
stats = torch.empty((B, engine.STEP_STATS_LEN), dtype=torch.float64, device='cuda')
ncs = torch.empty(B, dtype=torch.int32, device='cuda')
eng.reset(wl.arrivals)
prev = 0
for p in range(L):
    eng.step_device(dev[p].data_ptr(), True, stats.data_ptr(), ncs.data_ptr())
    eng.sync()
    m = eng.memo_stats()
    print('step', p, 'lookaheads', m['lookaheads'] - prev, flush=True)
    prev = m['lookaheads']
