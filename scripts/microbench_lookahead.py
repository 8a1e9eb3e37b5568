"""Sweeps the lookahead kernel's CTA size / residency on the BASELINE-sized template (run on the GPU box).

    python scripts/microbench_lookahead.py [n_lookaheads] [degree]
"""
import json
import os
import subprocess
import sys

CHILD = r'''
import sys, json, time
import numpy as np
sys.path.insert(0, '.')
from ddls_b200 import synth, engine
from ddls_b200.template_builder import build_template, RampShape
n, degree = int(sys.argv[1]), int(sys.argv[2])
t = build_template(synth.resnet_like_graph(), degree, RampShape(4, 4, 4))
eng = engine.RampEngine(n_episodes=1, n_cluster_workers=64, max_jobs=1, trace_cap=4096)
tid = eng.register_template(t)
ids = np.full(n, tid, dtype=np.int32)
eng.run_lookaheads(ids[:256])
best = 1e9
for _ in range(3):
    res, ms = eng.run_lookaheads(ids)
    best = min(best, ms)
assert (res['status'] == 0).all()
T = int(res['n_ticks'][0])
print(json.dumps(dict(ms=best, n=n, N=t.n_ops, E=t.n_deps, T=T, jct=float(res['jct'][0]),
                      lookaheads_per_s=n / best * 1e3, alg_GBps=n * t.algorithmic_bytes(T) / best / 1e6)))
'''

if __name__ == '__main__':
    n = sys.argv[1] if len(sys.argv) > 1 else '4096'
    degree = sys.argv[2] if len(sys.argv) > 2 else '16'
    for nt in (32, 64, 128, 256):
        for cps in (0,):
            env = dict(os.environ, RAMP_LOOKAHEAD_THREADS=str(nt))
            if cps:
                env['RAMP_LOOKAHEAD_CTAS_PER_SM'] = str(cps)
            else:
                env.pop('RAMP_LOOKAHEAD_CTAS_PER_SM', None)
            try:
                out = subprocess.run([sys.executable, '-c', CHILD, n, degree], env=env, capture_output=True, text=True, timeout=600)
                line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
            except subprocess.TimeoutExpired:
                line = 'timeout'
            print(f'nt={nt} ctas_per_sm={cps or "max"}: {line}', flush=True)
