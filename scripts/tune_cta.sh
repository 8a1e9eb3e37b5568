#!/bin/bash
# builds the library with different CTA-kernel register budgets and probes latency at bench-like item counts (GPU box)
set -e
for MW in 16 24 32; do
  for FC in 512 256; do
    python -c "
import sys; sys.path.insert(0,'.')
from ddls_b200 import build
build.build(force=True, extra_flags=['-DRAMP_CTA_MIN_WARPS=$MW','-DRAMP_CTA_F_CAP=$FC'])
" > /dev/null 2>&1
    for NT in 64 128; do
      echo "== min_warps=$MW f_cap=$FC cta_threads=$NT"
      RAMP_LOOKAHEAD_MODE=cta RAMP_LOOKAHEAD_CTA_THREADS=$NT timeout 120 python - <<'PY'
import sys, json
import numpy as np
sys.path.insert(0, '.')
from ddls_b200 import synth, engine
from ddls_b200.template_builder import build_template, RampShape
t = build_template(synth.resnet_like_graph(), 16, RampShape(4, 4, 4))
eng = engine.RampEngine(n_episodes=1, n_cluster_workers=64, max_jobs=1, trace_cap=4096)
tid = eng.register_template(t)
eng.run_lookaheads(np.full(64, tid, dtype=np.int32))
out = []
for n in (1, 592, 920, 1184, 2368, 4096):
    ids = np.full(n, tid, dtype=np.int32)
    best = min(eng.run_lookaheads(ids)[1] for _ in range(3))
    out.append((n, round(best, 2)))
print(out)
PY
    done
  done
done
