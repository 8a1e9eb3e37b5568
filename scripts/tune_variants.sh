#!/bin/bash
# usage (on the GPU box): scripts/tune_variants.sh  -- benches every prebuilt library variant under build_variants/
cp ddls_b200/libramp_b200.so /tmp/lib_orig.so
for lib in build_variants/lib_*.so; do
  cp $lib ddls_b200/libramp_b200.so
  echo -n "== $(basename $lib): "
  timeout 120 python bench.py --steps 32 --warmup 8 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
done
cp /tmp/lib_orig.so ddls_b200/libramp_b200.so
