#!/bin/bash
# usage (GPU box): scripts/final_records.sh <tag>  -- the measurement records committed under profiles/
TAG=${1:-v7}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 32 --warmup 8 > gpurun_out/bench_1gpu_$TAG.json 2> gpurun_out/bench_1gpu_$TAG.err
tail -c 300 gpurun_out/bench_1gpu_$TAG.json; echo
timeout 300 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -c 300 gpurun_out/bench_ref_$TAG.json; echo
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
tail -2 gpurun_out/launches_$TAG.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ramp_lookahead -s 2 -c 13 -o gpurun_out/prof_bench_$TAG python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out/prof_bench_$TAG.ncu-rep
