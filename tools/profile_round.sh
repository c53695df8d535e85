#!/bin/bash
# usage (under gpurun, one GPU): bash tools/profile_round.sh <tag>
# Writes into gpurun_out/:  <tag>_bench.json (the default bench line, never run under a profiler),
# <tag>_launches.csv (ncu launch list of `bench.py --steps 2 --warmup 1`, per-launch durations) and
# <tag>_full.ncu-rep (ncu --set full of one encode + one decode launch at full size, with source).
set -u
tag=${1:-r2}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(encode|decode|scan|compact)' -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(en|de)code_ops_wide' --launch-skip 2 -c 2 -f -o gpurun_out/${tag}_full \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-configs > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/ | tail -8
