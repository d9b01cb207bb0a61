#!/bin/bash
# One gpurun call that produces every artefact under profiles/ (run: gpurun --timeout 1500 -- 'bash scripts/profile_round.sh r01').
# Each ncu run is preceded (&&) by the same command without ncu, as the profiling recipe requires.
set -u
R=${1:-r01}
O=gpurun_out
mkdir -p $O
# 1. launch list of the benchmark command (eager launches: one row per kernel; cold-cache, serialised)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/${R}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/${R}_ncu_launches.log 2>&1
# 2. full-set capture of every library kernel of ONE forward (second forward of the script: warm caches / packs)
python scripts/profile_forward.py > $O/${R}_plain_fwd.log 2>&1 &&
SKIP=$(grep -o "skip=[0-9]*" $O/${R}_plain_fwd.log | cut -d= -f2) && COUNT=$(grep -o "count=[0-9]*" $O/${R}_plain_fwd.log | cut -d= -f2) &&
ncu --set full --clock-control none --import-source on -k regex:"tgp::" -s $SKIP -c $COUNT -o $O/${R}_forward -f \
    python scripts/profile_forward.py > $O/${R}_ncu_forward.log 2>&1
tail -2 $O/${R}_ncu_forward.log
