#!/bin/bash
# One gpurun call that produces every artefact under profiles/ (run: gpurun --timeout 1500 -- 'bash scripts/profile_round.sh r01').
# Each ncu run is preceded (&&) by the same command without ncu, as the profiling recipe requires.
set -u
R=${1:-r01}
O=gpurun_out
mkdir -p $O
if [ "${2:-all}" = "all" ]; then
# 1. launch list of the benchmark command (eager launches: one row per kernel; cold-cache, serialised)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-train > $O/${R}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-train > $O/${R}_ncu_launches.log 2>&1
fi
# (ncu matches the base name, without the tgp:: namespace)
KERNELS="^(concat_rows|decode_max|edge_record|gather_max|gather_rows|gemm_naive|gemm_skinny|gemm_simt|gemm_tc|knn_tc|knn_xyz|knn_feat|layer_conv|nearest|orl_|rownorm|select_rows|split_tf32|split_mixed|surface_conv|direction_norm)"
# 2. every library kernel of ONE forward (third forward of the script: warm caches / packs) with the sections the
#    roofline needs; the report stays on the box (it exceeds the 64 MiB return limit), only its raw CSV page comes back
python scripts/profile_forward.py > $O/${R}_plain_fwd.log 2>&1 &&
SKIP=$(grep -o "skip=[0-9]*" $O/${R}_plain_fwd.log | cut -d= -f2) && COUNT=$(grep -o "count=[0-9]*" $O/${R}_plain_fwd.log | cut -d= -f2) &&
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section WarpStateStats \
    --section LaunchStats --section Occupancy --section SchedulerStats --metrics dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"$KERNELS" -s $SKIP -c $COUNT -o /tmp/${R}_forward -f \
    python scripts/profile_forward.py > $O/${R}_ncu_forward.log 2>&1 &&
ncu -i /tmp/${R}_forward.ncu-rep --page raw --csv > $O/${R}_forward_raw.csv
tail -2 $O/${R}_ncu_forward.log
if [ "${2:-all}" = "fwdonly" ]; then exit 0; fi
# 3. the dominant kernel (tcgen05 GEMM), full set: the 17 launches of one forward.  The report stays on the box (gpurun returns
#    at most 64 MiB); its raw page comes back as CSV, and the heads' per-point first-layer launch (the largest single launch:
#    index 11 of the forward's gemm_tc launches) is captured once more on its own with source correlation, small enough to keep.
python scripts/profile_forward.py > $O/${R}_plain_fwd2.log 2>&1 &&
ncu --set full --clock-control none -k regex:"^gemm_tc" -s 34 -c 17 -o /tmp/${R}_gemm_tc -f \
    python scripts/profile_forward.py > $O/${R}_ncu_gemm.log 2>&1 &&
ncu -i /tmp/${R}_gemm_tc.ncu-rep --page raw --csv > $O/${R}_gemm_tc_raw.csv
tail -2 $O/${R}_ncu_gemm.log
python scripts/profile_forward.py > $O/${R}_plain_fwd3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"^gemm_tc" -s 45 -c 1 -o $O/${R}_gemm_stage1 -f \
    python scripts/profile_forward.py > $O/${R}_ncu_stage1.log 2>&1
tail -2 $O/${R}_ncu_stage1.log
ls -la $O | tail -20
