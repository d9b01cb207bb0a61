#!/bin/bash
# ncu sections for the BACKWARD / loss / optimiser kernels of one training step (B = 64 clouds on one GPU):
#   gpurun --timeout 1200 -- 'bash scripts/profile_train.sh r02t'   ->  gpurun_out/r02t_forward_raw.csv
# (the same command runs first without ncu, as the profiling recipe requires)
set -u
R=${1:-r02t}
O=gpurun_out
mkdir -p $O
KERNELS="(bwd|chamfer|dcd|scatter_add|colsum|colred|act_bwd|bn_|affine_act|split_mixed|split_tf32_t|ranger|directions_bwd|gemm_tn|splitk)"
B=64 python scripts/train_one.py > $O/${R}_plain.log 2>&1 &&
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section WarpStateStats \
    --section LaunchStats --section Occupancy --section SchedulerStats --metrics dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"$KERNELS" -c 600 -o /tmp/${R}_train -f \
    env B=64 python scripts/train_one.py > $O/${R}_ncu.log 2>&1 &&
ncu -i /tmp/${R}_train.ncu-rep --page raw --csv > $O/${R}_forward_raw.csv
tail -2 $O/${R}_ncu.log
# launch list of the inference benchmark command (eager launches, no train record), as profile_round.sh step 1
if [ "${2:-}" = "launches" ]; then
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-train > $O/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-train > $O/r02_ncu_launches.log 2>&1
fi
