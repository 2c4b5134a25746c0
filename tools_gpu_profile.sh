#!/bin/bash
# ncu: launch list + one full capture of the dominant kernels (C3 megakernel at 2 spp, C2 primary)
mkdir -p gpurun_out
CMD3="python bench.py --workload c3 --spp 2 --steps 2 --warmup 1 --no-cpu-baseline"
CMD2="python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline"
$CMD3 > gpurun_out/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv $CMD3 > gpurun_out/ncu_launch_c3.log 2>&1
$CMD3 > gpurun_out/plain_c3b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pathtrace_mega -c 1 -o gpurun_out/prof_mega_c3 $CMD3 > gpurun_out/ncu_full_c3.log 2>&1
$CMD2 > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_primary -c 1 -o gpurun_out/prof_primary_c2 $CMD2 > gpurun_out/ncu_full_c2.log 2>&1
tail -3 gpurun_out/ncu_full_c3.log gpurun_out/ncu_full_c2.log
ls -la gpurun_out
