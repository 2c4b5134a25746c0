#!/bin/bash
# ncu: launch list + full captures of the default kernels (C3 at 2 spp, one tile group so launches are not interleaved; C2 primary)
mkdir -p gpurun_out
export B200RT_WF_GROUPS=1
CMD3="python bench.py --workload c3 --spp 2 --steps 1 --warmup 1 --no-cpu-baseline --flags ${FLAGS:-0}"
CMD2="python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline"
$CMD3 > gpurun_out/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_wf_c3.csv $CMD3 > gpurun_out/ncu_launch_c3.log 2>&1
$CMD3 > gpurun_out/plain_c3b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wf_trace -s 2 -c 2 -o gpurun_out/prof_wf_trace_c3 $CMD3 > gpurun_out/ncu_full_trace.log 2>&1
$CMD3 > gpurun_out/plain_c3c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wf_shade -s 2 -c 1 -o gpurun_out/prof_wf_shade_c3 $CMD3 > gpurun_out/ncu_full_shade.log 2>&1
$CMD2 > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_primary -c 1 -o gpurun_out/prof_primary_c2 $CMD2 > gpurun_out/ncu_full_c2.log 2>&1
ls -la gpurun_out | tail -12
