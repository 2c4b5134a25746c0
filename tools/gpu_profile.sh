#!/bin/bash
# ncu: launch list + full captures of the wavefront kernels (C3 at 2 spp)
mkdir -p gpurun_out
CMD3="python bench.py --workload c3 --spp 2 --steps 1 --warmup 1 --no-cpu-baseline --integrator 1 --flags ${FLAGS:-4}"
$CMD3 > gpurun_out/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_wf_c3.csv $CMD3 > gpurun_out/ncu_launch_c3.log 2>&1
$CMD3 > gpurun_out/plain_c3b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wf_trace -s 2 -c 2 -o gpurun_out/prof_wf_trace_c3 $CMD3 > gpurun_out/ncu_full_trace.log 2>&1
$CMD3 > gpurun_out/plain_c3c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wf_shade -s 2 -c 1 -o gpurun_out/prof_wf_shade_c3 $CMD3 > gpurun_out/ncu_full_shade.log 2>&1
tail -n 3 gpurun_out/ncu_full_trace.log gpurun_out/ncu_full_shade.log
ls -la gpurun_out
