#!/bin/bash
# round-2 evidence of the shipping kernels: launch list of one default frame, full captures of wf_trace_coop (58 % grid as in a frame) and wf_shade
mkdir -p gpurun_out
CMDL="python bench.py --steps 1 --warmup 1 --lean --no-cpu-baseline"
$CMDL > gpurun_out/r2i_plain_launch.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 3450 -c 3500 --csv --log-file gpurun_out/r2i_launches_default_frame.csv $CMDL > gpurun_out/r2i_ncu_launch.log 2>&1
tail -2 gpurun_out/r2i_ncu_launch.log
export B200RT_WF_GROUPS=1 B200RT_TRACE_GRID_PCT=58
CMD3="python bench.py --workload c3 --spp 2 --steps 1 --warmup 1 --lean --no-cpu-baseline"
$CMD3 > gpurun_out/r2i_plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wf_trace_coop -s 2 -c 2 -o gpurun_out/r2i_prof_wf_trace_c3 -f $CMD3 > gpurun_out/r2i_ncu_trace.log 2>&1
tail -2 gpurun_out/r2i_ncu_trace.log
ncu --set full --clock-control none --import-source on -k regex:wf_shade -s 2 -c 1 -o gpurun_out/r2i_prof_wf_shade_c3 -f $CMD3 > gpurun_out/r2i_ncu_shade.log 2>&1
tail -2 gpurun_out/r2i_ncu_shade.log
