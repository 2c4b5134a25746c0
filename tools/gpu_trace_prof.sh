#!/bin/bash
# ncu full capture of the wavefront trace kernel only (C3 at 2 spp, one tile group); FLAGS selects the traversal layout
mkdir -p gpurun_out
export B200RT_WF_GROUPS=1
CMD3="python bench.py --workload c3 --spp 2 --steps 1 --warmup 1 --no-cpu-baseline --flags ${FLAGS:-0}"
$CMD3 > gpurun_out/plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wf_trace -s ${SKIP:-2} -c ${COUNT:-2} -o gpurun_out/prof_wf_trace_c3_f${FLAGS:-0} -f $CMD3 > gpurun_out/ncu_full_trace.log 2>&1
tail -3 gpurun_out/ncu_full_trace.log
