#!/bin/bash
# per-iteration wavefront timing (diagnostic): usage wf_timing.sh <flags> [spp]
mkdir -p gpurun_out
B200RT_WF_TIMING=1 python bench.py --workload c3 --spp ${2:-4} --steps 1 --warmup 1 --no-cpu-baseline --integrator 1 --flags $1 2> gpurun_out/wf_timing_f$1.log > /dev/null
grep "wf total" gpurun_out/wf_timing_f$1.log | tail -2
awk '/^wf it/ {n++; if (n<=45) print}' gpurun_out/wf_timing_f$1.log | tail -45
