#!/bin/bash
# usage: gpu_env_sweep.sh "<ENV=val ...>|<integrator> <flags>" ...    (env SPP)
mkdir -p gpurun_out
for cfg in "$@"; do envs="${cfg%%|*}"; rest="${cfg##*|}"; set -- $rest
env $envs python bench.py --workload ${WORKLOAD:-c3} --spp ${SPP:-16} --steps 3 --warmup 2 --no-cpu-baseline --integrator $1 --flags $2 2>gpurun_out/err.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print('%-50s i$1 f$2: %.1f Mrays/s  %.2f ms/step  spp/s %.1f M' % ('$envs', d['value'], d['ms_per_step'], d['spp_per_s']/1e6))
"
tail -2 gpurun_out/err.log
done
