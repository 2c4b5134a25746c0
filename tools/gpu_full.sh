#!/bin/bash
# full check: gpu tests, default bench (C3 1080p 64spp + CPU baseline), reference arm, C2 bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -2 gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
python bench.py --workload c2 --steps 5 > gpurun_out/bench_c2.json 2>gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2:', d['value'], 'Mrays/s', d['ms_per_step'], 'ms', 'cpu', d['cpu_baseline'])"
python bench.py --workload c1 --steps 5 > gpurun_out/bench_c1.json 2>gpurun_out/bench_c1.err; cat gpurun_out/bench_c1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C1:', d['value'], 'Mrays/s', d['ms_per_step'], 'ms', 'cpu', d['cpu_baseline'])"
