#!/bin/bash
mkdir -p gpurun_out
nproc; free -g | head -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
python bench.py --workload c3 --spp 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_4spp.json 2> gpurun_out/bench_c3_4spp.err; tail -2 gpurun_out/bench_c3_4spp.err; cat gpurun_out/bench_c3_4spp.json
