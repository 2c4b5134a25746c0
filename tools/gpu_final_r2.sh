#!/bin/bash
# round-2 final check on one B200: GPU tests, the default bench line, the reference arm, the other single-GPU workloads, smoke()
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2i_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_tests.log; tail -3 gpurun_out/r2i_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2i_bench_c3.json 2> gpurun_out/r2i_bench_c3.err; echo "bench rc=$?"; tail -1 gpurun_out/r2i_bench_c3.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2i_bench_reference.json 2> gpurun_out/r2i_bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2i_bench_reference.json
for w in ${EXTRA_WORKLOADS:-c1 c2 c3k}; do timeout 300 python bench.py --workload $w --steps 5 > gpurun_out/r2i_bench_$w.json 2> gpurun_out/r2i_bench_$w.err; echo "$w rc=$?"; done
python - <<'P'
import json
for w in ("c3", "c1", "c2", "c3k"):
    try:
        d = json.loads(open(f"gpurun_out/r2i_bench_{w}.json").read())
        print(w, round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 1), " cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 3), " frac", round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(w, "failed", e)
P
