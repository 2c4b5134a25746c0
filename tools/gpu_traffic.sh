#!/bin/bash
# DRAM traffic + duration of every kernel of one C3 frame at 8 spp (metrics-only ncu pass), and of the C2 primary kernel
mkdir -p gpurun_out
CMD3="python bench.py --workload c3 --spp 8 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD3 > gpurun_out/plain_traffic.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none -k regex:"wf_|k_untile" --csv --log-file gpurun_out/traffic_c3_8spp.csv $CMD3 > gpurun_out/ncu_traffic.log 2>&1
tail -n 2 gpurun_out/ncu_traffic.log | cut -c1-300
wc -l gpurun_out/traffic_c3_8spp.csv
