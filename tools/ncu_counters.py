#!/usr/bin/env python
"""Condenses an .ncu-rep of the dominant kernel into profiles/r2_counters.json (read by bench.py: `roofline.ncu`).

    python tools/ncu_counters.py <workload> <rep> [rays_of_first_launch] [bytes_per_ray] [note]

For every captured launch: duration, grid, registers, occupancy, active lanes per instruction, issue-slot utilisation, L1 / L2 hit
rates, L2 sector throughput and DRAM throughput as % of peak, DRAM bytes. With the ray count of the first launch: measured DRAM
bytes per ray and their ratio to the algorithmic bytes per ray."""
import csv, io, json, os, subprocess, sys

workload, rep = sys.argv[1], sys.argv[2]
rays = float(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] else None
bpr = float(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4] else None
note = sys.argv[5] if len(sys.argv) > 5 else ""
KEYS = {"gpu__time_duration.sum": "duration", "launch__grid_size": "grid", "launch__registers_per_thread": "registers",
        "launch__shared_mem_per_block_static": "static_smem_bytes",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_instruction",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
        "sm__inst_executed.avg.per_cycle_active": "ipc",
        "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
        "lts__t_sectors.avg.pct_of_peak_sustained_elapsed": "l2_sectors_pct_of_peak",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct_of_peak",
        "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait_per_issue",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard_per_issue"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "second": 1.0, "nsecond": 1e-9}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
launches = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
    for k, name in KEYS.items():
        if k in hdr:
            v = float(r[hdr.index(k)].replace(",", ""))
            u = units[hdr.index(k)]
            if name in ("dram_read", "dram_write"):
                d[name + "_bytes"] = v * UNIT.get(u, 1.0)
            elif name == "duration":
                d["duration_ms"] = v * UNIT.get(u, 1.0) * 1e3
            elif name == "static_smem_bytes":
                d[name] = v * UNIT.get(u.split("/")[0], 1.0)
            else:
                d[name] = v
    launches.append(d)
out = {"source": os.path.basename(rep), "note": note, "launches": launches}
if rays and launches and "dram_read_bytes" in launches[0]:
    dpr = (launches[0]["dram_read_bytes"] + launches[0]["dram_write_bytes"]) / rays
    out["first_launch_rays"] = rays
    out["dram_bytes_per_ray"] = dpr
    if bpr:
        out["algorithmic_bytes_per_ray"] = bpr
        out["dram_to_algorithmic_ratio"] = dpr / bpr
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_counters.json")
allc = json.load(open(path)) if os.path.exists(path) else {}
allc[workload] = out
json.dump(allc, open(path, "w"), indent=1)
sys.stdout.write(json.dumps(out)[:600] + "\n")
