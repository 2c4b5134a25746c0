#!/usr/bin/env python
"""Reads an .ncu-rep (needs ncu on PATH) and prints (a) headline counters per captured launch and (b) a basic-block level
breakdown of the first launch from the SASS source page: instructions executed, average active lanes, stall-sample share.

    python tools/ncu_blocks.py gpurun_out/prof_wf_trace_c3_f0.ncu-rep [min_share_pct]
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("----", r[hdr.index("Kernel Name")][:70])
    for k in KEYS:
        if k in hdr:
            print(f"  {k:92s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = []; blocks.append(cur); continue
    if cur is not None:
        cur.append(r)
b = blocks[0]; h = b[0]; data = b[1:]
ia, ie, it, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
tot_e = sum(int(r[ie]) for r in data); tot_t = sum(int(r[it]) for r in data); tot_s = sum(int(r[isamp]) for r in data)
print(f"first launch: {tot_e} warp instructions, {tot_t / tot_e:.2f} lanes on average, {tot_s} stall samples")
grp = []
for i, r in enumerate(data):
    e = int(r[ie]); t = int(r[it]); s = int(r[isamp])
    o = (i, r[ia].strip()[:50], e, t / e if e else 0.0, s)
    if grp and grp[-1][-1][2] == e and abs(grp[-1][-1][3] - o[3]) < 0.01: grp[-1].append(o)
    else: grp.append([o])
for g in grp:
    n = len(g); e = g[0][2]; s = sum(x[4] for x in g); share = n * e / tot_e * 100
    if share > min_share or s / tot_s * 100 > min_share:
        print(f"[{g[0][0]:4d}-{g[-1][0]:4d}] n={n:3d} exec={e:9d} lanes={g[0][3]:5.1f} inst%={share:5.1f} samples%={s / tot_s * 100:5.1f}  {g[0][1]}")
