#!/usr/bin/env python
"""Reads an .ncu-rep and aggregates the first kernel's per-source-line stall samples by source FILE and by stall reason, then
lists the hottest source lines.   python tools/ncu_lines.py <rep> [n_lines]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None
per_file = collections.defaultdict(lambda: collections.Counter()); lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or cur_file is None: continue
    if not r[0].strip().isdigit(): continue          # SASS child rows of a source line
    d = {}
    for k, v in zip(hdr, r): d.setdefault(k, v)      # ("Source" appears twice: keep the CUDA one)
    try: samples = int(d["# Samples"]); inst = int(d["Instructions Executed"])
    except Exception: continue
    c = per_file[cur_file]; c["samples"] += samples; c["inst"] += inst
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            try: c[k] += int(d[k])
            except Exception: pass
    lines.append((samples, inst, cur_file, d["Line No"], d["Source"].strip()[:110], {k: int(d[k]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and d[k].isdigit() and int(d[k])}))
tot = sum(c["samples"] for c in per_file.values()) or 1
for f, c in sorted(per_file.items(), key=lambda kv: -kv[1]["samples"]):
    reasons = ", ".join(f"{k[6:]} {100 * v / max(1, c['samples']):.0f}%" for k, v in c.most_common() if k.startswith("stall_") and v > 0.04 * c["samples"])
    print(f"{f:22s} samples {100 * c['samples'] / tot:5.1f}%  warp-inst {c['inst']:>13d}   {reasons}")
print()
for s, i, f, ln, src, st in sorted(lines, key=lambda t: -t[0])[:top]:
    top_r = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * s / tot:5.2f}%  {f}:{ln:>4s}  {src}\n        [{top_r}]")
