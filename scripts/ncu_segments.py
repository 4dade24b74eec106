#!/usr/bin/env python
"""Summarise an ncu report of the tiled sweep: headline metrics + time share per kernel phase
(source-page samples split at barriers / TMA instructions).  python scripts/ncu_segments.py <rep>"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]
r = rows[2]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-90s %s %s" % (w, r[i][:110], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, x in enumerate(rows) if x and x[0] == "Address"]
h = rows[hi[0]]
data = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
si, ie, ws = h.index("# Samples"), h.index("Instructions Executed"), h.index("L1 Wavefronts Shared")
tot = sum(int(x[si]) for x in data if x[si].isdigit())
cur = {"start": 0, "samples": 0, "inst": 0, "wave": 0, "n": 0}
segs = []
for k, x in enumerate(data):
    cur["samples"] += int(x[si]) if x[si].isdigit() else 0
    cur["inst"] += int(x[ie] or 0)
    cur["wave"] += int(x[ws] or 0)
    cur["n"] += 1
    if re.search(r"BAR\.SYNC|SYNCS|UTMALDG|UTMASTG|BRA", x[1]):
        cur["end"], cur["op"] = k, x[1].strip()[:50]
        segs.append(cur)
        cur = {"start": k + 1, "samples": 0, "inst": 0, "wave": 0, "n": 0}
segs.append(cur)
print("SASS instructions %d, samples %d" % (len(data), tot))
for s in segs:
    if s["samples"] > tot * 0.004:
        print("%5d-%5s n=%4d samples %5.1f%% warp-inst %9d smem-wavefronts %9d  %s" % (
            s["start"], s.get("end", "end"), s["n"], 100 * s["samples"] / tot, s["inst"], s["wave"], s.get("op", "")))
