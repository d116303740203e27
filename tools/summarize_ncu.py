"""Summarise an ncu report (.ncu-rep) into a small text file for profiles/.  Usage: summarize_ncu.py rep out.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]
with open(out, "w") as f:
    f.write(f"# ncu --set full summary of {rep.split('/')[-1]} (per launch; cold-cache, serialised replays)\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        for k in KEEP:
            if k in d:
                f.write(f"{k:90s} {d[k]:>22s} {u.get(k, '')}\n")
        f.write("# warp stall reasons (warps per issue-active cycle)\n")
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"):
                f.write(f"{k:90s} {d[k]:>22s}\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 2:
        h = srows[1]
        si, so, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
        data = srows[2:]
        op_s, op_e = collections.Counter(), collections.Counter()
        for r in data:
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[so].strip())
            if not m:
                continue
            op_s[m.group(2)] += int(r[si] or 0)
            op_e[m.group(2)] += int(r[ie] or 0)
        ts, te = sum(op_s.values()), sum(op_e.values())
        f.write(f"# SASS: {len(data)} static instructions, {te} warp-instructions executed, {ts} stall samples\n")
        f.write("# opcode            share of samples   share of executed\n")
        for o, c in op_s.most_common(16):
            f.write(f"{o:18s} {c / ts:8.3f} {op_e[o] / te:18.3f}\n")
print(open(out).read())
