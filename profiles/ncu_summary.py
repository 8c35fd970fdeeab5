"""Summarise an .ncu-rep of the solver kernel into a small text file for profiles/ (run where ncu is installed).
Usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/solver_rXX.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
print("# ncu --set full --clock-control none, kernel:", rows[2][rows[0].index("Kernel Name")] if "Kernel Name" in rows[0] else "?")
for h, u, v in zip(rows[0], rows[1], rows[2]):
    if h in keys:
        print("%-70s %-14s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(int(r[ix[h]]) for r in data) for h in st}
T = sum(tot.values())
print("# warp stall samples (all):", ", ".join("%s %.1f%%" % (k[6:], 100 * v / T) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]))
ops = {}
for r in data:
    parts = r[ix["Source"]].split()
    o = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
    ops[o] = ops.get(o, 0) + int(r[ix["Instructions Executed"]])
S = sum(ops.values())
print("# executed warp-instructions by opcode:", ", ".join("%s %.1f%%" % (k, 100 * v / S) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
print("# TMA / mbarrier SASS present:", ", ".join(k for k in ("UTMALDG", "UBLKCP", "SYNCS", "ELECT") if k in ops))
