"""Summarise an .ncu-rep (CPU side): roofline-relevant raw metrics (DRAM bytes / throughput, tensor-pipe activity,
L2 / SM throughput, launch shape) for every captured launch + the top stalled SASS lines of the first one.
    python tools/ncu_summary.py <file.ncu-rep> [n_lines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for li, vals in enumerate(rows[2:]):
    print(f"---- launch {li}")
    for i, h in enumerate(hdr):
        if h in want:
            print(f"{h:80s} {vals[i]:>20s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
data = []
for r in rows[2:]:
    if len(r) != len(h) or not r[ix["# Samples"]].isdigit():
        break  # a second launch's table starts here
    data.append(r)
tot = sum(int(r[ix["# Samples"]]) for r in data)
ops = {}
for r in data:
    s = r[ix["Source"]]
    for m in ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "MUFU"):
        if m in s:
            ops[m] = ops.get(m, 0) + 1
print("SASS mnemonics (static count, first launch):", ops)
print("total stall samples", tot)
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(r[ix["# Samples"]].rjust(7), r[ix["Source"]][:110])
