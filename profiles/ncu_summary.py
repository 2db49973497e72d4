#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per kernel + stall breakdown + hottest SASS lines.
Usage: python profiles/ncu_summary.py report.ncu-rep [kernel-regex]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]


def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"] + (["-k", f"regex:{kre}"] if kre else [])))))
hdr, units = raw[0], raw[1]
ix = {h: i for i, h in enumerate(hdr)}
for r in raw[2:]:
    print("==", r[ix["Kernel Name"]][:100])
    for w in WANT:
        if w in ix:
            print(f"   {w} [{units[ix[w]]}] = {r[ix[w]]}")
src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv"] + (["-k", f"regex:{kre}"] if kre else [])))))
# the source page repeats a 2-line header per kernel; take the first kernel only
start = next(i for i, r in enumerate(src) if r and r[0] == "Address")
end = next((i for i in range(start + 1, len(src)) if src[i] and src[i][0] == "Kernel Name"), len(src))
h = src[start]
data = [r for r in src[start + 1:end] if len(r) == len(h)]
hx = {n: i for i, n in enumerate(h)}
S = lambda r: int(r[hx["# Samples"]]) if r[hx["# Samples"]].isdigit() else 0
tot = sum(S(r) for r in data) or 1
stall = collections.Counter()
for r in data:
    for n in h:
        if n.startswith("stall_") and "Not Issued" not in n and r[hx[n]].isdigit():
            stall[n] += int(r[hx[n]])
print("-- stall reasons (all samples) of first matching kernel, total samples", tot)
for k, v in stall.most_common(8):
    print(f"   {k:24s} {100 * v / tot:5.1f}%")
print("-- hottest SASS (>=1% of samples), program order")
for n, r in enumerate(data):
    if S(r) >= 0.01 * tot:
        print(f"   {n:5d} {100 * S(r) / tot:5.1f}%  {r[hx['Source']].strip()[:100]}")
