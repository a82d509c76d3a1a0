#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [bytes_of_columns_per_warp_unit]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
get = lambda name: (vals[hdr.index(name)], units[hdr.index(name)]) if name in hdr else ("n/a", "")
print(f"# ncu summary of {rep.split('/')[-1]}")
print("kernel:", get("Kernel Name")[0][:110])
print("grid/block:", get("Grid Size")[0], "/", get("Block Size")[0], " registers/thread:", get("launch__registers_per_thread")[0])
for key in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
            "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed_per_warp.ratio", "launch__occupancy_limit_registers",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"):
    v, u = get(key)
    print(f"{key:75s} {v:>18s} {u}")
print("stall reasons (warps per issue-active cycle):")
st = []
for i, h in enumerate(hdr):
    m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio", h)
    if m:
        try:
            st.append((float(vals[i]), m.group(1)))
        except ValueError:
            pass
for v, n in sorted(st, reverse=True)[:8]:
    print(f"   {n:28s} {v:6.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h2 = rows[1]
si, ei = h2.index("Source"), h2.index("Instructions Executed")
mix, tot = collections.Counter(), 0
for r in rows[2:]:
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[si].strip()) if len(r) > ei else None
    if not m:
        continue
    op = m.group(2)
    op = ".".join(op.split(".")[:2]) if op.startswith("IMAD") else op.split(".")[0]
    try:
        n = int(r[ei])
    except ValueError:
        continue
    mix[op] += n
    tot += n
print(f"dynamic SASS mix ({tot} warp instructions, {len(rows) - 2} static):")
for op, n in mix.most_common(12):
    print(f"   {op:12s} {100.0 * n / tot:5.1f} %")

# ---- where the warps wait: SASS instructions with the most stall samples of the top reasons ----
try:
    cols = {h: i for i, h in enumerate(h2)}
    samp = cols.get("# Samples") or cols.get("Warp Stall Sampling (All Samples)") or cols.get("Samples")
    stall_cols = [(h, i) for h, i in cols.items() if h.startswith("stall_") or "stall" in h.lower()]
    addr = cols.get("Address")
    if samp is not None:
        scored = []
        for k, r in enumerate(rows[2:]):
            try:
                scored.append((int(r[samp]), k, r))
            except (ValueError, IndexError):
                pass
        tot_s = sum(x[0] for x in scored) or 1
        print(f"top SASS instructions by stall samples ({tot_s} samples; columns: {[h for h, _ in stall_cols][:12]}):")
        for n_s, k, r in sorted(scored, reverse=True)[:30]:
            detail = " ".join(f"{h.replace('stall_', '')}={r[i]}" for h, i in stall_cols if i < len(r) and r[i] not in ("0", "", "0.00"))
            print(f"   #{k:5d} {100.0 * n_s / tot_s:5.2f}%  {r[si].strip()[:70]:70s} {detail[:150]}")
except Exception as e:   # noqa: BLE001
    print("stall attribution unavailable:", e)
