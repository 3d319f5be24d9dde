#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch, --set full --import-source on) into markdown for profiles/.
usage: ncu_summary.py <report.ncu-rep> [launches.csv]"""
import collections
import csv
import subprocess
import sys

import os
rep = sys.argv[1]
# a report with several launches: NCU_LAUNCH=k selects the k-th (0-based)
sel = ["--launch-skip", os.environ["NCU_LAUNCH"], "--launch-count", "1"] if "NCU_LAUNCH" in os.environ else []
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout.splitlines()))
h, units, v = raw[0], raw[1], raw[-1]
m = {n: (v[i], units[i]) for i, n in enumerate(h)}
print(f"## {m['Kernel Name'][0][:150]}\n")
print(f"grid {m['launch__grid_size'][0]} x block {m['launch__block_size'][0]}, {m['launch__registers_per_thread'][0]} registers/thread\n")
want = [
    ("gpu__time_duration.sum", "duration"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "avg active threads / warp instruction (warp execution efficiency = x/32)"),
    ("smsp__sass_average_branch_targets_threads_uniform.pct", "branch targets uniform (SIMT divergence)"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC per SM (max 4)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_bytes.sum.per_second", "achieved L2 bandwidth"),
    ("dram__bytes_read.sum", "HBM bytes read"),
    ("dram__bytes_write.sum", "HBM bytes written"),
    ("dram__bytes.sum.per_second", "achieved HBM bandwidth"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "HBM throughput"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("sm__cycles_elapsed.max", "SM elapsed cycles (max)"),
]
print("| metric | value |\n|---|---|")
for k, label in want:
    if k in m:
        print(f"| {label} (`{k}`) | {m[k][0]} {m[k][1]} |")
print("\nstall reasons (warps stalled per issued instruction):\n")
st = sorted(((float(m[k][0]), k) for k in m if "average_warps_issue_stalled" in k and "per_issue_active.ratio" in k and "not_issued" not in k and m[k][0] not in ("", "n/a")), reverse=True)
for val, k in st[:8]:
    print(f"- {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {val:.2f}")

src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout.splitlines()))
# one section per launch: a "Kernel Name" row, the header row, then one row per SASS instruction
sections, cur_sec = [], None
for r in src:
    if r and r[0] == "Kernel Name":
        cur_sec = {"hdr": None, "rows": []}; sections.append(cur_sec)
    elif cur_sec is not None and cur_sec["hdr"] is None:
        cur_sec["hdr"] = r
    elif cur_sec is not None and len(r) == len(cur_sec["hdr"]):
        cur_sec["rows"].append(r)
sec = sections[int(os.environ.get("NCU_LAUNCH", len(sections) - 1))]
hdr, data = sec["hdr"], sec["rows"]
ci = {n: i for i, n in enumerate(hdr)}
tot_i = sum(float(r[ci["Instructions Executed"]] or 0) for r in data)
tot_s = sum(float(r[ci["# Samples"]] or 0) for r in data)
blocks, cur = [], None
for k, r in enumerate(data):
    ie = float(r[ci["Instructions Executed"]] or 0); te = float(r[ci["Thread Instructions Executed"]] or 0); sm = float(r[ci["# Samples"]] or 0)
    key = round(ie / 1e6, 1)
    if cur is None or cur["key"] != key:
        cur = {"key": key, "start": k, "n": 0, "ie": ie, "te": 0.0, "sm": 0.0, "ops": []}
        blocks.append(cur)
    cur["n"] += 1; cur["sm"] += 100 * sm / max(1, tot_s); cur["te"] += te
    cur["ops"].append(r[ci["Source"]].split()[0] if r[ci["Source"]] else "")
print("\nhot SASS regions (runs of instructions with the same execution count):\n")
print("| first SASS idx | #instr | executions (M) | avg threads | % stall samples | % of warp instructions | dominant opcodes |\n|---|---|---|---|---|---|---|")
for b in blocks:
    share = 100 * b["ie"] * b["n"] / max(1, tot_i)
    if share > 1.5 or b["sm"] > 2.0:
        ops = ", ".join(f"{o} x{c}" for o, c in collections.Counter(b["ops"]).most_common(4))
        print(f"| {b['start']} | {b['n']} | {b['ie'] / 1e6:.1f} | {b['te'] / max(1, b['ie'] * b['n']):.1f} | {b['sm']:.1f} | {share:.1f} | {ops} |")

if len(sys.argv) > 2:
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split("(")[0][:80]; val = float(r[vi].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += val
    tot = sum(x for _, x in agg.values())
    print("\n## launch list (gpu__time_duration.sum, cold-cache serialised: compare shares)\n")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (n, val) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {val / 1e6:.3f} | {100 * val / tot:.1f}% |")
