#!/usr/bin/env python
"""Turn `ncu --set full` captures (.ncu-rep, scratch under gpurun_out/) into the small markdown summaries that are kept
under profiles/.  Usage: python tools/ncu_summary.py OUT.md REP1.ncu-rep [REP2.ncu-rep ...]
Each report may hold several launches; every launch becomes one section."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("duration", "gpu__time_duration.sum"),
    ("grid / block", None),
    ("registers per thread", "launch__registers_per_thread"),
    ("dynamic + static smem per CTA", None),
    ("resident CTAs/SM limited by (regs, smem, warps)", None),
    ("achieved occupancy (warps active, % of peak)", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("SM throughput (% of peak)", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots busy (%)", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("FMA-pipe cycles active (%; IMAD.WIDE issues at half rate, so ~50 % is the ceiling)", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("ALU-pipe cycles active (%)", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("warp instructions executed", "smsp__inst_executed.sum"),
    ("active threads per warp instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("DRAM bytes read", "dram__bytes_read.sum"),
    ("DRAM bytes written", "dram__bytes_write.sum"),
    ("DRAM throughput (% of peak)", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 sector hit rate (%)", "lts__t_sector_hit_rate.pct"),
    ("L1 sector hit rate (%)", "l1tex__t_sector_hit_rate.pct"),
    ("local-memory (spill) load / store instructions", None),
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def section(d, units, rep):
    g = lambda k: d.get(k, "n/a")
    lines = [f"### `{g('Kernel Name')}`", f"source: `{rep}` (ncu --set full --clock-control none, one launch; replayed passes, so the duration is not a bench number)", "", "| metric | value |", "|---|---|"]
    for label, key in KEYS:
        if key:
            lines.append(f"| {label} | {g(key)} {units.get(key, '')} |")
        elif label.startswith("grid"):
            lines.append(f"| {label} | {g('launch__grid_size')} x {g('launch__block_size')} |")
        elif label.startswith("dynamic"):
            lines.append(f"| {label} | {g('launch__shared_mem_per_block_dynamic')} + {g('launch__shared_mem_per_block_static')} {units.get('launch__shared_mem_per_block_static','')} |")
        elif label.startswith("resident"):
            lines.append(f"| {label} | {g('launch__occupancy_limit_registers')}, {g('launch__occupancy_limit_shared_mem')}, {g('launch__occupancy_limit_warps')} |")
        elif label.startswith("local-memory"):
            lines.append(f"| {label} | {g('smsp__inst_executed_op_local_ld.sum')} / {g('smsp__inst_executed_op_local_st.sum')} |")
    stalls = []
    for k, v in d.items():
        if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
            try:
                stalls.append((float(v.replace(",", "")), k[len("smsp__pcsamp_warps_issue_stalled_"):]))
            except ValueError:
                pass
    tot = sum(x for x, _ in stalls) or 1.0
    stalls.sort(reverse=True)
    lines.append("| warp-state samples (top 6) | " + ", ".join(f"{n} {100 * x / tot:.1f} %" for x, n in stalls[:6]) + " |")
    return "\n".join(lines) + "\n"


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    parts = []
    for rep in reps:
        rows, units = raw(rep)
        for d in rows:
            parts.append(section(d, units, rep.split("/")[-1]))
    head = open(out).read().split("<!-- ncu -->")[0] if False else ""
    with open(out, "w") as f:
        f.write(head + "\n".join(parts))


if __name__ == "__main__":
    main()
