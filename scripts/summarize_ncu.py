#!/usr/bin/env python
"""Summarise an ncu report (--set full) and an ncu launch list into markdown for profiles/.

usage: summarize_ncu.py <prof.ncu-rep> <launches.csv> <out.md> [title]
Runs on the CPU box: `ncu -i` only reads the report.
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    rep, launches, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else "ncu summary"
    md = [f"# {title}", ""]
    hdr, units, rows = raw_rows(rep)
    ki = hdr.index("Kernel Name")
    md += ["## `ncu --set full --clock-control none` (one launch per kernel)", "",
           "DRAM traffic (`dram__bytes_read.sum + dram__bytes_write.sum`) is per launch.", ""]
    for r in rows:
        md += [f"### `{r[ki]}`", "", "| metric | value |", "|---|---|"]
        for k, name in KEYS:
            if k in hdr:
                i = hdr.index(k)
                md.append(f"| {name} (`{k}`) | {r[i]} {units[i]} |")
        md.append("")
    # launch list
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    h = rows[0]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    tot = collections.OrderedDict()
    cnt = collections.Counter()
    for r in rows[1:]:
        name = r[kn].split("(")[0]
        tot[name] = tot.get(name, 0.0) + float(r[mv].replace(",", ""))
        cnt[name] += 1
    # the atomic micro-benchmark (k_mb_*, run once before the steps) is not part of a step
    step = {k: v for k, v in tot.items() if "k_mb_" not in k}
    total = sum(step.values())
    md += ["## launch list (`--metrics gpu__time_duration.sum --clock-control none`)", "",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.",
           "Shares are of the product kernels' total (the atomic micro-benchmark `k_mb_*`, which `bench.py` runs once",
           "outside the timed region, is listed without a share).", "",
           "| kernel | launches | total ns | share of the steps |", "|---|---|---|---|"]
    for name, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        share = f"{100 * v / total:.1f}%" if name in step else "—"
        md.append(f"| `{name}` | {cnt[name]} | {v:.0f} | {share} |")
    open(out, "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    main()
