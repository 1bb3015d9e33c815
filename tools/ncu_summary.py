#!/usr/bin/env python
"""Summarise ncu captures (.ncu-rep, read with `ncu -i ... --page raw --csv`) as a markdown table, and an ncu
launch list (--metrics gpu__time_duration.sum --csv) as per-kernel totals.

    python tools/ncu_summary.py rep  gpurun_out/a.ncu-rep [more.ncu-rep ...]  > profiles/x.md
    python tools/ncu_summary.py list gpurun_out/launches.csv                   > profiles/y.csv
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = OrderedDict([
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/shared %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers"),
    ("launch__shared_mem_per_block_dynamic", "dyn. smem / CTA"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall LG throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall MIO throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
    ("lts__t_sectors_op_red.sum", "L2 reduction sectors"),
    ("lts__t_sectors_op_atom.sum", "L2 atomic sectors"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
])


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def rep(paths):
    cols = []
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {"kernel": short(r[hdr.index("Kernel Name")]), "file": p.split("/")[-1]}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    v = r[i]
                    try:
                        v = f"{float(v):.4g}"
                    except ValueError:
                        pass
                    d[k] = f"{v} {units[i]}".strip()
            cols.append(d)
    print("| metric | " + " | ".join(f"`{c['kernel']}`" for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    print("| capture | " + " | ".join(c["file"] for c in cols) + " |")
    for k, label in KEYS.items():
        if any(k in c for c in cols):
            print(f"| {label} | " + " | ".join(c.get(k, "") for c in cols) + " |")


def launch_list(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot, cnt = {}, {}
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        k = short(r[ik])
        tot[k] = tot.get(k, 0.0) + float(r[iv].replace(",", "")) / 1e3
        cnt[k] = cnt.get(k, 0) + 1
    total = sum(tot.values())
    print("kernel,launches,total_us,mean_us,share_pct")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"\"{k}\",{cnt[k]},{tot[k]:.1f},{tot[k] / cnt[k]:.2f},{100 * tot[k] / total:.2f}")


if __name__ == "__main__":
    (rep if sys.argv[1] == "rep" else launch_list)(sys.argv[2:] if sys.argv[1] == "rep" else sys.argv[2])
