#!/usr/bin/env python
"""Turns ncu outputs under gpurun_out/ into the small text summaries committed under profiles/."""
import csv
import collections
import subprocess
import sys


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    data = [(r[ik].split("(")[0], float(r[iv].replace(",", ""))) for r in rows[1:]]
    agg = collections.OrderedDict()
    for k, v in data:
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {len(data)} launches, {tot / 1e3:.1f} us total (cold-cache, serialised: compare shares)\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:45s} launches={n:4d} total_us={v / 1e3:10.1f} share={v / tot:.3f}\n")


def full(rep, out, keys):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h = rows[0]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on: {rep}\n")
        for r in rows[2:]:
            f.write(f"## {r[h.index('Kernel Name')][:90]}  grid={r[h.index('Grid Size')]} block={r[h.index('Block Size')]}\n")
            for i, k in enumerate(h):
                if any(k.startswith(x) for x in keys):
                    f.write(f"{k} [{rows[1][i]}] = {r[i]}\n")


KEYS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct", "sm__pipe_alu_cycles_active.avg.pct", "sm__pipe_fma_cycles_active.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak", "gpu__dram_throughput", "lts__t_bytes.sum", "lts__t_sector_hit_rate",
        "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak", "sm__throughput.avg.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu",
        "l1tex__t_sector_hit_rate", "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active", "launch__shared_mem_per_block", "smsp__warp_issue_stalled"]
if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], KEYS)
