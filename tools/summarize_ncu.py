#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into profiles/ (tracked).
  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01a_launches.md "<title>"
  python tools/summarize_ncu.py full gpurun_out/prof_sweep.ncu-rep profiles/r01a_sweep_full.md "<title>"
"""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "lts__t_bytes.sum",
            "l1tex__t_bytes.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
            "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
            "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def launches(src, dst, title):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, mi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= mi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        if "at::" in name or "internal::kernel" in name:
            name = "[torch data generation] " + name[:60]
        v = float(r[mi].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0, set()])
        a[0] += 1
        a[1] += v
        a[2].add(r[gi])
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nSource: `{src}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised "
                "launches: compare SHARES, not absolutes).\n\n| kernel | launches | total us | mean us | share | grids |\n|---|---:|---:|---:|---:|---|\n")
        for k, (c, t, g) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            gs = ", ".join(sorted(g)[:4])
            f.write(f"| `{k[:100]}` | {c} | {t/1e3:.1f} | {t/c/1e3:.2f} | {t/tot*100:.1f}% | {gs} |\n")
        f.write(f"\nTotal: {sum(a[0] for a in agg.values())} launches, {tot/1e6:.3f} ms.\n")


def full(src, dst, title):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nSource: `{src}` (ncu --set full --clock-control none --import-source on), read with "
                "`ncu -i ... --page raw --csv`.\n")
        for r in rows[2:]:
            f.write(f"\n## `{r[hdr.index('Kernel Name')]}` grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in RAW_KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
