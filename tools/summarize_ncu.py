#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into tracked text files under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_<tag>.csv profiles/<tag>_launches.md
    python tools/summarize_ncu.py full     gpurun_out/prof_<tag>.ncu-rep profiles/<tag>_full.md
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows[1:]:
        ns = float(r[vi].replace(",", ""))
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1; a[1] += ns; tot += ns
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over bench.py (eager, --no-graph); per-launch times are\n"
                "cold-cache and serialised, so compare SHARES with bench.py's `kernel_classes`, not absolutes.\n\n")
        f.write(f"launches captured: {len(rows) - 1}, total {tot / 1e6:.3f} ms\n\n| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ns / 1e6:.3f} | {ns / n / 1e3:.2f} | {ns / tot:.3f} |\n")
    print(open(dst).read())


def full(src, dst):
    # src: a .ncu-rep (converted here) or the `ncu -i rep --page raw --csv` text made on the GPU box
    out = open(src).read() if src.endswith(".csv") else subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n`ncu --set full --clock-control none --import-source on`; one block per captured launch.\n")
        for r in rows[2:]:
            f.write(f"\n## {short(r[hdr.index('Kernel Name')])}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n\n")
            for k in KEEP:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"- {k}: {r[i]} {units[i]}\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
