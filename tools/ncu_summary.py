#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/<name>_launches.md
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep      > profiles/<name>_full.md
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.pct",
    "smsp__average_warp_latency_issue_stalled_barrier.pct",
]


def launches(path):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        agg.setdefault(r[ki], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | mean us | share of captured GPU time |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k[:110]}` | {len(v)} | {sum(v) / len(v):.1f} | {100 * sum(v) / tot:.1f} % |")
    print(f"\ntotal captured: {tot / 1e3:.2f} ms over {sum(len(v) for v in agg.values())} launches "
          "(ncu serialises launches and runs them cold-cache: compare shares, not absolutes)")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"### `{r[ki][:120]}`\n\n| metric | value | unit |\n|---|---:|---|")
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {m} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
