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


LABELS = (("ca_bwd_kernel", "ca_bwd_bf16"), ("sa_bwd_kernel", "sa_bwd_bf16"),
          ("sa_fwd_kernel", "sa_fwd_bf16"), ("ca_fwd_kernel", "ca_fwd_bf16"), ("ce_feat_kernel", "ce_feat"),
          ("prep_feat_kernel", "prep_feat"), ("prep_kernel", "prep_bf16"), ("finalize_kernel", "finalize_bf16"),
          ("bwd_kernel", "bwd_bf16"))


def traffic(path, batch):
    """profiles/traffic.json: DRAM bytes (read + write) per launch of each kernel of the step, keyed by the
    labels bench.py uses, averaged over the captured launches."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    mult = lambda u: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    acc = collections.OrderedDict()
    for r in rows[2:]:
        label = next((lab for key, lab in LABELS if key in r[ki]), None)
        if label is None:
            continue
        b = float(r[ri].replace(",", "")) * mult(units[ri]) + float(r[wi].replace(",", "")) * mult(units[wi])
        acc.setdefault(label, []).append(b)
    print(json.dumps({"batch": int(batch), "source": path.split("/")[-1], "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch",
                      "kernels": {k: sum(v) / len(v) for k, v in acc.items()}}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else 4096)
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
