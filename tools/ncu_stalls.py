"""Summarise an ncu report: key raw metrics + stall reasons aggregated from the source page (SASS level)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit", "smsp__average_warp_latency_per_inst_issued.ratio"]
print(vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
for h, v in zip(hdr, vals):
    if any(h == k or h.startswith(k) for k in keys):
        print(f"  {h} = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}; data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    try: n = int(r[idx["# Samples"]])
    except ValueError: continue
    data.append((n, r))
    for s in stalls:
        try: tot[s] += int(r[idx[s]])
        except ValueError: pass
T = sum(tot.values()) or 1
print("stall samples:", T)
for s, v in sorted(tot.items(), key=lambda t: -t[1])[:8]: print(f"  {s:26s} {v:7d} {100 * v / T:5.1f}%")
print("hottest instructions:")
for n, r in sorted(data, key=lambda t: -t[0])[:topn]:
    top = sorted(((int(r[idx[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"  {n:5d} {r[idx['Source']][:70]:70s} {top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]}")
