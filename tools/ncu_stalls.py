"""Summarise an ncu report: per kernel id, key raw metrics + stall reasons aggregated from the source page."""
import csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kid = sys.argv[3] if len(sys.argv) > 3 else None
args = ["ncu", "-i", rep, "--page", "source", "--csv"]
if kid is not None: args += ["--kernel-id", kid] if False else []
src = subprocess.run(args, capture_output=True, text=True).stdout
# the source page concatenates kernels: split on the "Kernel Name" marker rows
blocks, cur = [], []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        if cur: blocks.append(cur)
        cur = [r]
    else:
        cur.append(r)
if cur: blocks.append(cur)
for bi, rows in enumerate(blocks):
    if kid is not None and str(bi) != kid: continue
    print(f"=== kernel #{bi}: {rows[0][1][:90]}")
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
    print("stall samples:", T, " instructions:", len(data))
    print("  " + "  ".join(f"{s[6:]}={100 * v / T:.1f}%" for s, v in sorted(tot.items(), key=lambda t: -t[1])[:9]))
    for n, r in sorted(data, key=lambda t: -t[0])[:topn]:
        top = sorted(((int(r[idx[s]] or 0), s) for s in stalls), reverse=True)[:2]
        print(f"  {n:5d} {r[idx['Source']][:64]:64s} {top[0][1][6:]}={top[0][0]} {top[1][1][6:]}={top[1][0]}")
