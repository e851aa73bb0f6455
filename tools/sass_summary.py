"""Per-kernel SASS instruction counts of libmmrca.so (cuobjdump -sass): the evidence that the tile kernels are
tcgen05 / TMEM / bulk-copy code.  Runs anywhere nvcc's cuobjdump is installed (no GPU).

  python tools/sass_summary.py > profiles/sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "garbage_classification_rca_b200", "libmmrca.so")
MNEMONICS = [("UTCHMMA", "tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM)"),
             ("UTCQMMA", "tcgen05.mma kind::tf32 / i8 family"),
             ("LDTM", "tcgen05.ld (TMEM -> registers)"),
             ("STTM", "tcgen05.st (registers -> TMEM)"),
             ("UTCBAR", "tcgen05.commit -> mbarrier"),
             ("UTCATOMSWS", "tcgen05.alloc / dealloc"),
             ("UBLKCP", "cp.async.bulk (1-D bulk copy engine, global <-> shared)"),
             ("UBLKPF", "cp.async.bulk.prefetch.L2"),
             ("UTMALDG", "cp.async.bulk.tensor load (tensor-map TMA)"),
             ("UTMASTG", "cp.async.bulk.tensor store (tensor-map TMA)"),
             ("SYNCS", "mbarrier operations"),
             ("HMMA", "legacy mma.sync (must be 0)"),
             ("RED", "red.global (fire-and-forget reductions)"),
             ("FFMA", "fp32 FMA")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur).replace("mmrca::", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn, _ in MNEMONICS:
                if op.startswith(mn):
                    counts[cur][mn] += 1
    cols = [m for m, _ in MNEMONICS]
    print("# SASS summary of libmmrca.so (sm_100a), `cuobjdump -sass`, instruction counts per kernel\n")
    for mn, what in MNEMONICS:
        print(f"- `{mn}`: {what}")
    print("\n| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    tot = collections.Counter()
    for k, c in counts.items():
        print(f"| `{k}` | {c['_total']} | " + " | ".join(str(c[m]) for m in cols) + " |")
        tot.update(c)
    print(f"| **all** | {tot['_total']} | " + " | ".join(str(tot[m]) for m in cols) + " |")


if __name__ == "__main__":
    sys.exit(main())
