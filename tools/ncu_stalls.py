"""Summarise a `ncu --page source --csv` dump (tools/ncu_kernel.sh): stall reasons, per-opcode samples, hottest lines.

    python tools/ncu_stalls.py gpurun_out/<tag>_source.csv [top]
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
hdr, data = rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, n = Counter(), 0
for r in data:
    for s in stalls:
        tot[s] += int(r[idx[s]] or 0)
    n += int(r[idx["# Samples"]] or 0)
print(rows[0][1][:120])
print("samples", n)
for s, v in tot.most_common(8):
    print(f"  {s:28s} {v:8d} {100 * v / n:5.1f} %")
ops, ex = Counter(), Counter()
for r in data:
    src = r[idx["Source"]].strip().split()
    op = src[1] if src[0].startswith("@") else src[0]
    ops[op] += int(r[idx["# Samples"]] or 0)
    ex[op] += int(r[idx["Instructions Executed"]] or 0)
print("opcode: samples, warp-level executions")
for op, v in ops.most_common(top_n):
    print(f"  {op:24s} {v:8d} {ex[op]:12d}")
print("total warp-level instructions", sum(ex.values()))
print("hottest instructions")
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:top_n]:
    st = {s[6:]: int(r[idx[s]] or 0) for s in stalls}
    print(f"  {r[idx['# Samples']]:>6s} {r[idx['Instructions Executed']]:>9s}  {r[idx['Source']].strip()[:84]:84s} "
          f"{sorted(st.items(), key=lambda x: -x[1])[:2]}")
