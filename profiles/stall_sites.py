"""Hottest stall sites of one kernel from `ncu --page source --csv -k regex:<kernel>` (rows come out twice)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN = hdr.index("Source"), hdr.index("# Samples")
cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in rows[2:] if len(r) > iN and r[iN].isdigit()][::2]
tot = sum(int(r[iN]) for r in data)
agg = {c: sum(int(r[hdr.index(c)] or 0) for r in data) for c in cols}
print("samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -int(r[iN]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    st = {c: int(r[hdr.index(c)]) for c in cols if int(r[hdr.index(c)] or 0) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{int(r[iN]):6d} {100.0 * int(r[iN]) / tot:5.1f}%  {r[iS].strip()[:64]:64s} {top}")
