"""Opcode mix of one kernel from `ncu --page source --csv`: executed warp-instructions and stall samples per opcode."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ex, st = Counter(), Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    toks = r[iS].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:2]) if op.split(".")[0] in ("ATOMS", "LDS", "STS", "LDG", "STG", "RED", "F2I", "I2F") else "")
    e = int(r[iE] or 0)
    ex[op] += e
    st[op] += int(r[iN] or 0)
    tot += e
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print(f"total warp-instr {tot}  (per unit: {tot / div:.1f})")
for op, e in ex.most_common(28):
    print(f"{op:14s} {e:12d} {100.0 * e / tot:6.2f}%  per-unit {e / div:8.2f}   stall-samples {st[op]}")
