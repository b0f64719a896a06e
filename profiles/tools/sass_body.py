"""Static opcode count of a kernel's per-particle compute body.

usage: sass_body.py <object file> <substring of the mangled kernel name> [first-marker [last-marker]]
Disassembles with cuobjdump, takes the instructions from the first <first-marker> (default LDS.128) to the next
<last-marker> (default BSYNC) and prints the opcode histogram: the body of `if (on) body.compute(...)` in walk_chunks.
"""
import re
import subprocess
import sys
from collections import Counter

obj, name = sys.argv[1], sys.argv[2]
first = sys.argv[3] if len(sys.argv) > 3 else "LDS.128"
last = sys.argv[4] if len(sys.argv) > 4 else "BSYNC"
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
sel = [f for f in funcs if name in f.split("\n", 1)[0]]
if not sel:
    sys.exit(f"no function matching {name}")
for f in sel:
    title = f.split("\n", 1)[0]
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append(m.group(2).strip())
    ops = [(i.split()[1] if i.startswith("@") else i.split()[0]) for i in ins]
    try:
        a = next(k for k, o in enumerate(ops) if o.startswith(first))
        b = next(k for k in range(a, len(ops)) if ops[k].startswith(last))
    except StopIteration:
        print(title, ": markers not found; total", len(ops))
        continue
    body = ops[a:b]
    c = Counter(o.split(".")[0] + ("." + o.split(".")[1] if o.split(".")[0] in ("LDS", "STS", "LDG", "STG", "F2I", "ATOMS", "IMAD") and len(o.split(".")) > 1 else "") for o in body)
    print(f"{title[:60]}: total {len(ops)} instrs, body [{a},{b}) = {len(body)}")
    print("   " + "  ".join(f"{k}:{v}" for k, v in c.most_common(24)))
